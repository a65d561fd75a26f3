import csv, sys, subprocess, io
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]
rows=[r for r in rows[2:] if len(r)==len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); iss=hdr.index('Warp Stall Sampling (All Samples)')
cols={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ie]) for r in rows); ts=sum(int(r[iss]) for r in rows)
print('total inst',tot,'samples',ts)
stall=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[cols[h]] or 0) for r in rows) for h in stall}
print({k:round(v/ts*100,1) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:9]})
top=sorted(enumerate(rows), key=lambda x:-int(x[1][iss]))[:topn]
for n,r in sorted(top):
    d={h:int(r[cols[h]]) for h in stall if r[cols[h]] not in ('','0')}
    d=dict(sorted(d.items(), key=lambda x:-x[1])[:2])
    print(f"{n:4d} {int(r[ie])/1e6:6.2f}M st={int(r[iss])/ts*100:5.2f}% {r[ia][:64]:64s} {d}")
