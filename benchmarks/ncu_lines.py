"""Per-source-line share of executed warp instructions of one kernel in an ncu report (needs --import-source on).
    python benchmarks/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kre], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
        ie = hdr.index('Instructions Executed'); st = hdr.index('Warp Stall Sampling (All Samples)')
        agg.append((int(r[ie] or 0), int(r[st] or 0), cur, int(r[0]), r[1].strip()[:100]))
tot = sum(a[0] for a in agg); ts = sum(a[1] for a in agg) or 1
print('warp instructions', tot, 'stall samples', ts)
for a in sorted(agg, reverse=True)[:topn]:
    print(f"{a[0]/tot*100:5.1f}% inst {a[1]/ts*100:5.1f}% stall  {a[2]}:{a[3]}  {a[4]}")
