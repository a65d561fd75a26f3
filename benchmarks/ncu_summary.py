import csv, sys, subprocess, io
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [('gpu__time_duration.sum','us'),('launch__grid_size','grid'),('launch__registers_per_thread','regs'),
 ('smsp__inst_executed.sum','inst'),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),
 ('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),('sm__throughput.avg.pct_of_peak_sustained_elapsed','sm%'),
 ('dram__bytes_read.sum','rd'),('dram__bytes_write.sum','wr'),('dram__throughput.avg.pct_of_peak_sustained_elapsed','dram%'),
 ('lts__t_bytes.sum','l2bytes'),('sm__cycles_active.avg','cyc_act'),('sm__cycles_elapsed.max','cyc_el'),
 ('smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','st_long'),('smsp__warp_issue_stalled_barrier_per_warp_active.pct','st_bar'),
 ('smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','st_short'),('smsp__warp_issue_stalled_sleeping_per_warp_active.pct','st_sleep'),
 ('smsp__warp_issue_stalled_membar_per_warp_active.pct','st_membar'),('smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','st_lg'),('smsp__warp_issue_stalled_wait_per_warp_active.pct','st_wait')]
idx = {n: hdr.index(n) for n,_ in want if n in hdr}
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    name = r[ki].replace('bsplat::','').replace('void ','')[:44]
    out = [f"{name:44s}"]
    for n,lab in want:
        if n in idx:
            v = r[idx[n]]; u = units[idx[n]]
            try: f = float(v.replace(',',''))
            except: f = None
            if f is None: out.append(f"{lab}={v}")
            elif lab in ('rd','wr','l2bytes'):
                mult = {'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
                out.append(f"{lab}={f*mult/1e6:.1f}MB")
            elif lab=='us':
                mult = {'ns':1e-3,'us':1,'ms':1e3,'usecond':1,'nsecond':1e-3,'msecond':1e3}.get(u,1)
                out.append(f"t={f*mult:.1f}us")
            elif lab=='inst': out.append(f"inst={f/1e6:.2f}M")
            else: out.append(f"{lab}={f:.0f}" if abs(f)>=100 else f"{lab}={f:.1f}")
    print(' '.join(out))
