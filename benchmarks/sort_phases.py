"""Where a onesweep pass spends its time: per-CTA SM clocks at the phase boundaries (A/B build with -DBSPLAT_PHASES,
mojosplat_b200/csrc/ab/libbsplat_phases.so; see radix_sort.cu).
    BSPLAT_LIB=$PWD/mojosplat_b200/csrc/ab/libbsplat_phases.so python benchmarks/sort_phases.py [config]"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import _lib, synthetic

cfg = sys.argv[1] if len(sys.argv) > 1 else "config3_1m_1080p"
dev = torch.device("cuda:0")
sc = synthetic.make_scene(cfg)
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for k in range(4):
    flush.zero_()
    _, info = ms.render_fused(*g, sc.camera, bg, 16, timing=True)
torch.cuda.synchronize()
L = _lib.load()
buf = np.zeros((8, 4096, 10), dtype=np.uint64)
rc = L.bsplat_debug_phases(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes))
assert rc == 0, rc
names = ["entry->dep wait", "load+rank", "publish+scan", "smem scatter", "look-back", "store"]
print("stage_ms", info["stage_ms"], "M", info["n_isect"])
for slot in range(8):
    b = buf[slot]
    rows = np.nonzero(b[:, 9])[0]
    if len(rows) == 0:
        continue
    b = b[rows].astype(np.int64)
    g0, g1 = b[:, 8], b[:, 9]
    print(f"pass slot {slot}: {len(rows)} tiles, kernel span {(g1.max() - g0.min()) / 1e3:.1f} us; first CTA entry -> last "
          f"CTA entry {(g0.max() - g0.min()) / 1e3:.1f} us; CTA lifetime median {np.median(g1 - g0) / 1e3:.1f} us "
          f"max {(g1 - g0).max() / 1e3:.1f} us")
    for i, nm in enumerate(names):
        d = (b[:, i + 1] - b[:, i]) / 1965.0  # us at 1965 MHz
        print(f"    {nm:18s} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us")
    # tiles by entry time: how late do late tiles start, when do they end
    order = np.argsort(g0)
    q = [0, len(order) // 4, len(order) // 2, 3 * len(order) // 4, len(order) - 1]
    print("    entry/exit (us since first entry) of tiles at quantiles of entry time:",
          [(round((g0[order[i]] - g0.min()) / 1e3, 1), round((g1[order[i]] - g0.min()) / 1e3, 1)) for i in q])

# count + scan (bin_count_scan2_kernel): thread 0 = the look-back path, first thread of the last warp = the coverage path
sb = np.zeros((1024, 16), dtype=np.uint64)
if hasattr(L, "bsplat_debug_scan_phases") and L.bsplat_debug_scan_phases(sb.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(sb.nbytes)) == 0:
    rows = np.nonzero(sb[:, 7])[0]
    b = sb[rows].astype(np.int64)
    print(f"count + scan: {len(rows)} CTAs")
    for nm, i0, i1 in [("entry->dep wait", 0, 1), ("index loads + gathers + rect stores", 1, 2), ("warp 0 coverage", 2, 3),
                       ("block scan (barrier)", 3, 4), ("look-back (warp 0)", 4, 5), ("barrier after look-back", 5, 6),
                       ("offset stores", 6, 7), ("[last warp] gathers -> after block scan", 8, 9),
                       ("[last warp] coverage", 9, 10), ("[last warp] coverage barrier", 10, 11),
                       ("[last warp] prefix + flush of the difference arrays", 11, 12)]:
        d = (b[:, i1] - b[:, i0]) / 1965.0
        print(f"    {nm:52s} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us")
