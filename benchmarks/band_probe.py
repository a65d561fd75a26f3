"""One rank's share of a row-band frame on ONE GPU (config 5 split W ways): per-band frame time, and -- under
`ncu --metrics gpu__time_duration.sum` -- the per-kernel list of that band.
    python benchmarks/band_probe.py [world=8] [rank=all]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mojosplat_b200 import parallel, synthetic
from mojosplat_b200.projection import project_gaussians_cuda

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
sc = synthetic.make_scene("config5_6m_4k")
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
cam = sc.camera
proj = project_gaussians_cuda(g[0], g[1], g[2], g[3], cam)
cost = parallel.tile_row_cost(proj[0], proj[3], cam.H, cam.W, 16).tolist()
bands = parallel.balanced_row_bands(cost, world)
out = {"bands": bands, "ms": []}
for r, band in enumerate(bands):
    if only is not None and r != only:
        continue
    rb = parallel.RowBandRenderer(sc.N, cam)
    rb.bands = [band]
    rb._resize(int(1.25 * sum(cost[band[0]:band[1]])) + 65536)
    rb.render(*g, cam, bg); rb.check()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rb.render(*g, cam, bg); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out["ms"].append(round(min(ts), 4))
    del rb
print(json.dumps(out))
