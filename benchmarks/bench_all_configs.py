#!/usr/bin/env python
"""Single-GPU frame latency of every BASELINE.json config (and both rule sets), one JSON line per case.
Frame = bsplat_render_fwd (fused, one read-back), CUDA events, 512 MiB L2 flush between frames, best and median
of 10.  Secondary table for DESIGN.md; the driver's contract is bench.py."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import mojosplat_b200 as ms  # noqa: E402
from mojosplat_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
cases = [("config1_1k_256", "torch"), ("config2_100k_1080p", "torch"), ("config3_1m_1080p", "torch"),
         ("config3_1m_1080p", "gsplat"), ("config3_dense_1m_1080p", "torch"), ("config4_3m_1080p", "torch"),
         ("config5_6m_4k", "torch"), ("config5_6m_4k", "gsplat")]
for name, sem in cases:
    sc = synthetic.make_scene(name)
    g = [t.to(dev) for t in sc.gaussians()]
    bg = sc.background.to(dev)
    semv = _lib.SEM_TORCH if sem == "torch" else _lib.SEM_GSPLAT
    for _ in range(3):
        img, aux = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True)
    ts, stages = [], None
    for k in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ms.render_fused(*g, sc.camera, bg, 16, semantics=semv)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    _, info = ms.render_fused(*g, sc.camera, bg, 16, semantics=semv, return_aux=True, timing=True)
    ts.sort()
    print(json.dumps({"workload": name, "semantics": sem, "N": sc.N, "width": sc.camera.W, "height": sc.camera.H,
                      "M": info["n_isect"], "frame_ms_best": ts[0], "frame_ms_median": ts[len(ts) // 2],
                      "stage_ms": dict(zip(["projection", "depth_sort+count_scan", "emit+tile_sort", "raster"],
                                           info["stage_ms"]))}), flush=True)
    del g, img, aux
    torch.cuda.empty_cache()
