"""Per-stage device times of one fused frame (L2 flushed between frames) and of the stand-alone rasterizer call.
    python benchmarks/stage_probe.py [config] [semantics] [packed]
A/B knobs of the rasterizer launch go through the BSPLAT_DEBUG environment variable (noprepass, roworder)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import _lib, rasterization, synthetic

cfg = sys.argv[1] if len(sys.argv) > 1 else "config3_1m_1080p"
sem = _lib.SEM_GSPLAT if len(sys.argv) > 2 and sys.argv[2] == "gsplat" else _lib.SEM_TORCH
packed = len(sys.argv) > 3 and sys.argv[3] == "packed"
dev = torch.device("cuda:0")
sc = synthetic.make_scene(cfg)
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
stage = np.zeros(4)
n = 12
for k in range(n + 3):
    flush.zero_()
    _, info = ms.render_fused(*g, sc.camera, bg, 16, semantics=sem, timing=True, packed=packed)
    if k >= 3:
        stage += np.array(info["stage_ms"])
stage /= n
ms.render_fused(*g, sc.camera, bg, 16, semantics=sem, return_aux=True)
img, aux = ms.render_fused(*g, sc.camera, bg, 16, semantics=sem, return_aux=True)
ts = []
for k in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    img2 = rasterization.rasterize_gaussians_cuda(aux["means2d"], aux["conics"], g[4], g[3], bg, aux["tile_ranges"],
                                                  aux["sorted_ids"], sc.camera, 16)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
tp = []
out = ms.projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera, semantics=sem)
for k in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ms.projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera, semantics=sem, out=out)
    b.record()
    torch.cuda.synchronize()
    tp.append(a.elapsed_time(b))
tf = []
for k in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ms.projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera, semantics=sem, out=out, allow_fma=True)
    b.record()
    torch.cuda.synchronize()
    tf.append(a.elapsed_time(b))
tq = []
for k in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ms.projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera, semantics=sem, out=out, fast_math=True)
    b.record()
    torch.cuda.synchronize()
    tq.append(a.elapsed_time(b))
print(json.dumps({"config": cfg, "projection_alone_ms": round(float(np.median(tp)), 4),
                  "projection_alone_fma_build_ms": round(float(np.median(tf)), 4),
                  "projection_alone_fast_math_ms": round(float(np.median(tq)), 4), "semantics": sem, "packed": packed, "debug": os.environ.get("BSPLAT_DEBUG", ""),
                  "M": info["n_isect"], "stage_ms": [round(float(x), 4) for x in stage],
                  "frame_ms": round(float(stage.sum()), 4),
                  "standalone_raster_call_ms (tile order + record kernel + raster)": round(float(np.median(ts)), 4),
                  "same_image": bool(torch.equal(img, img2)),
                  "image_sha1": __import__("hashlib").sha1(img.cpu().numpy().tobytes()).hexdigest()[:16]}))
