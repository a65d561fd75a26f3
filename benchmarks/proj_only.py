"""The stand-alone projection stage (bsplat_project_fwd) a few times -- ncu target.
    python benchmarks/proj_only.py [config] [exact|fma|fast]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mojosplat_b200 import synthetic
from mojosplat_b200.projection import project_gaussians_cuda

sc = synthetic.make_scene(sys.argv[1] if len(sys.argv) > 1 else "config3_1m_1080p")
variant = sys.argv[2] if len(sys.argv) > 2 else "exact"
dev = torch.device("cuda:0")
g = [t.to(dev) for t in sc.gaussians()]
for _ in range(3):
    out = project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera, allow_fma=variant == "fma",
                                 fast_math=variant == "fast")
torch.cuda.synchronize()
print("ok")
