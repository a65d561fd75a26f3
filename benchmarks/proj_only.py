"""The stand-alone projection stage (bsplat_project_fwd) a few times -- ncu target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import synthetic

sc = synthetic.make_scene(sys.argv[1] if len(sys.argv) > 1 else "config3_1m_1080p")
dev = torch.device("cuda:0")
g = [t.to(dev) for t in sc.gaussians()]
for _ in range(3):
    out = ms.project_gaussians(g[0], g[1], g[2], g[3], sc.camera)
torch.cuda.synchronize()
print("ok")
