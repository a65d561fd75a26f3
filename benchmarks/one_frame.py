import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic
dev = torch.device('cuda:0')
sc = synthetic.make_scene(sys.argv[1] if len(sys.argv) > 1 else 'config3_1m_1080p')
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for k in range(n):
    img = ms.render_fused(*g, sc.camera, bg, 16)
# ... and the stand-alone projection stage (the 72 B / Gaussian kernel of the HBM roofline)
ms.project_gaussians(g[0], g[1], g[2], g[3], sc.camera)
torch.cuda.synchronize()
print("ok")
