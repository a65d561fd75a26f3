import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, rasterization
dev = torch.device('cuda:0')
cfg = sys.argv[1] if len(sys.argv) > 1 else 'config3_1m_1080p'
N = int(sys.argv[2]) if len(sys.argv) > 2 else None
sc = synthetic.make_scene(cfg, N=N)
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
print(cfg, 'N', sc.N, 'M', aux['n_isect'])
t = [aux['means2d'].clone().requires_grad_(True), aux['conics'].clone().requires_grad_(True), g[4].clone().requires_grad_(True), g[3].clone().requires_grad_(True)]
gimg = torch.randn_like(img)
for k in range(4):
    for x in t: x.grad = None
    a, b, c = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    a.record()
    out = rasterization.rasterize_gaussians_diff(*t, bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16)
    b.record()
    out.backward(gimg)
    c.record(); torch.cuda.synchronize()
    print(f"train fwd {a.elapsed_time(b):.3f} ms  bwd {b.elapsed_time(c):.3f} ms")
print('fwd equals fused fast within', float((out - img).abs().max()))
print('grad norms', [float(x.grad.norm()) for x in t])
