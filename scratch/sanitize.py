import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, rasterization, _lib
from mojosplat_b200.pipeline import OverlappedPipeline, GraphRenderer
dev = torch.device('cuda:0')
for cfg, N, sem in [('config1_1k_256', None, 'cuda'), ('config3_1m_1080p', 30000, 'cuda'), ('config3_1m_1080p', 30000, 'cuda_gsplat'), ('config2_100k_1080p', 3000, 'cuda')]:
    sc = synthetic.make_scene(cfg, N=N)
    g = [t.to(dev) for t in sc.gaussians()]
    bg = sc.background.to(dev)
    img = ms.render_gaussians(*g, sc.camera, background_color=bg, backend=sem)
    img2, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
    img2, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
    for mode in ['fast', 'warp', 'single', 'faithful', 'fast_nocull']:
        rasterization.rasterize_gaussians_cuda(aux['means2d'], aux['conics'], g[4], g[3], bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16, mode=mode)
    t = [aux['means2d'].clone().requires_grad_(True), aux['conics'].clone().requires_grad_(True), g[4].clone().requires_grad_(True), g[3].clone().requires_grad_(True)]
    out = rasterization.rasterize_gaussians_diff(*t, bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16)
    out.sum().backward()
    pipe = OverlappedPipeline(dev, sc.N, sc.camera.W, sc.camera.H)
    cams = synthetic.orbit_cameras(4, sc.camera.W, sc.camera.H, sc.camera.fx)
    pipe.render(*g, cams, bg); pipe.check()
    small = OverlappedPipeline(dev, sc.N, sc.camera.W, sc.camera.H, m_capacity=500)
    small.render(*g, cams[:2], bg); small.check()
    coeffs = torch.randn(sc.N, 16, 3, device=dev)
    ms.eval_sh(3, coeffs, g[0], sc.camera)
    torch.cuda.synchronize()
    print(cfg, sem, 'ok', float(img.mean()))
