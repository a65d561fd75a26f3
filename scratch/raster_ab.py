import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, rasterization
dev = torch.device('cuda:0')
cfg = sys.argv[1] if len(sys.argv) > 1 else 'config3_1m_1080p'
sc = synthetic.make_scene(cfg)
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ref = None
for mode in ['fast', 'mbar']:
    ts = []
    for k in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = rasterization.rasterize_gaussians_cuda(aux['means2d'], aux['conics'], g[4], g[3], bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16, mode=mode)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None: ref = out
    print(f"{mode:12s} {min(ts[2:]):.4f} ms  maxdiff_vs_first={float((out-ref).abs().max()):.3e} frac>1e-4={float(((out-ref).abs()>1e-4).float().mean()):.2e}")
th = aux['tile_ranges'].shape[0]
for mode in ['fast', 'mbar']:
    for rows in [(0, th), (1, th - 1), (0, 1)]:
        ts = []
        out = torch.zeros_like(img)
        for k in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rasterization.rasterize_gaussians_cuda(aux['means2d'], aux['conics'], g[4], g[3], bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16, mode=mode, tile_rows=rows, out=out)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(mode, rows, f"{min(ts[2:]):.4f} ms")
