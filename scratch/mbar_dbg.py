import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, rasterization
dev = torch.device('cuda:0')
for cfg, N in [('config1_1k_256', None), ('config3_1m_1080p', 100000), ('config3_1m_1080p', None)]:
    sc = synthetic.make_scene(cfg, N=N)
    g = [t.to(dev) for t in sc.gaussians()]
    bg = sc.background.to(dev)
    img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
    img, aux = ms.render_fused(*g, sc.camera, bg, 16, return_aux=True)
    args = (aux['means2d'], aux['conics'], g[4], g[3], bg, aux['tile_ranges'], aux['sorted_ids'], sc.camera, 16)
    a = rasterization.rasterize_gaussians_cuda(*args, mode='fast')
    b = rasterization.rasterize_gaussians_cuda(*args, mode='mbar')
    bad = ((a - b).abs() > 1e-6) | ~torch.isfinite(b)
    badpix = bad.any(-1)
    H, W = badpix.shape
    th, tw = aux['tile_ranges'].shape[:2]
    pad = torch.zeros((th * 16, tw * 16), dtype=torch.bool, device=dev); pad[:H, :W] = badpix
    tiles_bad = pad.reshape(th, 16, tw, 16).permute(0, 2, 1, 3).reshape(th, tw, 256).any(-1)
    L = (aux['tile_ranges'][..., 1] - aux['tile_ranges'][..., 0])
    print(cfg, N, 'bad pixels', int(badpix.sum()), 'bad tiles', int(tiles_bad.sum()), 'of', th * tw)
    if tiles_bad.any():
        lb = L[tiles_bad]
        print('  list lengths of bad tiles: min', int(lb.min()), 'median', int(lb.median()), 'max', int(lb.max()))
        lg = L[~tiles_bad]
        print('  list lengths of good tiles: min', int(lg.min()), 'median', int(lg.median()), 'max', int(lg.max()))
        # sub-block pattern in the first bad tile
        idx = tiles_bad.nonzero()[0]
        ty, tx = int(idx[0]), int(idx[1])
        blk = pad[ty*16:(ty+1)*16, tx*16:(tx+1)*16].int()
        print('  first bad tile', ty, tx, 'len', int(L[ty, tx])); print(blk.cpu().numpy())
