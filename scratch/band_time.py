import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, parallel
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
sc = synthetic.make_scene('config5_6m_4k')
g = [t.to(dev) for t in sc.gaussians()]
bg = sc.background.to(dev)
cam = sc.camera
rb = parallel.RowBandRenderer(sc.N, cam)
rb._resize(70_000_000)
for band in [(0, 135), (0, 67), (67, 135), (0, 22), (54, 67), (113, 135)]:
    rb.bands = [band]
    rb.render(*g, cam, bg); torch.cuda.synchronize()
    ts = []
    for k in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rb.render(*g, cam, bg); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(band, f"{min(ts):.3f} ms", "M_band", rb.check())
