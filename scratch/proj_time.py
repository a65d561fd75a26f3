import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic, projection
dev = torch.device('cuda:0')
for cfg in ['config3_1m_1080p', 'config5_6m_4k']:
    sc = synthetic.make_scene(cfg)
    g = [t.to(dev) for t in sc.gaussians()]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera)
    ts = []
    for k in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = projection.project_gaussians_cuda(g[0], g[1], g[2], g[3], sc.camera); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    print(cfg, f"{t*1e3:.1f} us", f"{72*sc.N/(t*1e-3)/1e9:.0f} GB/s", float(out[1].sum()))
