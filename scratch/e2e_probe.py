import sys, time, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import synthetic
from mojosplat_b200.pipeline import HostFramePipeline
dev = torch.device('cuda:0')
sc = synthetic.make_scene('config3_1m_1080p')
host = [t.pin_memory() for t in sc.gaussians()]
cam0 = sc.camera
cams = synthetic.orbit_cameras(64, cam0.W, cam0.H, cam0.fx)
out = torch.empty((3, cam0.H, cam0.W, 3), dtype=torch.float32).pin_memory()
def run(pipe, n=30, **kw):
    pipe.render(lambda k: host, cams[:4], sc.background, out, **kw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pipe.render(lambda k: host, [cams[k % 64] for k in range(n)], sc.background, out, **kw); b.record(); torch.cuda.synchronize()
    return n / (a.elapsed_time(b) * 1e-3)
for ins in (2, 3, 4):
    p = HostFramePipeline(dev, sc.N, cam0.W, cam0.H, in_slots=ins)
    print('in_slots', ins, 'fps', round(run(p)))
    del p
# raw PCIe: H2D only, D2H only, both
dst = [torch.empty_like(t, device=dev) for t in host]
img = torch.empty((cam0.H, cam0.W, 3), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bw(h2d, d2h, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if h2d:
            with torch.cuda.stream(s1):
                for d, h in zip(dst, host): d.copy_(h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                out[k % 3].copy_(img, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    return dt * 1e3
print('H2D only ms/frame', round(bw(True, False), 3), ' -> GB/s', round(56.0e-3 / bw(True, False) * 1e3, 1))
print('D2H only ms/frame', round(bw(False, True), 3), ' -> GB/s', round(24.9e-3 / bw(False, True) * 1e3, 1))
print('both ms/frame', round(bw(True, True), 3))
big_h = torch.empty(14_000_003, dtype=torch.float32).pin_memory(); big_d = torch.empty_like(big_h, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(30): big_d.copy_(big_h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 30
print('single 56 MB H2D ms', round(dt * 1e3, 3), 'GB/s', round(56.0e-3 / dt, 1))
