import sys, torch
sys.path.insert(0, '/root/repo')
import mojosplat_b200 as ms
from mojosplat_b200 import _lib, synthetic
from mojosplat_b200.pipeline import OverlappedPipeline
from ctypes import byref, c_size_t
dev = torch.device('cuda:0')
sc = synthetic.make_scene('config3_1m_1080p')
g = [t.to(dev) for t in sc.gaussians()]
cam0 = sc.camera
cams = synthetic.orbit_cameras(64, cam0.W, cam0.H, cam0.fx)
bg = sc.background.to(dev)
prio = int(sys.argv[1]) if len(sys.argv) > 1 else -1
pipe = OverlappedPipeline(dev, sc.N, cam0.W, cam0.H, slots=3)
if prio == 0:
    pipe.s_bin = torch.cuda.Stream(dev, priority=0)
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sbins = [torch.cuda.Stream(dev, priority=prio) for _ in range(nb)]
out = torch.empty((4, cam0.H, cam0.W, 3), device=dev)
n = 12
pipe.render(*g, cams[:n], bg, out=out); pipe.check()
# manual loop with timing events
E = lambda: torch.cuda.Event(enable_timing=True)
b0 = [E() for _ in range(n)]; b1 = [E() for _ in range(n)]; r1 = [E() for _ in range(n)]
infos = torch.zeros((n, 32), dtype=torch.uint8).pin_memory()
cs = [_lib.camera_struct(c) for c in cams[:n]]
torch.cuda.synchronize()
t0 = E(); t0.record(pipe.s_ras)
for sb in sbins: sb.wait_event(t0)
for k in range(n):
    slot = k % pipe.slots
    sb = sbins[k % nb]
    if k >= pipe.slots:
        sb.wait_event(r1[k - pipe.slots])
    b0[k].record(sb)
    b1[k].record(sb)
    needed = c_size_t(0)
    rc = pipe.L.bsplat_render_enqueue(pipe.N, _lib.ptr(g[0]), _lib.ptr(g[1]), _lib.ptr(g[2]), _lib.ptr(g[3]), _lib.ptr(g[4]), 3,
        __import__('ctypes').addressof(cs[k]), _lib.ptr(bg), 16, pipe.semantics, pipe.flags, _lib.ptr(out[k % 4]), _lib.ptr(pipe.ws[slot]), pipe.ws[slot].numel(),
        pipe.m_cap, byref(needed), infos[k].data_ptr(), sb.cuda_stream, pipe.s_ras.cuda_stream, b1[k].cuda_event)
    assert rc == 0
    r1[k].record(pipe.s_ras)
torch.cuda.synchronize()
for k in range(n):
    print(f"frame {k}: bin {t0.elapsed_time(b0[k]):7.3f} -> {t0.elapsed_time(b1[k]):7.3f}   raster end {t0.elapsed_time(r1[k]):7.3f}")
