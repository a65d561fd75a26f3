"""End-to-end (host buffers in / host image out) probes, one process per GPU:

    python benchmarks/e2e_probe.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... benchmarks/e2e_probe.py

Per rank: raw pinned H2D / D2H bandwidth (alone, both directions at once, and with all ranks copying at the same time),
the HostFramePipeline rates (every-frame upload and resident scene) with the host enqueue time per frame, and an event
timeline of the resident-scene pipeline (rasterizer end -> download start -> download end, per frame).
One JSON line per rank on stdout."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mojosplat_b200 import synthetic
from mojosplat_b200.pipeline import HostFramePipeline

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
numa = "--numa" in sys.argv
if numa:
    from mojosplat_b200 import hostmem
    hostmem.bind_to_device_node(local)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


sc = synthetic.make_scene("config3_1m_1080p")
host = [t.pin_memory() for t in sc.gaussians()]
cam0 = sc.camera
cams = synthetic.orbit_cameras(64, cam0.W, cam0.H, cam0.fx)
out = torch.empty((3, cam0.H, cam0.W, 3), dtype=torch.float32).pin_memory()
res = {"rank": rank, "world": world, "numa_bound": numa}
try:
    res["cpu_affinity"] = len(os.sched_getaffinity(0))
except Exception:
    pass

# ---- raw PCIe ----
dst = [torch.empty_like(t, device=dev) for t in host]
img = torch.empty((cam0.H, cam0.W, 3), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def bw(h2d, d2h, n=20):
    barrier()
    t0 = time.perf_counter()
    for k in range(n):
        if h2d:
            with torch.cuda.stream(s1):
                for d, h in zip(dst, host):
                    d.copy_(h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                out[k % 3].copy_(img, non_blocking=True)
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / n * 1e3


res["h2d_GBps_all_ranks_at_once"] = round(56.0e-3 / bw(True, False) * 1e3, 1)
res["d2h_GBps_all_ranks_at_once"] = round(24.9e-3 / bw(False, True) * 1e3, 1)
res["both_ms_per_frame_all_ranks_at_once"] = round(bw(True, True), 3)

# ---- pipelines ----
pipe = HostFramePipeline(dev, sc.N, cam0.W, cam0.H)
n = 100
views = [cams[k % 64] for k in range(n)]
for mode in ("every_frame", "once"):
    pipe.render(lambda k: host, views[:4], sc.background, out, upload=mode)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    stats = pipe.render(lambda k: host, views, sc.background, out, upload=mode, timeline=(mode == "once"))
    b.record()
    torch.cuda.synchronize(dev)
    res[f"fps_{mode}"] = round(n / (a.elapsed_time(b) * 1e-3), 1)
    res[f"host_enqueue_ms_per_frame_{mode}"] = round(pipe.last_host_enqueue_s / n * 1e3, 4)
    if mode == "once" and pipe.last_timeline is not None:
        tl = pipe.last_timeline
        k0 = 40  # steady state
        res["timeline_ms (frames 40..44: raster end, download start, download end, relative to frame 40's raster end)"] = [
            [round(x - tl[k0][0], 3) for x in tl[k]] for k in range(k0, k0 + 5)]
        per = [(tl[k + 1][0] - tl[k][0]) for k in range(20, n - 1)]
        res["raster_end_period_ms_mean"] = round(sum(per) / len(per), 4)
        res["download_ms_mean"] = round(sum(tl[k][2] - tl[k][1] for k in range(20, n)) / (n - 20), 4)
        res["download_wait_after_raster_ms_mean"] = round(sum(tl[k][1] - tl[k][0] for k in range(20, n)) / (n - 20), 4)
print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
