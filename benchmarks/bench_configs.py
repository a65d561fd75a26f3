#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 on 1/2/4/8 B200 -- not the driver's bench.py contract,
an additional measurement whose JSON lines are kept under profiles/.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_configs.py --config 4
    torchrun ...                                                    benchmarks/bench_configs.py --config 5

config 4: 3 M Gaussians x 64 views @1080p. Gaussians are drawn on rank 0 and NCCL-broadcast once (timed),
          views are split round-robin, no data-path collective; value = views/s (all ranks, max device time).
config 5: 6 M Gaussians, one 3840x2160 frame split into tile-row bands balanced by intersection count; every
          rank projects all N, bins + rasterizes its band, one all-gather of the bands; value = frame
          latency (ms, max over ranks) and the image is compared bit for bit with rank 0's single-GPU render.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--n-gaussians", type=int, default=None)
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION/INFO makes NCCL print a banner on stdout; the contract is one JSON line there
        # (the version banner is printed at every level but NONE: send NCCL's log to stderr instead)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    import mojosplat_b200 as ms
    from mojosplat_b200 import parallel, synthetic
    from mojosplat_b200.pipeline import OverlappedPipeline

    name = "config4_3m_1080p" if args.config == 4 else "config5_6m_4k"
    N = synthetic.CONFIGS[name][0] if args.n_gaussians is None else args.n_gaussians
    sc = synthetic.make_scene(name, N=N if rank == 0 else 1)
    cam0 = sc.camera
    W, H = cam0.W, cam0.H
    shapes = [(N, 3), (N, 3), (N, 4), (N,), (N, 3)]
    g = [t.to(dev) for t in sc.gaussians()] if rank == 0 else \
        [torch.empty(s, dtype=torch.float32, device=dev) for s in shapes]
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if world > 1:
        parallel.broadcast_gaussians(g, src=0)
    torch.cuda.synchronize(dev)
    bcast_ms = 1e3 * (time.perf_counter() - t0)
    bg = sc.background.to(dev)

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    line = {"config": args.config, "workload": name, "N": N, "n_gpus": world, "width": W, "height": H,
            "broadcast_ms": bcast_ms, "broadcast_bytes": 56 * N}
    if args.config == 4:
        cams = synthetic.orbit_cameras(args.views, W, H, cam0.fx)
        mine = parallel.split_views(len(cams), rank, world)
        my_cams = [cams[v] for v in mine]
        pipe = OverlappedPipeline(dev, N, W, H, slots=3, bin_streams=2, m_capacity=6 * N)
        ring = torch.empty((4, H, W, 3), dtype=torch.float32, device=dev)
        pipe.render(*g, my_cams[:4], bg, out=ring); pipe.check()
        best = 1e30
        for rep in range(max(1, args.steps // 5)):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pipe.render(*g, my_cams, bg, out=ring)
            b.record()
            torch.cuda.synchronize(dev)
            assert pipe.check() == 0
            best = min(best, maxr(a.elapsed_time(b)))
        line.update({"metric": "views/s, 3M Gaussians x 64 views @1920x1080", "value": len(cams) / (best * 1e-3),
                     "unit": "views/s", "ms_total": best, "views": len(cams), "views_per_rank": len(mine),
                     "M_max": pipe.last_M, "scaling": "strong (64 views fixed)",
                     "timing": "CUDA events around this rank's views, max over ranks, best of repeats"})
    else:
        ref = ms.render_fused(*g, cam0, bg, 16) if rank == 0 else None
        rb = parallel.RowBandRenderer(N, cam0, exchange=args.exchange)
        bands = rb.rebalance(g[0], g[1], g[2], g[3], cam0)   # once per sequence, not per frame
        img = rb.render(*g, cam0, bg); rb.check()            # warm-up
        ts = []
        for rep in range(args.steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            img = rb.render(*g, cam0, bg)
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(maxr(a.elapsed_time(b)))
        m_band = rb.check()
        same = bool(torch.equal(img, ref)) if rank == 0 else None
        line.update({"bands": bands, "M_band_rank0": m_band, "exchange": rb.exchange})
        # single-GPU fused frame for comparison (rank 0)
        t1 = None
        if rank == 0:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                ms.render_fused(*g, cam0, bg, 16)
            b.record(); torch.cuda.synchronize(dev)
            t1 = a.elapsed_time(b) / 3
        line.update({"metric": "frame latency, 6M Gaussians @3840x2160, tile-row bands", "value": min(ts),
                     "unit": "ms", "median_ms": sorted(ts)[len(ts) // 2], "higher_is_better": False,
                     "bit_identical_to_single_gpu": same, "single_gpu_fused_ms": t1,
                     "timing": "CUDA events around one sync-free band frame (project all, bin + rasterize the band) + the band exchange, max over ranks"})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
