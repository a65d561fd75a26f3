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
    sys.path.insert(0, str(ROOT / "benchmarks"))
    import sharded
    if args.config == 4:
        line = sharded.run_config4(dev, rank, world, args.n_gaussians, args.views, reps=max(1, args.steps // 5))
        line.update({"config": 4, "metric": "views/s, 3M Gaussians x 64 views @1920x1080", "value": line["views_per_s"],
                     "unit": "views/s"})
    else:
        line = sharded.run_config5(dev, rank, world, args.n_gaussians, steps=args.steps, exchanges=(args.exchange,))
        key = "latency_ms" if world == 1 else f"latency_ms_{args.exchange}"
        line.update({"config": 5, "metric": "frame latency, 6M Gaussians @3840x2160, tile-row bands", "value": line[key],
                     "unit": "ms", "higher_is_better": False, "exchange": args.exchange if world > 1 else "none",
                     "bit_identical_to_single_gpu": line["bit_identical"]})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
