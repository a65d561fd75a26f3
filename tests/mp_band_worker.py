"""Worker of tests/test_gpu_multiproc.py: one process per GPU (torch.distributed.run), NCCL.
Row-band frame with both band exchanges (NCCL all-gather; fused peer stores over NVLink) and the view split,
each compared bit for bit with the single-GPU render.  Prints "MP_OK <json>" from rank 0 on success."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    import mojosplat_b200 as ms
    from mojosplat_b200 import parallel, synthetic

    out = {}
    for cfg, N, sem in (("config5_6m_4k", 400_000, 0), ("config3_1m_1080p", 200_000, 1)):
        sc = synthetic.make_scene(cfg, N=N if rank == 0 else 1)
        shapes = [(N, 3), (N, 3), (N, 4), (N,), (N, 3)]
        g = [t.to(dev) for t in sc.gaussians()] if rank == 0 else \
            [torch.empty(s, dtype=torch.float32, device=dev) for s in shapes]
        parallel.broadcast_gaussians(g, src=0)
        cam = synthetic.make_scene(cfg, N=1).camera
        bg = torch.full((3,), 0.1, device=dev)
        ref = ms.render_fused(*g, cam, bg, 16, semantics=sem)
        for exchange in ("nccl", "p2p"):
            rb = parallel.RowBandRenderer(N, cam, semantics=sem, exchange=exchange)
            bands = rb.rebalance(g[0], g[1], g[2], g[3], cam)
            for _ in range(3):  # repeated frames reuse the peer-mapped buffers
                img = rb.render(*g, cam, bg)
                rb.check()
            same = torch.tensor([1 if torch.equal(img, ref) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)  # EVERY rank holds the full, identical frame
            assert int(same.item()) == 1, (cfg, exchange, rank)
            out[f"{cfg}/{exchange}"] = {"bands": bands, "bit_identical_on_all_ranks": True}
            del rb
        # view split: 2 views per rank, gathered, equal to local renders
        cams = synthetic.orbit_cameras(2 * world, cam.W // 4, cam.H // 4, cam.fx / 4)
        ids, full = parallel.render_views(*g, cams, bg, backend="cuda" if sem == 0 else "cuda_gsplat", gather=True)
        for k in (0, len(cams) - 1):
            loc = ms.render_fused(*g, cams[k], bg, 16, semantics=sem)
            assert torch.equal(full[k], loc), (cfg, "views", k)
        out[f"{cfg}/views"] = len(cams)
    dist.barrier()
    if rank == 0:
        print("MP_OK " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
