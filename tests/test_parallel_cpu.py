"""CPU: host logic of the multi-GPU paths (view split, row-band balancing, band / view assembly) with
world_size-2 gloo process groups.  The render itself is replaced by a deterministic stand-in; the GPU
versions are covered by tests/test_gpu_parallel.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mojosplat_b200 import parallel


def test_split_views_partition():
    for n, w in [(64, 1), (64, 2), (64, 8), (5, 8), (0, 4)]:
        seen = sorted(v for r in range(w) for v in parallel.split_views(n, r, w))
        assert seen == list(range(n))
        sizes = [len(parallel.split_views(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_balanced_row_bands():
    cost = [1.0] * 135
    for w in (1, 2, 4, 8):
        b = parallel.balanced_row_bands(cost, w)
        assert b[0][0] == 0 and b[-1][1] == 135 and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) <= 1
    # concentrated cost: bands follow the cost, stay contiguous and cover everything
    cost = [0.0] * 20 + [100.0] * 10 + [1.0] * 38
    b = parallel.balanced_row_bands(cost, 4)
    assert b[0][0] == 0 and b[-1][1] == 68 and all(b[i][1] == b[i + 1][0] for i in range(3))
    sums = [sum(cost[s:e]) for s, e in b]
    assert max(sums) <= 0.5 * sum(cost)
    # more ranks than rows: empty bands allowed, still a partition
    b = parallel.balanced_row_bands([1.0, 1.0], 4)
    assert b[0][0] == 0 and b[-1][1] == 2 and sum(e - s for s, e in b) == 2


def test_tile_row_cost_matches_exact_counts():
    from oracle import oracle
    g = torch.Generator().manual_seed(1)
    N, H, W = 500, 270, 480
    m2 = torch.rand(N, 2, generator=g) * torch.tensor([W, H])
    rad = (torch.rand(N, 2, generator=g) * 40).floor().int()
    M, counts = oracle.bin_count(m2.numpy(), rad.numpy(), W, H, 16)
    cost = parallel.tile_row_cost(m2, rad, H, W, 16)
    assert abs(float(cost.sum()) - M) <= 1e-3 * M


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- broadcast of the Gaussian arrays
        t = [torch.arange(12, dtype=torch.float32).reshape(4, 3) * (1 if rank == 0 else 0), torch.ones(4) * rank]
        parallel.broadcast_gaussians(t, src=0)
        assert torch.equal(t[0], torch.arange(12, dtype=torch.float32).reshape(4, 3)) and (t[1] == 0).all()

        # --- view split with gather: stand-in renderer paints each view with its index
        class Cam:  # only H, W are read by the gather path
            def __init__(self, k): self.H, self.W, self.k = 4, 6, k
        cams = [Cam(k) for k in range(5)]
        feats = torch.zeros(1, 3)
        fake = lambda cs: torch.stack([torch.full((4, 6, 3), float(c.k)) for c in cs])
        ids, full = parallel.render_views(None, None, None, None, feats, cams, gather=True, render_fn=fake)
        assert ids == list(range(5)) and full.shape == (5, 4, 6, 3)
        for k in range(5):
            assert (full[k] == k).all()
        mine, imgs = parallel.render_views(None, None, None, None, feats, cams, gather=False, render_fn=fake)
        assert mine == parallel.split_views(5, rank, world) and imgs.shape[0] == len(mine)

        # --- row-band assembly: every rank fills only its band of a 70-row image (tile size 16)
        bands = parallel.balanced_row_bands([3.0, 1.0, 1.0, 1.0, 2.0], world)  # 5 tile rows, H = 70
        img = torch.zeros(70, 8, 3)
        s, e = bands[rank][0] * 16, min(bands[rank][1] * 16, 70)
        img[s:e] = torch.arange(s, e, dtype=torch.float32)[:, None, None] + 1
        full = parallel.assemble_bands(img, bands, 16, rank, world)
        assert torch.equal(full[:, 0, 0], torch.arange(70, dtype=torch.float32) + 1)
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_view_split_and_band_assembly():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
