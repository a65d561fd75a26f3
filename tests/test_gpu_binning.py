"""GPU parity: binning (count+scan, emit, onesweep sort, tile ranges) -- bit-exact."""
import numpy as np
import pytest
import torch

import mojosplat_b200 as ms
from conftest import canonical_tie_order_equal, load_golden
from helpers import dev, oracle_project_scene
from mojosplat_b200 import _lib, binning, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu
CASES = ["config1_1k_256", "garden_6k_1080p", "dense_300_1080p", "teststyle_500_identity",
         "teststyle_500_offset", "odd_800_200x120_ts10", "bin_simple", "bin_edge", "bin_50_ts8",
         "bin_50_ts16", "bin_50_ts32", "bin_ties_400", "bin_empty"]


@pytest.mark.parametrize("algo", ["two_level", "single"])
@pytest.mark.parametrize("name", CASES)
def test_binning_vs_reference_golden(cuda_device, name, algo):
    g = load_golden(name)
    W, H = [int(v) for v in g["size"]]
    ts = int(g["tile_size"])
    if algo == "two_level":  # the public dispatcher (default algorithm)
        ids, ranges = ms.bin_gaussians_to_tiles(dev(g["means2d"], cuda_device), dev(g["radii"], cuda_device),
                                                dev(g["depths"], cuda_device), H, W, ts, backend="cuda")
    else:
        ids, ranges = binning.bin_gaussians_to_tiles_cuda(dev(g["means2d"], cuda_device), dev(g["radii"], cuda_device),
                                                          dev(g["depths"], cuda_device), H, W, ts, algo="single")
    assert ids.dtype == torch.int32 and ranges.dtype == torch.int32
    assert tuple(ranges.shape) == g["tile_ranges"].shape
    ids, ranges = ids.cpu().numpy(), ranges.cpu().numpy()
    assert np.array_equal(ranges, g["tile_ranges"])          # bit-exact vs the reference
    assert canonical_tie_order_equal(ids, g["sorted_ids"], ranges, g["depths"])
    # and exactly the oracle's canonical (stable) order
    o_ids, o_ranges = oracle.bin_tiles(g["means2d"], g["radii"], g["depths"], H, W, ts)
    assert np.array_equal(ids, o_ids) and np.array_equal(ranges, o_ranges)


@pytest.mark.parametrize("cfg,N,sem", [("config2_100k_1080p", 100_000, 0), ("config3_1m_1080p", 1_000_000, 0),
                                       ("config3_1m_1080p", 300_000, 1)])
def test_binning_full_size_bit_exact_vs_oracle(cuda_device, cfg, N, sem):
    sc = synthetic.make_scene(cfg, N=N)
    m2, con, dep, rad = oracle_project_scene(sc)
    cam = sc.camera
    o_ids, o_ranges, o_keys = oracle.bin_tiles(m2, rad, dep, cam.H, cam.W, 16, semantics=sem, return_keys=True)
    backend = "cuda" if sem == 0 else "cuda_gsplat"
    ids, ranges, keys, layout = binning.bin_gaussians_to_tiles_cuda(
        dev(m2, cuda_device), dev(rad, cuda_device), dev(dep, cuda_device), cam.H, cam.W, 16,
        semantics=sem, return_keys=True)
    assert ids.numel() == o_ids.shape[0]
    assert np.array_equal(ranges.cpu().numpy(), o_ranges)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    # the default two-level algorithm gives the same lists, bit for bit
    ids2, ranges2 = binning.bin_gaussians_to_tiles_cuda(
        dev(m2, cuda_device), dev(rad, cuda_device), dev(dep, cuda_device), cam.H, cam.W, 16, semantics=sem)
    assert torch.equal(ids2, ids) and torch.equal(ranges2, ranges)
    # size-independent properties: keys sorted, tile field consistent with the ranges
    k = keys.cpu().numpy().astype(np.uint64)
    assert (np.diff(k.astype(np.int64)) >= 0).all()
    tiles = (k >> np.uint64(layout.depth_bits)).astype(np.int64)
    assert np.array_equal(tiles, (o_keys >> np.uint64(32)).astype(np.int64))
    if cfg == "config3_1m_1080p" and N == 1_000_000 and sem == 0:
        assert abs(ids.numel() - 4_214_053) <= 8  # SURVEY 8d probe (made with the torch projection)


def test_binning_row_bands_match_full_frame(cuda_device):
    sc = synthetic.make_scene("config2_100k_1080p", N=20_000)
    m2, con, dep, rad = oracle_project_scene(sc)
    cam = sc.camera
    a = [dev(x, cuda_device) for x in (m2, rad, dep)]
    ids, ranges = binning.bin_gaussians_to_tiles_cuda(*a, cam.H, cam.W, 16, algo="single")
    ids, ranges = ids.cpu().numpy(), ranges.cpu().numpy()
    th = ranges.shape[0]
    for r0, r1 in [(0, 17), (17, 40), (40, th)]:
        b_ids, b_ranges = binning.bin_gaussians_to_tiles_cuda(*a, cam.H, cam.W, 16, tile_rows=(r0, r1))
        b_ids, b_ranges = b_ids.cpu().numpy(), b_ranges.cpu().numpy()
        for ty in range(r0, r1):
            for tx in range(0, ranges.shape[1], 7):
                s, e = ranges[ty, tx]; bs, be = b_ranges[ty, tx]
                assert np.array_equal(ids[s:e], b_ids[bs:be])
        outside = np.ones(th, bool); outside[r0:r1] = False
        assert (b_ranges[outside, :, 0] == b_ranges[outside, :, 1]).all()


def test_binning_structure_invariants(cuda_device):
    """tests/test_binning.py:78-100,150-165 of the reference."""
    g = torch.Generator().manual_seed(0)
    N, H, W = 300, 128, 128
    m2 = (torch.rand(N, 2, generator=g) * 128).to(cuda_device)
    rad = (torch.rand(N, 2, generator=g) * 10 + 5).to(cuda_device)
    dep = (torch.rand(N, generator=g) * 5 + 1).to(cuda_device)
    ids, ranges = ms.bin_gaussians_to_tiles(m2, rad, dep, H, W, 16)
    M = ids.numel()
    assert (ranges[..., 0] <= ranges[..., 1]).all() and ranges.max().item() <= M and M > N
    assert ids.min().item() >= 0 and ids.max().item() < N
    flat = ranges.reshape(-1, 2).cpu().numpy()
    assert flat[0, 0] == 0 and flat[-1, 1] == M and (flat[1:, 0] == flat[:-1, 1]).all()
    d = dep.cpu().numpy(); idn = ids.cpu().numpy()
    for s, e in flat:
        assert (np.diff(d[idn[s:e]]) >= 0).all()  # front-to-back inside every tile
    with pytest.raises(ValueError, match="Invalid backend"):
        ms.bin_gaussians_to_tiles(m2, rad, dep, H, W, 16, backend="nope")


@pytest.mark.parametrize("M", [0, 1, 31, 3071, 3072, 3073, 100_000, 2_500_000])
@pytest.mark.parametrize("bits", [(0, 8), (0, 13), (0, 40), (3, 45), (0, 64)])
def test_radix_sort_pairs(cuda_device, M, bits):
    from ctypes import byref, c_int32
    L = _lib.require_device(cuda_device)
    b0, b1 = bits
    g = torch.Generator().manual_seed(M + b1)
    keys = torch.randint(-(2 ** 63), 2 ** 63 - 1, (M,), generator=g, dtype=torch.int64)
    if b1 - b0 <= 13:
        pass  # many duplicates by construction (few live bits) -> exercises stability
    vals = torch.arange(M, dtype=torch.int32)
    ku = keys.numpy().astype(np.uint64)
    field = (ku >> np.uint64(b0)) & np.uint64((1 << (b1 - b0)) - 1 if b1 - b0 < 64 else 0xFFFFFFFFFFFFFFFF)
    order = np.argsort(field, kind="stable")
    k, ka = keys.to(cuda_device), torch.empty(M, dtype=torch.int64, device=cuda_device)
    v, va = vals.to(cuda_device), torch.empty(M, dtype=torch.int32, device=cuda_device)
    nbytes = L.bsplat_radix_sort_workspace_bytes(M, b0, b1)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=cuda_device)
    in_alt = c_int32(0)
    rc = L.bsplat_radix_sort_pairs(M, _lib.ptr(k), _lib.ptr(ka), _lib.ptr(v), _lib.ptr(va), b0, b1, _lib.ptr(ws),
                                   nbytes, byref(in_alt), _lib.stream_ptr(cuda_device))
    assert rc == 0
    torch.cuda.synchronize()
    rk, rv = (ka, va) if in_alt.value else (k, v)
    assert np.array_equal(rv.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(rk.cpu().numpy(), keys.numpy()[order])


@pytest.mark.parametrize("sem", [0, 1])
@pytest.mark.parametrize("W,H,ts", [(1920, 1080, 16), (1000, 700, 10), (3840, 2160, 16)])
def test_binning_mixed_huge_and_tiny_rectangles(cuda_device, W, H, ts, sem):
    """Rectangles from a single tile up to the whole image (thousands of tiles per Gaussian) in one warp of the
    pair-balanced emitter, odd widths (exact division by the rectangle width), zero radii, means far outside."""
    rng = np.random.default_rng(W + ts + sem)
    N = 3000
    means = np.stack([rng.uniform(-0.2 * W, 1.2 * W, N), rng.uniform(-0.2 * H, 1.2 * H, N)], 1).astype(np.float32)
    radii = rng.integers(0, 40, (N, 2)).astype(np.int32)
    huge = rng.choice(N, 60, replace=False)
    radii[huge] = rng.integers(200, 3 * max(W, H), (60, 2))
    radii[rng.choice(N, 300, replace=False)] = 0
    depths = rng.uniform(0.2, 50.0, N).astype(np.float32)
    depths[rng.choice(N, 200, replace=False)] = 7.0  # ties
    o_ids, o_ranges = oracle.bin_tiles(means, radii, depths, H, W, ts, semantics=sem)
    ids, ranges = binning.bin_gaussians_to_tiles_cuda(dev(means, cuda_device), dev(radii, cuda_device),
                                                      dev(depths, cuda_device), H, W, ts, semantics=sem)
    assert np.array_equal(ranges.cpu().numpy(), o_ranges)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert ids.numel() > 20 * N  # the huge ones dominate


@pytest.mark.parametrize("W,H,ts,N", [(1920, 1080, 4, 20_000),      # 480 x 270 = 129 600 tiles: 17 tile bits, 3 passes
                                       (3840, 2160, 4, 4_000),       # 960 x 540 = 518 400 tiles: 19 bits
                                       (4096, 4096, 1, 300)])        # 16.8 M tiles: 24 bits, 3 passes of 8
def test_binning_more_than_65536_tiles(cuda_device, W, H, ts, N):
    """Tile ids wider than two 8-bit digits (8K frames, small tiles): the tile sort takes ceil(bits / 8) passes.
    Both algorithms against the oracle, bit for bit."""
    g = torch.Generator().manual_seed(W + ts)
    m2 = torch.stack([torch.rand(N, generator=g) * (W + 40) - 20, torch.rand(N, generator=g) * (H + 40) - 20], 1)
    rad = torch.randint(0, 9 * ts, (N, 2), generator=g, dtype=torch.int32)
    dep = torch.rand(N, generator=g) * 20 + 0.1
    dep[::7] = dep[3]  # ties
    o_ids, o_ranges = oracle.bin_tiles(m2.numpy(), rad.numpy(), dep.numpy(), H, W, ts)
    assert o_ranges.shape[0] * o_ranges.shape[1] > 65536
    for algo in ("two_level", "single"):
        ids, ranges = binning.bin_gaussians_to_tiles_cuda(m2.to(cuda_device), rad.to(cuda_device), dep.to(cuda_device),
                                                          H, W, ts, algo=algo)
        assert np.array_equal(ranges.cpu().numpy(), o_ranges), algo
        assert np.array_equal(ids.cpu().numpy(), o_ids), algo


@pytest.mark.parametrize("cfg,N", [("config1_1k_256", None), ("config3_1m_1080p", 300_000)])
def test_packed_layout_identical_lists_and_image(cuda_device, cfg, N):
    """gsplat's packed (culled-compacted) layout (projection.mojo:73-87,213-244; tests/test_projection_mojo.py:238-247):
    Gaussians without a tile leave the frame before the depth sort.  Lists and image must not change."""
    sc = synthetic.make_scene(cfg, N=N)
    sc.opacities[::9] = 0.001   # opacity cull
    g = [t.to(cuda_device) for t in sc.gaussians()]
    cam = sc.camera.to(cuda_device)
    bg = sc.background.to(cuda_device)
    for sem in (_lib.SEM_TORCH, _lib.SEM_GSPLAT):
        a, aux_a = ms.render_fused(*g, cam, bg, semantics=sem, return_aux=True)
        a, aux_a = ms.render_fused(*g, cam, bg, semantics=sem, return_aux=True)  # (second call returns sorted_ids)
        b, aux_b = ms.render_fused(*g, cam, bg, semantics=sem, return_aux=True, packed=True)
        assert torch.equal(a, b)
        assert aux_a["n_isect"] == aux_b["n_isect"] and torch.equal(aux_a["tile_ranges"], aux_b["tile_ranges"])
        assert torch.equal(aux_a["sorted_ids"], aux_b["sorted_ids"])
        # the stage-level entry point too
        ids, ranges = binning.bin_gaussians_to_tiles_cuda(aux_a["means2d"], aux_a["radii"], aux_a["depths"], cam.H,
                                                          cam.W, 16, semantics=sem, packed=True)
        assert torch.equal(ids, aux_a["sorted_ids"]) and torch.equal(ranges, aux_a["tile_ranges"])
    culled = (aux_b["radii"] == 0).all(-1)
    assert culled[::9].all()      # the torch rules keep them in border tiles; the gsplat rules drop them
