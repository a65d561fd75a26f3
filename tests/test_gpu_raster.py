"""GPU parity: tile rasterizer (faithful and fast kernels) vs the CPU oracle on identical inputs.

Tolerance (north_star): 1e-4 max-abs at the reference's own test sizes (asserted with assert_allclose); at full
size the alpha-threshold / saturation discontinuities make single-ulp exp differences flip a handful
of pixels, so the gate is helpers.image_gate: >= 99.99 % of values within 1e-4 (+1e-4 rel), PSNR > 60 dB, max
error bounded by what one flipped decision can cause, and EVERY out-of-tolerance pixel reproduced by the oracle
with that decision forced the other way (oracle.raster_audit)."""
import numpy as np
import pytest
import torch

import mojosplat_b200 as ms
from conftest import load_golden
from helpers import camera_from_golden, dev, image_gate, oracle_project_scene
from mojosplat_b200 import rasterization, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu


def raster_inputs_from_golden(g, device):
    return [dev(g[k], device) for k in ("means2d", "conics", "colors", "opacities")], \
        dev(g["tile_ranges"], device), dev(g["sorted_ids"], device)


@pytest.mark.parametrize("name", ["config1_1k_256", "teststyle_500_offset", "teststyle_500_identity",
                                  "garden_6k_1080p", "dense_300_1080p"])
@pytest.mark.parametrize("mode", ["faithful", "fast", "fast_nocull"])
def test_raster_vs_oracle_golden_scenes(cuda_device, name, mode):
    g = load_golden(name)
    cam = camera_from_golden(g)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    ref = oracle.rasterize(g["means2d"], g["conics"], g["colors"], g["opacities"], bg, g["tile_ranges"],
                           g["sorted_ids"], cam.W, cam.H, 16)
    (m2, con, col, op), ranges, ids = raster_inputs_from_golden(g, cuda_device)
    img = rasterization.rasterize_gaussians_cuda(m2, con, col, op, dev(bg, cuda_device), ranges, ids, cam, 16,
                                                 mode=mode)
    assert img.shape == (cam.H, cam.W, 3) and img.dtype == torch.float32
    audit = dict(means2d=g["means2d"], conics=g["conics"], colors=g["colors"], opacities=g["opacities"],
                 background=bg, tile_ranges=g["tile_ranges"], sorted_ids=g["sorted_ids"], W=cam.W, H=cam.H)
    r = image_gate(img.cpu().numpy(), ref, audit=audit)
    assert r["ok"], r
    if mode == "faithful":
        assert r["frac_bad"] <= 2e-5, r  # same operation order: only exp-ulp flips remain


def make_test_style(N, seed):
    """tests/test_rasterization.py:24-36 distribution (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    means3d = torch.randn(N, 3, generator=g)
    means3d[:, 2] = torch.rand(N, generator=g) * 3.5 + 1.5
    log_scales = torch.ones(N, 3) * -2.0 + torch.randn(N, 3, generator=g) * 0.1
    quats = torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=1)
    opac = torch.rand(N, generator=g) * 0.45 + 0.5
    colors = torch.rand(N, 3, generator=g)
    return means3d, log_scales, quats, opac, colors


@pytest.mark.parametrize("N,HW,bgv", [(1, 64, 0.0), (5, 64, 0.0), (50, 64, 0.0), (200, 64, 0.0), (50, 64, 0.3),
                                      (100, 128, 0.0)])
def test_raster_reference_test_sizes_within_1e4(cuda_device, N, HW, bgv):
    """tests/test_rasterization.py:94-146 (mojo vs gsplat at atol=rtol=1e-4) with the oracle as the other side."""
    m, s, q, o, c = make_test_style(N, seed=N)
    cam = ms.Camera(R=torch.eye(3), T=torch.zeros(3), H=HW, W=HW, fx=100.0, fy=100.0, cx=HW / 2, cy=HW / 2)
    m2, con, dep, rad = oracle.project(m.numpy(), s.numpy(), q.numpy(), o.numpy(), cam.view_matrix.numpy(),
                                       cam.fx, cam.fy, cam.cx, cam.cy, HW, HW)
    ids, ranges = oracle.bin_tiles(m2, rad, dep, HW, HW, 16)
    bg = np.full(3, bgv, np.float32)
    ref = oracle.rasterize(m2, con, c.numpy(), o.numpy(), bg, ranges, ids, HW, HW, 16)
    for mode in ("fast", "faithful"):
        img = rasterization.rasterize_gaussians_cuda(dev(m2, cuda_device), dev(con, cuda_device), c.to(cuda_device),
                                                     o.to(cuda_device), dev(bg, cuda_device), dev(ranges, cuda_device),
                                                     dev(ids, cuda_device), cam, 16, mode=mode).cpu().numpy()
        np.testing.assert_allclose(img, ref, atol=1e-4, rtol=1e-4)


def test_raster_known_answers(cuda_device):
    """tests/test_rasterization.py:154-266."""
    cam = ms.Camera(R=torch.eye(3), T=torch.zeros(3), H=64, W=64, fx=100.0, fy=100.0, cx=32.0, cy=32.0)
    d = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt, device=cuda_device)
    m2, con = d([[32.0, 32.0]]), d([[0.05, 0.0, 0.05]])
    col, op = d([[1.0, 0.0, 0.0]]), d([0.9])
    bg = d([0.2, 0.4, 0.6])
    empty = torch.zeros((4, 4, 2), dtype=torch.int32, device=cuda_device)
    img = ms.rasterize_gaussians(m2, con, col, op, bg, empty, torch.zeros(0, dtype=torch.int32, device=cuda_device), cam)
    assert (img - bg).abs().max().item() <= 1e-6  # empty ranges => exact background
    ids, ranges = ms.bin_gaussians_to_tiles(m2, d([[30, 30]], torch.int32), d([2.0]), 64, 64, 16)
    img = ms.rasterize_gaussians(m2, con, col, op, torch.zeros(3, device=cuda_device), ranges.long(), ids, cam)
    assert img[32, 32, 0] > 0.1 and img[32, 32, 1] == 0 and torch.isfinite(img).all()
    vals = [ms.rasterize_gaussians(m2, con, col, d([o]), torch.zeros(3, device=cuda_device), ranges, ids, cam)[32, 32, 0].item()
            for o in (0.2, 0.5, 0.9)]
    assert vals[0] < vals[1] < vals[2]
    m2b, conb = d([[32.0, 32.0], [32.0, 32.0]]), d([[0.05, 0, 0.05]] * 2)
    ids, ranges = ms.bin_gaussians_to_tiles(m2b, d([[30, 30]] * 2, torch.int32), d([1.0, 3.0]), 64, 64, 16)
    img = ms.rasterize_gaussians(m2b, conb, d([[1.0, 0, 0], [0, 1.0, 0]]), d([[0.9], [0.9]]), torch.zeros(3, device=cuda_device),
                                 ranges, ids, cam)
    assert img[32, 32, 0] > img[32, 32, 1] > 0
    with pytest.raises(ValueError, match="Invalid backend"):
        ms.rasterize_gaussians(m2, con, col, op, bg, empty, ids, cam, backend="nope")


@pytest.mark.parametrize("ts", [8, 10, 32])
@pytest.mark.parametrize("C", [1, 3, 5])
def test_raster_other_tile_sizes_and_channels(cuda_device, ts, C):
    sc = synthetic.make_scene("config1_1k_256", N=400, seed=3)
    cam = sc.camera
    m2, con, dep, rad = oracle_project_scene(sc)
    ids, ranges = oracle.bin_tiles(m2, rad, dep, cam.H, cam.W, ts)
    g = torch.Generator().manual_seed(C)
    col = torch.rand(sc.N, C, generator=g)
    bg = np.linspace(0.1, 0.5, C).astype(np.float32)
    ref = oracle.rasterize(m2, con, col.numpy(), sc.opacities.numpy(), bg, ranges, ids, cam.W, cam.H, ts)
    img = ms.rasterize_gaussians(dev(m2, cuda_device), dev(con, cuda_device), col.to(cuda_device),
                                 sc.opacities.to(cuda_device), dev(bg, cuda_device), dev(ranges, cuda_device),
                                 dev(ids, cuda_device), cam, tile_size=ts)
    r = image_gate(img.cpu().numpy(), ref)
    assert r["ok"], r


@pytest.mark.parametrize("cfg,N", [("config3_1m_1080p", 1_000_000), ("config2_100k_1080p", 100_000)])
def test_raster_full_size(cuda_device, cfg, N):
    """Full BASELINE sizes: fast kernel vs oracle (gate), culling is exact (bit-identical to no-cull),
    and the GPU work counters equal the oracle's."""
    sc = synthetic.make_scene(cfg, N=N)
    cam = sc.camera
    m2, con, dep, rad = oracle_project_scene(sc)
    ids, ranges = oracle.bin_tiles(m2, rad, dep, cam.H, cam.W, 16)
    bg = np.full(3, 0.1, np.float32)
    ref, (e_all, e_pass) = oracle.rasterize(m2, con, sc.colors.numpy(), sc.opacities.numpy(), bg, ranges, ids,
                                            cam.W, cam.H, 16, return_stats=True)
    a = [dev(m2, cuda_device), dev(con, cuda_device), sc.colors.to(cuda_device), sc.opacities.to(cuda_device),
         dev(bg, cuda_device), dev(ranges, cuda_device), dev(ids, cuda_device), cam, 16]
    fast = rasterization.rasterize_gaussians_cuda(*a, mode="fast")
    nocull = rasterization.rasterize_gaussians_cuda(*a, mode="fast_nocull")
    assert torch.equal(fast, nocull)
    audit = dict(means2d=m2, conics=con, colors=sc.colors.numpy(), opacities=sc.opacities.numpy(), background=bg,
                 tile_ranges=ranges, sorted_ids=ids, W=cam.W, H=cam.H)
    r = image_gate(fast.cpu().numpy(), ref, audit=audit)
    assert r["ok"], r
    print(f"\n[parity {cfg}] fast vs oracle: {r}")
    img_f, g_all, g_pass = rasterization.rasterize_gaussians_stats(*a)
    rf = image_gate(img_f.cpu().numpy(), ref, frac_allowed=2e-5, audit=audit)
    assert rf["ok"], rf
    assert abs(g_all - e_all) <= 1e-5 * e_all and abs(g_pass - e_pass) <= 1e-5 * e_pass


@pytest.mark.parametrize("cfg,N,sem", [("config5_6m_4k", 1_500_000, 0), ("config3_1m_1080p", 1_000_000, 0)])
def test_long_list_prepass_is_exact(cuda_device, cfg, N, sem):
    """Fused frames compact very long tile lists (the border tiles the torch binning rules fill with culled Gaussians)
    in a multi-SM pre-pass and rasterize a private copy of the list; the stage-level call walks the full list with the
    in-kernel tile test.  Same image, bit for bit, and sorted_ids / tile_ranges are the full lists."""
    sc = synthetic.make_scene(cfg, N=N)
    sc.means3d[::3] *= 6.0   # push a third of the scene far outside the frustum: border lists of > 100 k entries
    g = [t.to(cuda_device) for t in sc.gaussians()]
    cam, bg = sc.camera, sc.background.to(cuda_device)
    ms.render_fused(*g, cam, bg, semantics=sem, return_aux=True)
    img, aux = ms.render_fused(*g, cam, bg, semantics=sem, return_aux=True)
    ranges = aux["tile_ranges"].reshape(-1, 2)
    longest = int((ranges[:, 1] - ranges[:, 0]).max())
    assert longest > 16384, longest
    ref = rasterization.rasterize_gaussians_cuda(aux["means2d"], aux["conics"], g[4], g[3], bg, aux["tile_ranges"],
                                                 aux["sorted_ids"], cam, 16, mode="fast")
    assert torch.equal(img, ref)
    nocull = rasterization.rasterize_gaussians_cuda(aux["means2d"], aux["conics"], g[4], g[3], bg, aux["tile_ranges"],
                                                    aux["sorted_ids"], cam, 16, mode="fast_nocull")
    assert torch.equal(img, nocull)
    assert int(ranges[-1, 1]) == aux["n_isect"] == aux["sorted_ids"].numel()
