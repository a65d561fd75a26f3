"""GPU parity: projection kernel (bsplat_project_fwd) vs the reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

import mojosplat_b200 as ms
from conftest import load_golden
from helpers import camera_from_golden, dev, oracle_project_scene, scene_on
from mojosplat_b200 import synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu
FULL_CASES = ["config1_1k_256", "garden_6k_1080p", "dense_300_1080p", "teststyle_500_identity",
              "teststyle_500_offset", "odd_800_200x120_ts10"]


def check_projection(out, ref, N, radii_slack):
    m2, con, dep, rad = [t.cpu().numpy() for t in out]
    rm2, rcon, rdep, rrad = ref
    assert m2.dtype == np.float32 and con.dtype == np.float32 and dep.dtype == np.float32
    assert rad.dtype == np.int32 and m2.shape == (N, 2) and con.shape == (N, 3) and dep.shape == (N,)
    # tolerance of the reference's own tests (test_rasterization.py:110): atol + rtol*|ref|
    np.testing.assert_allclose(m2, rm2, atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(con, rcon, atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(dep, rdep, atol=1e-4, rtol=1e-4)
    d = np.abs(rad.astype(np.int64) - rrad.astype(np.int64))
    assert d.max() <= 1
    assert (d > 0).sum() <= radii_slack, (d > 0).sum()
    assert ((rad > 0).all(-1) != (rrad > 0).all(-1)).sum() <= radii_slack


@pytest.mark.parametrize("name", FULL_CASES)
def test_projection_vs_reference_golden(cuda_device, name):
    g = load_golden(name)
    cam = camera_from_golden(g, cuda_device)
    out = ms.project_gaussians(dev(g["means3d"], cuda_device), dev(g["log_scales"], cuda_device),
                               dev(g["quats"], cuda_device), dev(g["opacities"], cuda_device), cam, backend="cuda")
    N = g["means3d"].shape[0]
    check_projection(out, (g["means2d"], g["conics"], g["depths"], g["radii"]), N, max(1, N // 2000))
    # the kernel rounds like the reference's torch ops (projection.cu header): means2d, depths and radii of the
    # unmodified reference bit for bit; conics wherever MKL's exp gave the correctly rounded scale (~97 % of the rows)
    m2, con, dep, rad = [t.cpu().numpy() for t in out]
    eq = lambda a, b: (a == b) | ((a != a) & (b != b))
    assert eq(m2, g["means2d"]).all() and eq(dep, g["depths"]).all()
    assert np.array_equal(rad, g["radii"])
    assert eq(con, g["conics"]).all(-1).mean() >= 0.95


@pytest.mark.parametrize("cfg,N", [("config3_1m_1080p", 1_000_000), ("config2_100k_1080p", 100_000)])
def test_projection_full_size_vs_oracle(cuda_device, cfg, N):
    sc = synthetic.make_scene(cfg, N=N)
    ref = oracle_project_scene(sc)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    out = ms.project_gaussians(m, s, q, o, cam, backend="b200")
    check_projection(out, ref, N, max(2, N // 100_000))
    # means2d and depths involve only products, sums, FMAs and IEEE divisions: the kernel's reciprocal-based
    # quotients (one correctly rounded reciprocal per denominator + an exact-remainder correction) must reproduce the
    # oracle's `/` bit for bit; radii and conics also depend on exp (libdevice vs glibc double exp, both < 1 ulp in
    # double: the float results agree except for a handful of arguments per million)
    assert np.array_equal(out[0].cpu().numpy(), ref[0])
    assert np.array_equal(out[2].cpu().numpy(), ref[2])
    assert (out[3].cpu().numpy() != ref[3]).any(-1).sum() <= 2
    con_rows = (out[1].cpu().numpy() == ref[1]).all(-1) | ~np.isfinite(ref[1]).all(-1)
    assert con_rows.mean() >= 0.9999, con_rows.mean()
    # the A/B build with FMA contraction: the VISIBLE Gaussians stay within the tolerances (radii may flip at integer
    # crossings a little more often); culled ones with catastrophic cancellation in (fx x + cx z) / z (|z| tiny, pixel
    # coordinates ~1e7) move by up to ~0.5 % -- one reason the default build rounds like the oracle
    from mojosplat_b200.projection import project_gaussians_cuda
    out_fma = [t.cpu().numpy() for t in project_gaussians_cuda(m, s, q, o, cam, allow_fma=True)]
    vis = (ref[3] > 0).all(-1)
    np.testing.assert_allclose(out_fma[0][vis], ref[0][vis], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(out_fma[1][vis], ref[1][vis], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(out_fma[2], ref[2], atol=1e-4, rtol=1e-4)
    d_exact = int((out[3].cpu().numpy() != ref[3]).any(-1).sum())
    d_fma = int((out_fma[3] != ref[3]).any(-1).sum())
    assert np.abs(out_fma[3].astype(np.int64) - ref[3]).max() <= 1 and d_fma <= max(8, N // 20_000)
    bad_culled = int((np.abs(out_fma[0] - ref[0]) > 1e-4 + 1e-4 * np.abs(ref[0])).any(-1).sum())
    print(f"\n[projection {cfg}] radii that differ from the oracle: exact build {d_exact}, FMA build {d_fma} of {N}; "
          f"means2d rows of the FMA build outside 1e-4 + 1e-4|ref| (all culled): {bad_culled}")
    # the fast-math build (BSPLAT_PROJ_FAST_MATH: MUFU exp / rcp / sqrt): an OPTION for callers that want the HBM-bound
    # kernel.  Visible Gaussians within the float tolerances; radii at most one off, and only next to integer crossings
    out_fast = [t.cpu().numpy() for t in project_gaussians_cuda(m, s, q, o, cam, fast_math=True)]
    np.testing.assert_allclose(out_fast[0][vis], ref[0][vis], atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(out_fast[1][vis], ref[1][vis], atol=1e-4, rtol=2e-4)
    np.testing.assert_allclose(out_fast[2], ref[2], atol=1e-4, rtol=1e-4)
    d_fast = int((out_fast[3] != ref[3]).any(-1).sum())
    assert np.abs(out_fast[3].astype(np.int64) - ref[3]).max() <= 1 and d_fast <= max(16, N // 2_000), d_fast
    print(f"[projection {cfg}] fast-math build: {d_fast} of {N} radii one off")


def test_projection_gsplat_semantics(cuda_device):
    sc = synthetic.make_scene("config1_1k_256")
    sc.opacities[::9] = 0.001  # opacity cull (test_projection_mojo.py:238-247)
    ref = oracle_project_scene(sc, oracle.SEM_GSPLAT)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    out = ms.project_gaussians(m, s, q, o.unsqueeze(-1), cam, backend="cuda_gsplat")
    check_projection(out, ref, sc.N, 2)
    rad = out[3].cpu().numpy()
    assert (rad[::9] == 0).all()
    culled = (rad == 0).all(-1)
    assert (out[0].cpu().numpy()[culled] == 0).all()  # culled rows are zeroed in this mode


def test_projection_geometry_known_answers(cuda_device):
    """tests/test_projection_mojo.py:203-258: on-axis -> centre, depth == z, behind camera -> culled."""
    cam = ms.Camera(R=torch.eye(3, device=cuda_device), T=torch.zeros(3, device=cuda_device), H=64, W=64,
                    fx=100.0, fy=100.0, cx=32.0, cy=32.0)
    m = torch.tensor([[0.0, 0.0, 2.0], [0.0, 0.0, -2.0], [0.3, -0.2, 4.0]], device=cuda_device)
    s = torch.log(torch.full((3, 3), 0.1, device=cuda_device))
    q = torch.tensor([[1.0, 0, 0, 0]] * 3, device=cuda_device)
    m2, con, dep, rad = ms.project_gaussians(m, s, q, torch.ones(3, 1, device=cuda_device), cam)
    assert abs(m2[0, 0].item() - 32) < 1e-3 and abs(m2[0, 1].item() - 32) < 1e-3
    assert abs(dep[0].item() - 2.0) < 1e-6 and abs(dep[2].item() - 4.0) < 1e-6
    assert (rad[1] == 0).all() and (rad[0] > 0).all()


def test_projection_extreme_depths_take_the_plain_division_path(cuda_device):
    """Denominators outside [2^-30, 2^30] (and zero) leave the reciprocal-based quotients: same results as the oracle."""
    sc = synthetic.make_scene("config1_1k_256", N=64)
    sc.means3d[:16] *= 1e-12     # z ~ camera offset only
    sc.means3d[16:32, 2] = 1e12  # far beyond 2^30
    sc.quats[32:40] *= 1e-20     # norms below the 1e-12 clamp (denormal squares)
    cam = sc.camera
    sc.means3d[40:44] = torch.tensor([0.0, 1.5, 5.0]) @ torch.eye(3)  # exactly the eye position: z = 0 in camera space
    ref = oracle_project_scene(sc)
    (m, s, q, o, c), camd = scene_on(sc, cuda_device)
    out = [t.cpu().numpy() for t in ms.project_gaussians(m, s, q, o, camd)]
    for a, b in zip(out[:3], ref[:3]):
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        np.testing.assert_allclose(a[fin], b[fin], atol=1e-4, rtol=1e-4)
    assert np.array_equal(out[3], ref[3])


def test_projection_empty_and_unaligned(cuda_device):
    cam = synthetic.make_camera(64, 64, 50.0).to(cuda_device)
    z = lambda *s: torch.zeros(*s, device=cuda_device)
    out = ms.project_gaussians(z(0, 3), z(0, 3), z(0, 4), z(0), cam)
    assert [tuple(t.shape) for t in out] == [(0, 2), (0, 3), (0,), (0, 2)]
    # N not a multiple of the block, tensors that are views at odd offsets
    sc = synthetic.make_scene("config1_1k_256", N=777)
    ref = oracle_project_scene(sc)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    out = ms.project_gaussians(m, s, q, o, cam)
    check_projection(out, ref, 777, 1)
