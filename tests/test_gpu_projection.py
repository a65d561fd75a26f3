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


@pytest.mark.parametrize("cfg,N", [("config3_1m_1080p", 1_000_000), ("config2_100k_1080p", 100_000)])
def test_projection_full_size_vs_oracle(cuda_device, cfg, N):
    sc = synthetic.make_scene(cfg, N=N)
    ref = oracle_project_scene(sc)
    (m, s, q, o, c), cam = scene_on(sc, cuda_device)
    out = ms.project_gaussians(m, s, q, o, cam, backend="b200")
    check_projection(out, ref, N, max(2, N // 100_000))


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
