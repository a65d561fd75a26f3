"""GPU: spherical-harmonics colours (bsplat_sh_eval) against the float64 numpy restatement."""
import numpy as np
import pytest
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import synthetic
from mojosplat_b200.sh import camera_position, eval_sh
from oracle.oracle_np import sh_eval_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree,K", [(0, 1), (0, 16), (1, 4), (2, 9), (3, 16), (2, 16), (1, 9)])
@pytest.mark.parametrize("N", [1, 777, 100_003])
def test_sh_matches_numpy(cuda_device, degree, K, N):
    g = torch.Generator().manual_seed(degree * 100 + K)
    coeffs = torch.randn(N, K, 3, generator=g) * 0.6
    means = torch.randn(N, 3, generator=g) * 2
    cam = synthetic.make_camera(640, 360, 300.0)
    pos = camera_position(cam).numpy()
    full = np.zeros((N, 16, 3)); full[:, :K] = coeffs.numpy()
    ref = sh_eval_np(degree, full, means.numpy(), pos)
    out = eval_sh(degree, coeffs.to(cuda_device), means.to(cuda_device), cam).cpu().numpy()
    assert out.shape == (N, 3) and out.dtype == np.float32
    np.testing.assert_allclose(out, ref, atol=2e-5, rtol=1e-5)
    assert (out >= 0).all() and (out == 0).any() == (ref <= 0).any()


def test_render_with_sh_coefficients(cuda_device):
    """render_gaussians(features=(N,K,3), sh_degree=d) == render of the evaluated colours; 2-D features keep the
    reference's placeholder behaviour (warning, first three channels)."""
    sc = synthetic.make_scene("config1_1k_256")
    m, s, q, o, c = [t.to(cuda_device) for t in sc.gaussians()]
    gen = torch.Generator().manual_seed(5)
    coeffs = (torch.randn(sc.N, 16, 3, generator=gen) * 0.4).to(cuda_device)
    cam = sc.camera
    img = ms.render_gaussians(m, s, q, o, coeffs, cam, sh_degree=3, background_color=sc.background.to(cuda_device))
    cols = eval_sh(3, coeffs, m, cam)
    ref = ms.render_gaussians(m, s, q, o, cols, cam, background_color=sc.background.to(cuda_device))
    assert torch.equal(img, ref)
    img0 = ms.render_gaussians(m, s, q, o, coeffs, cam, sh_degree=0, background_color=sc.background.to(cuda_device))
    assert not torch.equal(img0, img)
    with pytest.warns(UserWarning):
        feat = torch.cat([c, c], dim=1)
        ph = ms.render_gaussians(m, s, q, o, feat, cam, sh_degree=3,
                                 background_color=torch.cat([sc.background, sc.background]).to(cuda_device))
    assert torch.equal(ph, ms.render_gaussians(m, s, q, o, c, cam, background_color=sc.background.to(cuda_device)))
    with pytest.raises(ValueError):
        eval_sh(3, coeffs[:, :9], m, cam)


def test_sh_bandwidth_1m(cuda_device):
    """Degree 3 at 1 M Gaussians: 216 B per Gaussian of algorithmic traffic; report GB/s (no assertion on speed
    beyond 'not pathological')."""
    N = 1_000_000
    coeffs = torch.randn(N, 16, 3, device=cuda_device)
    means = torch.randn(N, 3, device=cuda_device)
    cam = synthetic.make_camera(1920, 1080, 1000.0)
    eval_sh(3, coeffs, means, cam)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        eval_sh(3, coeffs, means, cam)
    b.record(); torch.cuda.synchronize()
    ms_per = a.elapsed_time(b) / 10
    gbs = 216 * N / (ms_per * 1e-3) / 1e9
    print(f"sh_eval degree 3, 1M: {ms_per:.4f} ms, {gbs:.0f} GB/s")
    assert gbs > 500
