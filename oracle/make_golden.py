"""Generate tests/golden/*.npz by running the UNMODIFIED reference torch backend on the CPU.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The fixtures pin oracle/oracle.c (and through it the CUDA kernels) to the reference's own
outputs for projection (mojosplat/projection.py:285-346) and binning
(mojosplat/binning.py:108-262).  The reference has no runnable rasterizer here, so no image
fixture comes from it (see oracle.c header: rasterizer parity is unpinned vs gsplat).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mojosplat_b200 import synthetic  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = ROOT / "tests" / "golden"


def ref_camera(ref, cam):
    return ref.utils.Camera(R=cam.R, T=cam.T, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy,
                            cx=cam.cx, cy=cam.cy, near=cam.near, far=cam.far)


def cam_dict(cam):
    return dict(viewmat=cam.view_matrix.numpy().astype(np.float32),
                intr=np.array([cam.fx, cam.fy, cam.cx, cam.cy], np.float32),
                size=np.array([cam.W, cam.H], np.int32),
                clip=np.array([cam.near, cam.far], np.float32))


def run_case(ref, name, means3d, log_scales, quats, opacities, colors, cam, tile_size=16):
    with torch.no_grad():
        m2, con, dep, rad = ref.projection.project_gaussians(
            means3d, log_scales, quats, opacities, ref_camera(ref, cam), backend="torch")
        ids, ranges = ref.binning.bin_gaussians_to_tiles(
            m2, rad, dep, cam.H, cam.W, tile_size, backend="torch")
    np.savez_compressed(
        OUT / f"{name}.npz",
        means3d=means3d.numpy(), log_scales=log_scales.numpy(), quats=quats.numpy(),
        opacities=opacities.numpy(), colors=colors.numpy(), tile_size=np.int32(tile_size),
        means2d=m2.numpy(), conics=con.numpy(), depths=dep.numpy(), radii=rad.numpy(),
        sorted_ids=ids.numpy().astype(np.int32), tile_ranges=ranges.numpy().astype(np.int32),
        **cam_dict(cam))
    print(f"{name}: N={means3d.shape[0]} M={ids.numel()} visible={(rad > 0).all(-1).sum().item()}")


def run_binning_case(ref, name, means2d, radii, depths, H, W, tile_size):
    with torch.no_grad():
        ids, ranges = ref.binning.bin_gaussians_to_tiles(means2d, radii, depths, H, W, tile_size,
                                                         backend="torch")
    np.savez_compressed(
        OUT / f"{name}.npz", means2d=means2d.numpy(), radii=radii.numpy(), depths=depths.numpy(),
        size=np.array([W, H], np.int32), tile_size=np.int32(tile_size),
        sorted_ids=ids.numpy().astype(np.int32), tile_ranges=ranges.numpy().astype(np.int32))
    print(f"{name}: N={means2d.shape[0]} M={ids.numel()}")


def test_style_gaussians(N, seed):
    """tests/test_projection_mojo.py:36-50 distribution, CPU generator."""
    g = torch.Generator().manual_seed(seed)
    means3d = torch.randn(N, 3, generator=g) * 2.0
    means3d[:, 2] = means3d[:, 2].abs() + 1.0
    scales = torch.log(torch.rand(N, 3, generator=g) * 0.3 + 0.05)
    quats = torch.nn.functional.normalize(torch.randn(N, 4, generator=g), p=2, dim=-1)
    opac = torch.sigmoid(torch.randn(N, generator=g))
    colors = torch.rand(N, 3, generator=g)
    return means3d, scales, quats, opac, colors


def main():
    ref = ref_import.load()
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)

    # BASELINE.json configs[0]: render_sample-style 1 000 Gaussians at 256x256
    sc = synthetic.make_scene("config1_1k_256")
    run_case(ref, "config1_1k_256", *sc.gaussians(), sc.camera)

    # config-3 ("garden") distribution at 1080p, subsampled so the reference's Python loop is quick
    sc = synthetic.make_scene("config3_1m_1080p", N=6000, seed=7)
    run_case(ref, "garden_6k_1080p", *sc.gaussians(), sc.camera)

    # config-2 distribution (large splats, many tiles per Gaussian), small N
    sc = synthetic.make_scene("config2_100k_1080p", N=300, seed=11)
    run_case(ref, "dense_300_1080p", *sc.gaussians(), sc.camera)

    # reference test distribution, identity and offset cameras (tests/test_projection_mojo.py:16-31)
    for cname, T in (("identity", (0.0, 0.0, 0.0)), ("offset", (0.0, 0.0, 5.0))):
        cam = synthetic.Camera(R=torch.eye(3), T=torch.tensor(T), H=64, W=64, fx=100.0, fy=100.0,
                               cx=32.0, cy=32.0, near=0.1, far=100.0)
        run_case(ref, f"teststyle_500_{cname}", *test_style_gaussians(500, 42), cam)

    # non-square image, non-power-of-two tile size, non-centred principal point
    cam = synthetic.make_camera(200, 120, 150.0)
    cam = synthetic.Camera(R=cam.R, T=cam.T, H=120, W=200, fx=150.0, fy=140.0, cx=90.5, cy=70.25)
    m, s, q, o, c = synthetic.make_gaussians(800, seed=3, log_scale_mean=-2.5)
    run_case(ref, "odd_800_200x120_ts10", m, s, q, o, c, cam, tile_size=10)

    # binning-only cases with float radii (tests/test_binning.py:18-75 fixtures)
    run_binning_case(ref, "bin_simple", torch.tensor([[32.0, 32.0]]), torch.tensor([[8.0, 8.0]]),
                     torch.tensor([1.0]), 64, 64, 16)
    edge_m = torch.tensor([[0.0, 0.0], [63.0, 63.0], [-10.0, 32.0], [74.0, 32.0], [32.0, -10.0],
                           [32.0, 74.0]])
    run_binning_case(ref, "bin_edge", edge_m, torch.full((6, 2), 5.0),
                     torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]), 64, 64, 16)
    g = torch.Generator().manual_seed(5)
    m2 = torch.rand(50, 2, generator=g) * 256
    rr = torch.rand(50, 2, generator=g) * 20 + 5
    dd = torch.rand(50, generator=g) * 10 + 0.5
    for ts in (8, 16, 32):
        run_binning_case(ref, f"bin_50_ts{ts}", m2, rr, dd, 256, 256, ts)
    # negative / zero / huge depths and repeated depths (tie order)
    g = torch.Generator().manual_seed(9)
    m2 = torch.rand(400, 2, generator=g) * torch.tensor([320.0, 200.0]) - 10.0
    rr = (torch.rand(400, 2, generator=g) * 30).floor()
    dd = torch.randn(400, generator=g) * 5.0
    dd[::7] = 2.5
    dd[3::11] = -1.0
    dd[5] = 0.0
    dd[6] = 1e20
    run_binning_case(ref, "bin_ties_400", m2, rr, dd, 200, 320, 16)
    run_binning_case(ref, "bin_empty", torch.zeros(0, 2), torch.zeros(0, 2), torch.zeros(0), 64, 64, 16)


if __name__ == "__main__":
    main()
