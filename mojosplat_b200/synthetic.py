"""Synthetic scenes of BASELINE.json's configs.

Everything is drawn on the CPU with ``torch.Generator().manual_seed(seed)`` so that the CPU
CPU checker and the CUDA kernels see bit-identical inputs (the reference seeds the CUDA generator,
tests/test_rasterization.py:26-27, which is not reproducible without a GPU).  Distribution
and draw order follow the reference's render_sample.py:86-102.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .utils import Camera


def look_at(eye: torch.Tensor, target: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """World->camera 4x4, gsplat convention +X right / +Y down / +Z forward
    (same construction as the reference's render_sample.py:12-30)."""
    eye, target, up = eye.float(), target.float(), up.float()
    forward = torch.nn.functional.normalize(target - eye, dim=0)
    right = torch.nn.functional.normalize(torch.linalg.cross(forward, up), dim=0)
    down = torch.linalg.cross(right, forward)
    Rt = torch.stack([right, down, forward], dim=0)
    vm = torch.eye(4, dtype=torch.float32)
    vm[:3, :3] = Rt
    vm[:3, 3] = -(Rt @ eye)
    return vm


def make_camera(W: int, H: int, focal: float, eye=(0.0, 1.5, 5.0), target=(0.0, 0.0, 0.0),
                up=(0.0, 1.0, 0.0), near: float = 0.1, far: float = 100.0) -> Camera:
    vm = look_at(torch.tensor(eye), torch.tensor(target), torch.tensor(up))
    return Camera(R=vm[:3, :3].contiguous(), T=vm[:3, 3].contiguous(), H=H, W=W,
                  fx=float(focal), fy=float(focal), cx=W / 2.0, cy=H / 2.0, near=near, far=far)


def orbit_cameras(n_views: int, W: int, H: int, focal: float, radius: float = 5.0,
                  height: float = 1.5) -> list:
    """Config 4: poses on a circle, eye_k = (r sin t_k, h, r cos t_k), looking at the origin."""
    cams = []
    for k in range(n_views):
        t = 2.0 * math.pi * k / n_views
        cams.append(make_camera(W, H, focal, eye=(radius * math.sin(t), height, radius * math.cos(t))))
    return cams


def make_gaussians(N: int, seed: int = 42, log_scale_mean: float = -2.0,
                   log_scale_std: float = 0.3, spread: float = 2.0, channels: int = 3):
    """render_sample.py:86-102 distribution, drawn in the same order, on the CPU."""
    g = torch.Generator().manual_seed(seed)
    means3d = torch.randn(N, 3, generator=g) * spread
    log_scales = torch.ones(N, 3) * log_scale_mean + torch.randn(N, 3, generator=g) * log_scale_std
    quats = torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=1)
    opacities = torch.sigmoid(torch.randn(N, generator=g) + 1.0)
    colors = torch.rand(N, channels, generator=g)
    return (means3d.float().contiguous(), log_scales.float().contiguous(),
            quats.float().contiguous(), opacities.float().contiguous(),
            colors.float().contiguous())


@dataclass
class Scene:
    name: str
    means3d: torch.Tensor
    log_scales: torch.Tensor
    quats: torch.Tensor
    opacities: torch.Tensor
    colors: torch.Tensor
    camera: Camera
    background: torch.Tensor

    @property
    def N(self) -> int:
        return self.means3d.shape[0]

    def gaussians(self):
        return self.means3d, self.log_scales, self.quats, self.opacities, self.colors


# name -> (N, W, H, focal, log_scale_mean, log_scale_std)
CONFIGS = {
    "config1_1k_256": (1_000, 256, 256, 500.0 * 256 / 1920, -2.0, 0.3),
    "config2_100k_1080p": (100_000, 1920, 1080, 500.0, -2.0, 0.3),
    "config3_1m_1080p": (1_000_000, 1920, 1080, 1000.0, -4.5, 0.5),
    "config3_dense_1m_1080p": (1_000_000, 1920, 1080, 500.0, -2.0, 0.3),
    "config4_3m_1080p": (3_000_000, 1920, 1080, 1000.0, -4.5, 0.5),
    "config5_6m_4k": (6_000_000, 3840, 2160, 2000.0, -4.5, 0.5),
}


def make_scene(name: str, N: int | None = None, seed: int = 42) -> Scene:
    n, W, H, focal, lsm, lss = CONFIGS[name]
    n = n if N is None else N
    m, s, q, o, c = make_gaussians(n, seed=seed, log_scale_mean=lsm, log_scale_std=lss)
    cam = make_camera(W, H, focal)
    return Scene(name, m, s, q, o, c, cam, torch.full((3,), 0.1))
