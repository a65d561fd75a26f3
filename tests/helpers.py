"""Shared helpers for the parity tests (CUDA path vs oracle / golden fixtures)."""
import math

import numpy as np
import torch

import mojosplat_b200 as ms
from mojosplat_b200 import synthetic
from oracle import oracle


def camera_from_golden(g, device="cpu"):
    vm = torch.from_numpy(np.asarray(g["viewmat"], np.float32)).to(device)
    fx, fy, cx, cy = [float(v) for v in g["intr"]]
    W, H = [int(v) for v in g["size"]]
    near, far = [float(v) for v in g["clip"]]
    return ms.Camera(R=vm[:3, :3].contiguous(), T=vm[:3, 3].contiguous(), H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy,
                     near=near, far=far)


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def oracle_project_scene(sc, semantics=oracle.SEM_TORCH):
    cam = sc.camera
    return oracle.project(sc.means3d.numpy(), sc.log_scales.numpy(), sc.quats.numpy(), sc.opacities.numpy(),
                          cam.view_matrix.numpy(), cam.fx, cam.fy, cam.cx, cam.cy, cam.W, cam.H, cam.near,
                          cam.far, 0.3, semantics)


def scene_on(sc, device):
    return [t.to(device) for t in sc.gaussians()], sc.camera.to(device)


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def image_gate(img, ref, frac_allowed=1e-4, atol=1e-4, rtol=1e-4):
    """SURVEY H4 gate: >= (1 - frac_allowed) of the values within atol + rtol*|ref|, PSNR > 60 dB,
    and every outlier explained by a single alpha-threshold / saturation flip (bounded)."""
    img = np.asarray(img); ref = np.asarray(ref)
    err = np.abs(img - ref)
    bad = err > (atol + rtol * np.abs(ref))
    return dict(frac_bad=float(bad.mean()), max_err=float(err.max()), psnr=psnr(img, ref),
                ok=bool(bad.mean() <= frac_allowed and psnr(img, ref) > 60.0))
