"""How the reference's torch ops round -- the evidence behind oracle.c's arithmetic.  BUILD CONTAINER ONLY
(imports the unmodified reference from /root/reference through oracle/ref_import.py).  TEST INFRASTRUCTURE ONLY.

For every step of mojosplat/projection.py:51-283 the reference function is called on a golden fixture's inputs and
its result compared, bit for bit, with candidate restatements (numpy float32; fma(a, b, c) emulated through float64,
exact for float32 operands up to a 2^-29 double-rounding chance):

    python oracle/bisect_rounding.py [fixture]

Findings (torch 2.11.0 CPU, MKL 2024.2; every fixture of tests/golden agrees):
    F.normalize                      sum of squares left to right, sqrt, divide                      100 %
    quat -> R                        one rounding per operation                                      100 %
    M M^T   ("ij,kj->ik")            product + sum, every operation rounded                          100 %  (FMA chain 10 %)
    R mu    ("cij,nj->cni")          c = a0 b0; c = fma(a1, b1, c); c = fma(a2, b2, c)               100 %  (separate 90 %)
    (R S) R^T ("cij,njk,clk->cnil")  left to right, both products FMA chains                         100 %  (separate 9 %)
    (J S) J^T ("ij,jk,kl->il")       left to right, both products product + sum                      100 %  (FMA chains 44 %)
    K mu_c  ("ij,nj->ni")            FMA chain                                                       100 %
    torch.exp                        MKL VML: == correctly rounded exp for 98.9 % of 4 M arguments (glibc expf 60.7 %,
                                     SLEEF u10 90.4 %) -- not reproducible without MKL
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_golden  # noqa: E402
from helpers import camera_from_golden  # noqa: E402
from oracle import ref_import  # noqa: E402

f32 = np.float32


def fma(a, b, c):
    return (np.asarray(a, f32).astype(np.float64) * np.asarray(b, f32).astype(np.float64)
            + np.asarray(c, f32).astype(np.float64)).astype(f32)


def mm(A, B, fused):
    """C[n,i,k] = sum_j A[n,i,j] B[n,j,k], j ascending; fused: FMA chain, else product + sum."""
    C = np.zeros((A.shape[0], A.shape[1], B.shape[2]), f32)
    for i in range(A.shape[1]):
        for k in range(B.shape[2]):
            c = (A[:, i, 0] * B[:, 0, k]).astype(f32)
            for j in range(1, A.shape[2]):
                c = fma(A[:, i, j], B[:, j, k], c) if fused else (c + (A[:, i, j] * B[:, j, k]).astype(f32)).astype(f32)
            C[:, i, k] = c
    return C


def frac(a, b):
    return round(float(((a == b) | ((a != a) & (b != b))).reshape(a.shape[0], -1).all(1).mean()), 4)


def main(name="garden_6k_1080p"):
    P = ref_import.load().projection
    g = load_golden(name)
    cam = camera_from_golden(g)
    q, ls, mu = torch.from_numpy(g["quats"]), torch.from_numpy(g["log_scales"]), torch.from_numpy(g["means3d"])
    N = q.shape[0]
    s_t = torch.exp(ls).numpy()
    print("exp: glibc", frac(np.exp(g["log_scales"]).astype(f32), s_t), " correctly rounded",
          frac(np.exp(g["log_scales"].astype(np.float64)).astype(f32), s_t))
    qn_t = torch.nn.functional.normalize(q, p=2, dim=-1).numpy()
    qq = g["quats"]
    n = np.sqrt(((qq[:, 0] * qq[:, 0] + qq[:, 1] * qq[:, 1]) + qq[:, 2] * qq[:, 2]) + qq[:, 3] * qq[:, 3]).astype(f32)
    print("normalize (sequential sum):", frac((qq / n[:, None]).astype(f32), qn_t))
    cov_t, _ = P._quat_scale_to_covar_preci(q, torch.exp(ls), True, False)
    cov_t = cov_t.numpy()
    M = P._quat_to_rotmat(q).numpy() * s_t[:, None, :]
    Mt = np.transpose(M, (0, 2, 1))
    print("M M^T: product+sum", frac(mm(M, Mt, False), cov_t), " FMA chain", frac(mm(M, Mt, True), cov_t))
    vm = cam.view_matrix
    mc_t, cc_t = P._world_to_cam(mu[None], torch.from_numpy(cov_t)[None], vm[None, None])
    mc_t, cc_t = mc_t[0, 0].numpy(), cc_t[0, 0].numpy()
    Rv = np.broadcast_to(vm[:3, :3].numpy().astype(f32), (N, 3, 3)).copy()
    RvT = np.transpose(Rv, (0, 2, 1))
    t = vm[:3, 3].numpy()
    for fused in (True, False):
        mc = mm(Rv, g["means3d"][:, :, None], fused)[:, :, 0] + t
        print(f"R mu ({'FMA chain' if fused else 'product+sum'}):", frac(mc, mc_t),
              " (R S) R^T:", frac(mm(mm(Rv, cov_t, fused), RvT, fused), cc_t),
              " R (S R^T):", frac(mm(Rv, mm(cov_t, RvT, fused), fused), cc_t))
    m2_t, c2_t = P._persp_proj(torch.from_numpy(mc_t)[None, None], torch.from_numpy(cc_t)[None, None],
                               cam.Ks[None, None], cam.W, cam.H)
    m2_t, c2_t = m2_t[0, 0].numpy(), c2_t[0, 0].numpy()
    fx, fy, cx, cy, W, H = [f32(v) for v in (cam.fx, cam.fy, cam.cx, cam.cy, cam.W, cam.H)]
    tz = mc_t[:, 2]
    tz2 = (tz * tz).astype(f32)
    tanx, tany = f32(0.5) * W / fx, f32(0.5) * H / fy
    lxp, lxn = (W - cx) / fx + f32(0.3) * tanx, cx / fx + f32(0.3) * tanx
    lyp, lyn = (H - cy) / fy + f32(0.3) * tany, cy / fy + f32(0.3) * tany
    with np.errstate(all="ignore"):
        tx = (tz * np.clip((mc_t[:, 0] / tz).astype(f32), -lxn, lxp)).astype(f32)
        ty = (tz * np.clip((mc_t[:, 1] / tz).astype(f32), -lyn, lyp)).astype(f32)
        J = np.zeros((N, 2, 3), f32)
        J[:, 0, 0] = fx / tz; J[:, 0, 2] = -fx * tx / tz2; J[:, 1, 1] = fy / tz; J[:, 1, 2] = -fy * ty / tz2
        Jt = np.transpose(J, (0, 2, 1))
        for f1 in (True, False):
            for f2 in (True, False):
                print(f"(J S) J^T first {'FMA' if f1 else 'sum'} second {'FMA' if f2 else 'sum'}:",
                      frac(mm(mm(J, cc_t, f1), Jt, f2), c2_t))
        K = np.zeros((N, 2, 3), f32)
        K[:, 0, 0] = fx; K[:, 0, 2] = cx; K[:, 1, 1] = fy; K[:, 1, 2] = cy
        for fused in (True, False):
            m2 = mm(K, mc_t[:, :, None], fused)[:, :, 0] / tz[:, None]
            print(f"K mu_c / z ({'FMA chain' if fused else 'product+sum'}):", frac(m2.astype(f32), m2_t))


if __name__ == "__main__":
    main(*sys.argv[1:])
