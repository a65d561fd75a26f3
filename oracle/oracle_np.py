"""Independent numpy restatement of the rasterizer (kernels/rasterization.mojo:75-162).

Vectorised over the pixels of one tile, Python loop over the tile's Gaussians: small cases
only.  It exists to cross-check oracle.c's rasterizer, which has no runnable reference
counterpart in this environment (see oracle.c header).  TEST INFRASTRUCTURE ONLY.
"""
import math

import numpy as np


def rasterize_np(means2d, conics, colors, opacities, background, tile_ranges, sorted_ids,
                 W, H, tile_size=16):
    f = np.float32
    means2d = np.asarray(means2d, f); conics = np.asarray(conics, f)
    colors = np.asarray(colors, f); opacities = np.asarray(opacities, f).reshape(-1)
    background = np.asarray(background, f).reshape(-1)
    C = colors.shape[1]
    th, tw = math.ceil(H / tile_size), math.ceil(W / tile_size)
    ranges = np.asarray(tile_ranges).reshape(th, tw, 2)
    img = np.zeros((H, W, C), f)
    for tr in range(th):
        for tc in range(tw):
            i0, j0 = tr * tile_size, tc * tile_size
            i1, j1 = min(i0 + tile_size, H), min(j0 + tile_size, W)
            py, px = np.meshgrid(np.arange(i0, i1, dtype=f) + f(0.5),
                                 np.arange(j0, j1, dtype=f) + f(0.5), indexing="ij")
            T = np.ones_like(px)
            done = np.zeros(px.shape, bool)
            out = np.zeros(px.shape + (C,), f)
            for k in range(int(ranges[tr, tc, 0]), int(ranges[tr, tc, 1])):
                g = int(sorted_ids[k])
                dx = means2d[g, 0] - px
                dy = means2d[g, 1] - py
                a, b, c = conics[g]
                sigma = f(0.5) * (a * dx * dx + c * dy * dy) + b * dx * dy
                alpha = np.minimum(opacities[g] * np.exp(-sigma, dtype=f), f(0.999))
                skip = (sigma < 0) | (alpha < f(1.0 / 255.0))
                next_T = T * (f(1.0) - alpha)
                stop = (~skip) & (~done) & (next_T <= f(1e-4))
                done |= stop
                live = (~skip) & (~done)
                vis = alpha * T
                out[live] += colors[g][None, :] * vis[live][:, None]
                T = np.where(live, next_T, T)
            img[i0:i1, j0:j1] = out + T[..., None] * background
    return img


def sh_eval_np(degree, coeffs, means3d, campos):
    """float64 restatement of the standard 3DGS spherical-harmonics colour (real SH, degree <= 3):
    colour = max(sum_k Y_k(dir) c_k + 0.5, 0), dir = normalize(mean - campos).  The reference only has a
    placeholder (render.py:82-87); this follows the published basis (constants of the 3DGS paper's code).
    TEST INFRASTRUCTURE ONLY."""
    c = np.asarray(coeffs, np.float64)
    d = np.asarray(means3d, np.float64) - np.asarray(campos, np.float64)[None, :]
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-15)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    out = 0.28209479177387814 * c[:, 0]
    if degree >= 1:
        out = out + 0.4886025119029199 * (-y * c[:, 1] + z * c[:, 2] - x * c[:, 3])
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out = out + (1.0925484305920792 * xy * c[:, 4] - 1.0925484305920792 * yz * c[:, 5]
                     + 0.31539156525252005 * (2 * zz - xx - yy) * c[:, 6] - 1.0925484305920792 * xz * c[:, 7]
                     + 0.5462742152960396 * (xx - yy) * c[:, 8])
    if degree >= 3:
        out = out + (-0.5900435899266435 * y * (3 * xx - yy) * c[:, 9] + 2.890611442640554 * xy * z * c[:, 10]
                     - 0.4570457994644658 * y * (4 * zz - xx - yy) * c[:, 11]
                     + 0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy) * c[:, 12]
                     - 0.4570457994644658 * x * (4 * zz - xx - yy) * c[:, 13]
                     + 1.445305721320277 * z * (xx - yy) * c[:, 14] - 0.5900435899266435 * x * (xx - 3 * yy) * c[:, 15])
    return np.maximum(out + 0.5, 0.0)
