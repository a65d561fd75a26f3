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
