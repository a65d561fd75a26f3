// Shared by the forward (rasterize.cu) and the training-side (rasterize_bwd.cu) rasterizers.
#pragma once
#include "common.cuh"

namespace bsplat {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLog2AlphaThreshold = -7.994353436858858f;  // log2(1/255)

// Conservative test "can this Gaussian reach alpha >= 1/255 anywhere in the pixel rectangle [X0, X1] x [Y0, Y1]
// (pixel centres)?"  min of q over the rectangle = min over the (<= 2) edges facing the mean; along the edge
// u = ue the quadratic is D ue^2 + C (v - hy ue)^2 with D = A + B hy / 2 (and symmetrically for v = ve), so each
// edge costs a clamp and two FMAs.  NaN-safe: anything odd counts as a hit.
__device__ __forceinline__ bool pair_cull_hit(const float mx, const float my, const float A, const float B,
                                              const float C, const float tau, const float hy, const float hx,
                                              const float X0, const float X1, const float Y0, const float Y1) {
    const float u0 = mx - X1, u1 = mx - X0;
    const float v0 = my - Y1, v1 = my - Y0;
    const bool zu = (u0 <= 0.0f) && (u1 >= 0.0f);
    const bool zv = (v0 <= 0.0f) && (v1 >= 0.0f);
    float qmin = 0.0f;
    if (!(zu && zv)) {
        float q1 = INFINITY, q2 = INFINITY;
        if (!zu) {
            const float ue = (u0 > 0.0f) ? u0 : u1;
            const float vstar = hy * ue;
            const float dv = vstar - fminf(fmaxf(vstar, v0), v1);
            q1 = fmaf(fmaf(0.5f * B, hy, A) * ue, ue, C * dv * dv);
        }
        if (!zv) {
            const float ve = (v0 > 0.0f) ? v0 : v1;
            const float ustar = hx * ve;
            const float du = ustar - fminf(fmaxf(ustar, u0), u1);
            q2 = fmaf(fmaf(0.5f * B, hx, C) * ve, ve, A * du * du);
        }
        qmin = fminf(q1, q2);
    }
    const float um = fmaxf(fabsf(u0), fabsf(u1)), vm = fmaxf(fabsf(v0), fabsf(v1));
    const float slack = 4e-6f * (A * um * um + C * vm * vm) + 1e-3f;
    return !(qmin > tau + slack);
}

// Culling constants of one Gaussian in log2-folded units (A = 0.5 a log2e, B = b log2e, C = 0.5 c log2e):
// tau = log2(opacity) - log2(1/255) (+inf: never cull -- non-positive-definite conic or NaN opacity),
// hy = -B / (2C), hx = -B / (2A) (edge minimisers of the quadratic).
__device__ __forceinline__ void cull_constants(const float a, const float b, const float c, const float op,
                                               float& tau, float& hy, float& hx) {
    const float A = 0.5f * kLog2e * a, B = kLog2e * b, C = 0.5f * kLog2e * c;
    const float L = (op > 0.0f) ? __log2f(op) : -INFINITY;
    const bool pd = (A > 0.0f) && (C > 0.0f) && (4.0f * A * C - B * B > 0.0f);
    tau = pd ? (L - kLog2AlphaThreshold) : INFINITY;
    if (!(op == op)) tau = INFINITY;
    hy = pd ? __fdividef(-B, 2.0f * C) : 0.0f;
    hx = pd ? __fdividef(-B, 2.0f * A) : 0.0f;
}

}  // namespace bsplat
