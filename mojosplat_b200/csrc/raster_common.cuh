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

constexpr int kPairRec = 5;  // float4 per raster record

// The rasterizer's record of one Gaussian (5 x float4; every operand of the packed-pair walk stored duplicated):
//   q0 = {mx, mx, my, my}  q1 = {-A, -A, -B, -B}  q2 = {-C, -C, L, L}  q3 = {r, g, b, tau}  q4 = {hy, hx, special, Lt}
// One definition for the in-kernel staging, the record kernel and the projection epilogue, so that every path
// composites bit-identical values (products of two or three factors only: nothing here can be contracted).
__device__ __forceinline__ void pair_record_from(const float mx, const float my, const float ca, const float cb,
                                                 const float cc, const float op, const float cr, const float cg,
                                                 const float cbl, float4& q0, float4& q1, float4& q2, float4& q3,
                                                 float4& q4) {
    const float A = __fmul_rn(__fmul_rn(0.5f, kLog2e), ca), B = __fmul_rn(kLog2e, cb),
                C = __fmul_rn(__fmul_rn(0.5f, kLog2e), cc);
    // MUFU.LG2 (abs. error ~2^-22): alpha = 2^(L - q) stays within 1e-6 of o*exp(-sigma)
    const float L = (op > 0.0f) ? __log2f(op) : -INFINITY;
    const bool pd = (A > 0.0f) && (C > 0.0f) && (__fsub_rn(__fmul_rn(__fmul_rn(4.0f, A), C), __fmul_rn(B, B)) > 0.0f);
    float tau = pd ? __fsub_rn(L, kLog2AlphaThreshold) : INFINITY;
    if (!(op == op)) tau = INFINITY;  // NaN opacity: evaluate, never cull
    q0 = make_float4(mx, mx, my, my);
    q1 = make_float4(-A, -A, -B, -B);
    q2 = make_float4(-C, -C, L, L);
    q3 = make_float4(cr, cg, cbl, tau);
    // "plain" Gaussians (positive-definite conic, opacity <= 0.99, no NaN) have q >= 0 and alpha <= opacity by
    // construction: the walk may skip the sigma < 0 test and the 0.999 clamp (0.99, not 0.999: ex2.approx may
    // overshoot by an ulp).  q4.w is the bound of the sigma >= 0 test of the full walk: +inf for plain Gaussians,
    // so both walks treat them identically.  The edge minimisers hy, hx feed the conservative culling bound only:
    // approximate division is inside its slack.
    const bool plain = pd && (op <= 0.99f);
    q4 = make_float4(pd ? __fdividef(-B, __fmul_rn(2.0f, C)) : 0.0f, pd ? __fdividef(-B, __fmul_rn(2.0f, A)) : 0.0f,
                     plain ? 0.f : 1.f, plain ? INFINITY : L);
}

}  // namespace bsplat
