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
    // branch-free (lanes of a warp test different Gaussians: divergent branches here cost more than the arithmetic)
    const float hB = 0.5f * B;
    const float ue = (u0 > 0.0f) ? u0 : u1;
    const float vstar = hy * ue;
    const float dv = vstar - fminf(fmaxf(vstar, v0), v1);
    float q1 = fmaf(fmaf(hB, hy, A) * ue, ue, C * dv * dv);
    const float ve = (v0 > 0.0f) ? v0 : v1;
    const float ustar = hx * ve;
    const float du = ustar - fminf(fmaxf(ustar, u0), u1);
    float q2 = fmaf(fmaf(hB, hx, C) * ve, ve, A * du * du);
    q1 = zu ? INFINITY : q1;
    q2 = zv ? INFINITY : q2;
    float qmin = fminf(q1, q2);
    qmin = (zu && zv) ? 0.0f : qmin;
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

constexpr int kPairRec = 3;  // float4 per raster record (48 B: a 12-word stride keeps per-lane LDS.128 conflict-free)

// The rasterizer's record of one Gaussian (3 x float4), operands stored ONCE: sm_100's packed FP32 instructions take a
// scalar register broadcast to both halves (`FFMA2 R, R.F32x2, R.F32, R.F32x2`), so the two-pixel walk needs no
// duplicated operands.
//   q0 = {mx, my, -A, -B}  q1 = {-C, L, r, g}  q2 = {b, tau*, hy, hx}
// tau* = tau with the lowest mantissa bit carrying the "special" flag (tau only feeds the conservative culling bound,
// whose slack is 1e-3; tau = +inf -- never cull -- only occurs for special Gaussians and becomes a NaN, which the
// NaN-safe bound also counts as a hit).
// One definition for the in-kernel staging, the record kernel and the projection epilogue, so that every path
// composites bit-identical values (products of two or three factors only: nothing here can be contracted).
__device__ __forceinline__ void pair_record_from(const float mx, const float my, const float ca, const float cb,
                                                 const float cc, const float op, const float cr, const float cg,
                                                 const float cbl, float4& q0, float4& q1, float4& q2) {
    const float A = __fmul_rn(__fmul_rn(0.5f, kLog2e), ca), B = __fmul_rn(kLog2e, cb),
                C = __fmul_rn(__fmul_rn(0.5f, kLog2e), cc);
    // MUFU.LG2 (abs. error ~2^-22): alpha = 2^(L - q) stays within 1e-6 of o*exp(-sigma)
    const float L = (op > 0.0f) ? __log2f(op) : -INFINITY;
    const bool pd = (A > 0.0f) && (C > 0.0f) && (__fsub_rn(__fmul_rn(__fmul_rn(4.0f, A), C), __fmul_rn(B, B)) > 0.0f);
    float tau = pd ? __fsub_rn(L, kLog2AlphaThreshold) : INFINITY;
    if (!(op == op)) tau = INFINITY;  // NaN opacity: evaluate, never cull
    // "plain" Gaussians (positive-definite conic, opacity <= 0.99, no NaN) have q >= 0 and alpha <= opacity by
    // construction: the walk may skip the sigma < 0 test and the 0.999 clamp (0.99, not 0.999: ex2.approx may
    // overshoot by an ulp).  The full walk applies the sigma >= 0 test (power <= L) to special Gaussians only, so
    // both walks treat plain ones identically.  The edge minimisers hy, hx feed the conservative culling bound only:
    // approximate division is inside its slack.
    const bool plain = pd && (op <= 0.99f);
    const uint32_t tb = (__float_as_uint(tau) & ~1u) | (plain ? 0u : 1u);
    q0 = make_float4(mx, my, -A, -B);
    q1 = make_float4(-C, L, cr, cg);
    q2 = make_float4(cbl, __uint_as_float(tb), pd ? __fdividef(-B, __fmul_rn(2.0f, C)) : 0.0f,
                     pd ? __fdividef(-B, __fmul_rn(2.0f, A)) : 0.0f);
}

// A record that can never hit (past the end of a list / invalid id, rasterization.mojo:109 guard): tau = -inf
__device__ __forceinline__ void pair_record_none(float4& q0, float4& q1, float4& q2) {
    q0 = make_float4(0.f, 0.f, 0.f, 0.f);
    q1 = make_float4(0.f, -INFINITY, 0.f, 0.f);
    q2 = make_float4(0.f, -INFINITY, 0.f, 0.f);
}

// The culling test on a record (global or shared memory): block = pixel-centre rectangle [X0, X1] x [Y0, Y1].
__device__ __forceinline__ bool pair_record_hit(const float4* r, const float X0, const float X1, const float Y0,
                                                const float Y1, bool* special) {
    const float4 p0 = r[0];
    const float nC = reinterpret_cast<const float*>(r + 1)[0];
    const float4 p2 = r[2];
    if (special) *special = (__float_as_uint(p2.y) & 1u) != 0u;
    return pair_cull_hit(p0.x, p0.y, -p0.z, -p0.w, -nC, p2.y, p2.z, p2.w, X0, X1, Y0, Y1);
}

// The same on a 32-bit shared-memory address held in a register (the walks use these: the compiler otherwise rebuilds
// the shared window base -- S2UR SR_CgaCtaId + UMOV + ULEA -- at every use).
__device__ __forceinline__ float4 lds_f4(const uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f2(const uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f1(const uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds_i32(const uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ bool pair_record_hit_at(const uint32_t addr, const float X0, const float X1, const float Y0,
                                                   const float Y1, bool* special) {
    const float4 p0 = lds_f4(addr);
    const float nC = lds_f1(addr + 16u);
    const float4 p2 = lds_f4(addr + 32u);
    if (special) *special = (__float_as_uint(p2.y) & 1u) != 0u;
    return pair_cull_hit(p0.x, p0.y, -p0.z, -p0.w, -nC, p2.y, p2.z, p2.w, X0, X1, Y0, Y1);
}

}  // namespace bsplat
