// Stage 3 -- tile rasterizer: front-to-back alpha compositing of the per-tile sorted lists.
//
// Replaces the MAX op `rasterize_to_pixels_3dgs_fwd` (mojosplat/kernels/rasterization.mojo:16-162,
// called from mojosplat/rasterization.py:127-186) and gsplat's rasterize_to_pixels
// (rasterization.py:81-124).  Per pixel p = (j+0.5, i+0.5), for the tile's Gaussians front to back:
//   sigma = 0.5 (a dx^2 + c dy^2) + b dx dy ; alpha = min(0.999, o exp(-sigma))
//   skip if sigma < 0 or alpha < 1/255 ; stop if T (1-alpha) <= 1e-4 ; out += colour alpha T ; T *= 1-alpha
//   image = out + T background                                   (rasterization.mojo:138-162)
//
// Two kernels:
//   raster_faithful_kernel  any tile size <= 32 and any channel count; arithmetic in the exact
//       operation order of the Mojo kernel (separately rounded products, expf).  Parity anchor
//       and fallback for non-default shapes.
//   raster_fast_kernel      16x16 tiles, RGB.  One warp owns an 8x4 pixel block; staged Gaussians
//       carry log2-folded conics so a pixel costs 2 FADD + 2 FMUL + 3 FFMA + 1 MUFU.EX2; each warp
//       first tests 32 staged Gaussians at once (one per lane) against its 8x4 block with an exact
//       conservative ellipse/rectangle bound and then only walks the survivors (warp ballot),
//       warps retire when all 32 pixels are saturated, the CTA retires on __syncthreads_count,
//       the tile is written with 128-bit stores.  Skipping is exact: a skipped Gaussian has
//       alpha < 1/255 on every pixel of the block, so the composited result is unchanged.
// No tensor cores: no stage of this path is a dense contraction.
#include <type_traits>

#include "raster_common.cuh"

namespace bsplat {

constexpr float kAlphaThreshold = 1.0f / 255.0f;

// ------------------------------------------------------------------------------------------
// faithful kernel
// ------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(1024) raster_faithful_kernel(const int64_t N, const int cdim, const int c0,
                                       const float* __restrict__ means2d, const float* __restrict__ conics,
                                       const float* __restrict__ colors, const float* __restrict__ opacities,
                                       const float* __restrict__ background,
                                       const int32_t* __restrict__ tile_ranges,
                                       const int32_t* __restrict__ sorted_ids, const int W, const int H,
                                       const int ts, const int tiles_w, const int row_begin,
                                       float* __restrict__ image, unsigned long long* __restrict__ stats,
                                       const unsigned long long* __restrict__ m_dev) {
    extern __shared__ float s_buf[];
    const int nthreads = ts * ts;
    float* s_mx = s_buf;
    float* s_my = s_mx + nthreads;
    float* s_a = s_my + nthreads;
    float* s_b = s_a + nthreads;
    float* s_c = s_b + nthreads;
    float* s_o = s_c + nthreads;
    float* s_col = s_o + nthreads;  // [nthreads][CH]

    const int tid = threadIdx.y * ts + threadIdx.x;
    const int tile_row = row_begin + (int)blockIdx.y;
    const int tile = tile_row * tiles_w + blockIdx.x;
    const int i = tile_row * ts + threadIdx.y;
    const int j = blockIdx.x * ts + threadIdx.x;
    const bool inside = (i < H) && (j < W);
    bool done = !inside;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;

    const int32_t r0 = tile_ranges[2 * tile], r1 = tile_ranges[2 * tile + 1];
    float T = 1.0f;
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
    unsigned int n_eval = 0, n_pass = 0;

    for (int32_t b0 = r0; b0 < r1; b0 += nthreads) {
        // barrier before the staging buffers are overwritten; also the whole-tile early exit
        if (__syncthreads_count(done) >= nthreads) break;
        const int32_t idx = b0 + tid;
        if (idx < r1) {
            const int32_t g = sorted_ids[idx];
            if (g >= 0 && (int64_t)g < N) {
                s_mx[tid] = means2d[2 * (int64_t)g];
                s_my[tid] = means2d[2 * (int64_t)g + 1];
                s_a[tid] = conics[3 * (int64_t)g];
                s_b[tid] = conics[3 * (int64_t)g + 1];
                s_c[tid] = conics[3 * (int64_t)g + 2];
                s_o[tid] = opacities[g];
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    s_col[tid * CH + c] = (c0 + c < cdim) ? colors[(int64_t)g * cdim + c0 + c] : 0.0f;
            } else {
                // out-of-range id: the Mojo staging guard (rasterization.mojo:109) skips it
                s_mx[tid] = s_my[tid] = s_a[tid] = s_b[tid] = s_c[tid] = 0.0f;
                s_o[tid] = __int_as_float(0x7fc00000);  // NaN marks "not a Gaussian"
            }
        }
        __syncthreads();
        if (!done) {
            const int bs = min(nthreads, (int)(r1 - b0));
            for (int t = 0; t < bs; ++t) {
                const float op = s_o[t];
                if (op != op) continue;
                ++n_eval;
                const float dx = __fsub_rn(s_mx[t], px), dy = __fsub_rn(s_my[t], py);
                const float a = s_a[t], b = s_b[t], c = s_c[t];
                // 0.5 * (a*dx*dx + c*dy*dy) + b*dx*dy, every operation rounded separately
                const float q = __fadd_rn(__fmul_rn(__fmul_rn(a, dx), dx), __fmul_rn(__fmul_rn(c, dy), dy));
                const float sigma = __fadd_rn(__fmul_rn(0.5f, q), __fmul_rn(__fmul_rn(b, dx), dy));
                float alpha = __fmul_rn(op, expf(-sigma));
                alpha = fminf(alpha, 0.999f);
                if (sigma < 0.0f || alpha < kAlphaThreshold) continue;
                const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                if (next_T <= 1e-4f) { done = true; break; }
                const float vis = __fmul_rn(alpha, T);
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) acc[ch] = __fadd_rn(acc[ch], __fmul_rn(s_col[t * CH + ch], vis));
                T = next_T;
                ++n_pass;
            }
        }
    }
    if (inside) {
        // sync-free frames: no intersections at all => all-zero image (render.py:73-76), decided on the device
        const bool empty = (m_dev != nullptr) && (*m_dev == 0ull);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
            if (c0 + ch < cdim)
                image[((int64_t)i * W + j) * cdim + c0 + ch] =
                    empty ? 0.0f : __fadd_rn(acc[ch], __fmul_rn(T, background[c0 + ch]));
    }
    if (stats != nullptr) {
        // blocks may end in a partial warp (tile sizes like 10): plain per-thread atomics
        if (n_eval) atomicAdd(stats, (unsigned long long)n_eval);
        if (n_pass) atomicAdd(stats + 1, (unsigned long long)n_pass);
    }
}

// ------------------------------------------------------------------------------------------
// fast kernel (tile 16, RGB)
// ------------------------------------------------------------------------------------------
constexpr int kFastTile = 16;
constexpr int kFastThreads = 256;
constexpr int kFastBatch = 256;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Staged Gaussian t (log2-folded, 48 B): s_g[3t] = {mx, my, A, B}, s_g[3t+1] = {C, L, hy, hx},
// s_g[3t+2] = {r, g, b, tau} with A = 0.5 a log2e, B = b log2e, C = 0.5 c log2e, L = log2(opacity),
// hy = -B/(2C), hx = -B/(2A) (edge minimisers of the quadratic), tau = L - log2(1/255)
// (+inf: never cull, -inf: not a Gaussian).  alpha = 2^(L - (A dx^2 + B dx dy + C dy^2)).
template <bool kCull>
__global__ void __launch_bounds__(kFastThreads)
raster_fast_kernel(const int64_t N, const float* __restrict__ means2d, const float* __restrict__ conics,
                   const float* __restrict__ colors, const float* __restrict__ opacities,
                   const float* __restrict__ background, const int32_t* __restrict__ tile_ranges,
                   const int32_t* __restrict__ tile_order, const int first_tile,
                   const int32_t* __restrict__ sorted_ids, const int W, const int H, const int tiles_w,
                   float* __restrict__ image, const int vec_store,
                   const unsigned long long* __restrict__ m_dev) {
    __shared__ float4 s_g[kFastBatch * 3];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // heavy tiles first (tile_order, longest lists first) so that no long list starts at the tail
    const int tile = tile_order ? __ldg(tile_order + blockIdx.x) : first_tile + (int)blockIdx.x;
    const int tile_y = tile / tiles_w, tile_x = tile - tile_y * tiles_w;
    // warp -> 8x4 pixel block inside the tile; lane -> pixel inside the block
    const int bx = tile_x * kFastTile + (warp & 1) * 8;
    const int by = tile_y * kFastTile + (warp >> 1) * 4;
    const int j = bx + (lane & 7);
    const int i = by + (lane >> 3);
    const bool inside = (i < H) && (j < W);
    // a finished pixel gets an infinite x: every later Gaussian then fails the alpha test by itself,
    // and "done" is simply px == inf (no separate flag to maintain in the inner loop)
    float px = inside ? (float)j + 0.5f : INFINITY;
    const float py = (float)i + 0.5f;
    // pixel-centre extent of this warp's block (clipped blocks only get more conservative)
    const float X0 = (float)bx + 0.5f, X1 = (float)bx + 7.5f;
    const float Y0 = (float)by + 0.5f, Y1 = (float)by + 3.5f;

    const int32_t r0 = tile_ranges[2 * tile], r1 = tile_ranges[2 * tile + 1];
    float T = 1.0f, accr = 0.0f, accg = 0.0f, accb = 0.0f;

    for (int32_t b0 = r0; b0 < r1; b0 += kFastBatch) {
        if (__syncthreads_count(!(px < INFINITY)) >= kFastThreads) break;
        const int32_t idx = b0 + tid;
        if (idx < r1) {
            const int32_t g = __ldg(sorted_ids + idx);
            float4 ga, gb, gc;
            if (g >= 0 && (int64_t)g < N) {
                const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + g);
                const float ca = __ldg(conics + 3 * (int64_t)g), cb = __ldg(conics + 3 * (int64_t)g + 1),
                            cc = __ldg(conics + 3 * (int64_t)g + 2);
                const float op = __ldg(opacities + g);
                const float A = 0.5f * kLog2e * ca, B = kLog2e * cb, C = 0.5f * kLog2e * cc;
                // accurate log2 keeps alpha = 2^(L - q) within a few ulp of o*exp(-sigma)
                const float L = (op > 0.0f) ? log2f(op) : -INFINITY;
                const bool pd = (A > 0.0f) && (C > 0.0f) && (4.0f * A * C - B * B > 0.0f);
                ga = make_float4(m.x, m.y, A, B);
                gb = make_float4(C, L, pd ? -B / (2.0f * C) : 0.0f, pd ? -B / (2.0f * A) : 0.0f);
                float tau = pd ? (L - kLog2AlphaThreshold) : INFINITY;
                if (!(op == op)) tau = INFINITY;  // NaN opacity: evaluate, never cull
                gc = make_float4(__ldg(colors + 3 * (int64_t)g), __ldg(colors + 3 * (int64_t)g + 1),
                                 __ldg(colors + 3 * (int64_t)g + 2), tau);
            } else {
                ga = make_float4(0.f, 0.f, 0.f, 0.f);
                gb = make_float4(0.f, -INFINITY, 0.f, 0.f);
                gc = make_float4(0.f, 0.f, 0.f, -INFINITY);
            }
            s_g[3 * tid] = ga; s_g[3 * tid + 1] = gb; s_g[3 * tid + 2] = gc;
        }
        __syncthreads();

        const int bs = min(kFastBatch, (int)(r1 - b0));
        // warp-uniform loop; a warp whose 32 pixels are all saturated just falls through
        for (int c0 = 0; c0 < bs; c0 += 32) {
            if (__all_sync(0xffffffffu, !(px < INFINITY))) break;
            unsigned int mask;
            if (kCull) {
                bool hit = false;
                // lane l tests Gaussian c0 + 31 - l: the earliest Gaussian is the HIGHEST ballot bit, so the
                // walk below needs a single FLO (clz) per survivor instead of BREV + FLO
                const int gi = c0 + 31 - lane;
                if (gi < bs) {
                    const float4 a = s_g[3 * gi];
                    const float4 b = s_g[3 * gi + 1];
                    const float tau = s_g[3 * gi + 2].w;
                    // u = mx - x over the block, v = my - y
                    const float u0 = a.x - X1, u1 = a.x - X0;
                    const float v0 = a.y - Y1, v1 = a.y - Y0;
                    const bool zu = (u0 <= 0.0f) && (u1 >= 0.0f);
                    const bool zv = (v0 <= 0.0f) && (v1 >= 0.0f);
                    // min of q over the block = min over the (<= 2) edges facing the mean; along the edge
                    // u = ue the quadratic is D ue^2 + C (v - hy ue)^2 with D = A - B^2/(4C) = A + B hy / 2
                    // (and symmetrically E = C + B hx / 2), so each edge costs a clamp and two FMAs.
                    float qmin = 0.0f;
                    if (!(zu && zv)) {
                        float q1 = INFINITY, q2 = INFINITY;
                        if (!zu) {
                            const float ue = (u0 > 0.0f) ? u0 : u1;
                            const float vstar = b.z * ue;
                            const float dv = vstar - fminf(fmaxf(vstar, v0), v1);
                            q1 = fmaf(fmaf(0.5f * a.w, b.z, a.z) * ue, ue, b.x * dv * dv);
                        }
                        if (!zv) {
                            const float ve = (v0 > 0.0f) ? v0 : v1;
                            const float ustar = b.w * ve;
                            const float du = ustar - fminf(fmaxf(ustar, u0), u1);
                            q2 = fmaf(fmaf(0.5f * a.w, b.w, b.x) * ve, ve, a.z * du * du);
                        }
                        qmin = fminf(q1, q2);
                    }
                    const float um = fmaxf(fabsf(u0), fabsf(u1)), vm = fmaxf(fabsf(v0), fabsf(v1));
                    const float slack = 4e-6f * (a.z * um * um + b.x * vm * vm) + 1e-3f;
                    hit = !(qmin > tau + slack);  // NaN-safe: anything odd counts as a hit
                }
                mask = __ballot_sync(0xffffffffu, hit);
            } else {
                const int rem = bs - c0;
                mask = rem >= 32 ? 0xffffffffu : ~(0xffffffffu >> rem);  // Gaussian c0+k <-> bit 31-k
            }
            const float4* rec_hi = s_g + 3 * (c0 + 31);  // record of ballot bit 0; bit b is 3*b records earlier
            while (mask) {
                // highest set bit = next Gaussian, front to back (bfind -> a single FLO; written in PTX
                // because nvcc rewrites 31 - clz(x) into a longer clz-based sequence)
                unsigned int b_hi;
                asm("bfind.u32 %0, %1;" : "=r"(b_hi) : "r"(mask));
                mask ^= 1u << b_hi;
                const float4* r = rec_hi - 3 * (int)b_hi;
                const float4 a = r[0];
                const float2 b = *reinterpret_cast<const float2*>(r + 1);
                const float4 c = r[2];
                const float dx = a.x - px, dy = a.y - py;
                const float t1 = fmaf(a.w, dy, a.z * dx);
                const float q = fmaf(b.x * dy, dy, t1 * dx);
                const float power = b.y - q;
                // branch-free body: selects instead of divergent paths (90 % of the walked Gaussians
                // contribute to at least one pixel of the block, so a skip branch saves nothing)
                const bool pass = (q >= 0.0f) && (power >= kLog2AlphaThreshold);
                const float alpha = fminf(0.999f, ex2_approx(power));
                const float next_T = T * (1.0f - alpha);
                const bool live = next_T > 1e-4f;
                const float vis = (pass && live) ? alpha * T : 0.0f;
                accr = fmaf(c.x, vis, accr);
                accg = fmaf(c.y, vis, accg);
                accb = fmaf(c.z, vis, accb);
                T = (pass && live) ? next_T : T;
                // saturated: this Gaussian is not added (rasterization.mojo:146-150) and the pixel retires
                px = (pass && !live) ? INFINITY : px;
            }
        }
    }

    // ---- write the tile: through shared memory as 128-bit rows when the layout allows ----
    // sync-free frames: no intersections at all => all-zero image (render.py:73-76), decided on the device
    const float bgs = (m_dev != nullptr && *m_dev == 0ull) ? 0.0f : 1.0f;
    const float outr = fmaf(T, bgs * __ldg(background), accr), outg = fmaf(T, bgs * __ldg(background + 1), accg),
                outb = fmaf(T, bgs * __ldg(background + 2), accb);
    const bool full_tile = (tile_x * kFastTile + kFastTile <= W) && (tile_y * kFastTile + kFastTile <= H);
    if (vec_store && full_tile) {
        __syncthreads();  // staging buffers are dead from here on
        float* s_out = reinterpret_cast<float*>(s_g);  // 16 rows x 48 floats = 3 KB
        const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
        s_out[(ly * kFastTile + lx) * 3 + 0] = outr;
        s_out[(ly * kFastTile + lx) * 3 + 1] = outg;
        s_out[(ly * kFastTile + lx) * 3 + 2] = outb;
        __syncthreads();
        if (tid < 16 * 12) {
            const int row = tid / 12, c4 = tid % 12;
            const float4 v = reinterpret_cast<const float4*>(s_out)[row * 12 + c4];
            float* dst = image + ((int64_t)(tile_y * kFastTile + row) * W + tile_x * kFastTile) * 3;
            reinterpret_cast<float4*>(dst)[c4] = v;
        }
    } else if (inside) {
        float* dst = image + ((int64_t)i * W + j) * 3;
        dst[0] = outr; dst[1] = outg; dst[2] = outb;
    }
}

// ------------------------------------------------------------------------------------------
// pair kernel (tile 16, RGB): the default fast path.
// sm_100 executes packed FP32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot, two lanes of work), and the
// rasterizer is bound by instruction issue, not by the FP32 pipe.  So each lane owns TWO pixels -- (x, y) and
// (x, y + 4) of the warp's 8x8 block -- and the whole quadratic runs as f32x2 instructions on
// register pairs.  Staged records keep every per-Gaussian operand duplicated {v, v}, so the pairs come
// straight out of LDS.128 with no packing moves.  4 warps (128 threads) per 16x16 tile.
//   record t (80 B): s[5t] = {mx, mx, my, my}  s[5t+1] = {-A, -A, -B, -B}  s[5t+2] = {-C, -C, L, L}
//                    s[5t+3] = {r, g, b, tau}  s[5t+4] = {hy, hx, -, -}
//   power = L - (A dx^2 + B dx dy + C dy^2) as  fma2(fma2(-A, dx, -B dy), dx, fma2(-C dy, dy, L))
// A finished pixel carries -x = -inf (every later power is -inf or NaN and fails the alpha test).
// Same culling as above on the 8x8 block; a skipped Gaussian has alpha < 1/255 on all 64 pixels.
// ------------------------------------------------------------------------------------------
constexpr int kPairThreads = 128;
constexpr int kPairBatch = 256;
constexpr int kPairRec = 5;  // float4 per staged Gaussian

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

constexpr int kLongTile = 2048;  // lists longer than this get a tile-level pre-test (see the kernel)

struct LoopConsts {
    unsigned int one_u;
    float one_f, mone_f;
};

// The staged record of one Gaussian (5 x float4, layout above).  One definition for the in-kernel staging and for
// raster_pair_prep_kernel, so both paths composite bit-identical values.
__device__ __forceinline__ void pair_record(const int64_t g, const float* __restrict__ means2d,
                                            const float* __restrict__ conics, const float* __restrict__ colors,
                                            const float* __restrict__ opacities, float4& q0, float4& q1, float4& q2,
                                            float4& q3, float4& q4) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + g);
    const float ca = __ldg(conics + 3 * g), cb = __ldg(conics + 3 * g + 1), cc = __ldg(conics + 3 * g + 2);
    const float op = __ldg(opacities + g);
    const float A = 0.5f * kLog2e * ca, B = kLog2e * cb, C = 0.5f * kLog2e * cc;
    // MUFU.LG2 (abs. error ~2^-22): alpha = 2^(L - q) stays within 1e-6 of o*exp(-sigma)
    const float L = (op > 0.0f) ? __log2f(op) : -INFINITY;
    const bool pd = (A > 0.0f) && (C > 0.0f) && (4.0f * A * C - B * B > 0.0f);
    float tau = pd ? (L - kLog2AlphaThreshold) : INFINITY;
    if (!(op == op)) tau = INFINITY;  // NaN opacity: evaluate, never cull
    q0 = make_float4(m.x, m.x, m.y, m.y);
    q1 = make_float4(-A, -A, -B, -B);
    q2 = make_float4(-C, -C, L, L);
    q3 = make_float4(__ldg(colors + 3 * g), __ldg(colors + 3 * g + 1), __ldg(colors + 3 * g + 2), tau);
    // "plain" Gaussians (positive-definite conic, opacity <= 0.99, no NaN) have q >= 0 and alpha <= opacity by
    // construction: the walk may skip the sigma < 0 test and the 0.999 clamp (0.99, not 0.999: ex2.approx may
    // overshoot by an ulp).  q4.w is the bound of the sigma >= 0 test of the full walk: +inf for plain Gaussians,
    // so both walks treat them identically.  The edge minimisers hy, hx feed the conservative culling bound only:
    // approximate division is inside its slack.
    const bool plain = pd && (op <= 0.99f);
    q4 = make_float4(pd ? __fdividef(-B, 2.0f * C) : 0.0f, pd ? __fdividef(-B, 2.0f * A) : 0.0f, plain ? 0.f : 1.f,
                     plain ? INFINITY : L);
}

// Records of ALL Gaussians, once per frame (80 B each): with them the rasterizer's staging is a pure gather that
// cp.async can run one batch ahead, and the per-(tile, Gaussian) staging arithmetic disappears.
// With a list (row-band frames: the band's Gaussians in depth order, count on the device) only those get a record.
__global__ void __launch_bounds__(256)
raster_pair_prep_kernel(const int64_t N, const float* __restrict__ means2d, const float* __restrict__ conics,
                        const float* __restrict__ colors, const float* __restrict__ opacities,
                        float4* __restrict__ rec, const int32_t* __restrict__ list,
                        const unsigned long long* __restrict__ list_n) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (list != nullptr) {
        if (g >= (int64_t)(*list_n)) return;
        g = __ldg(list + g);
    }
    if (g < 0 || g >= N) return;
    float4 q0, q1, q2, q3, q4;
    pair_record(g, means2d, conics, colors, opacities, q0, q1, q2, q3, q4);
    float4* d = rec + kPairRec * g;
    d[0] = q0; d[1] = q1; d[2] = q2; d[3] = q3; d[4] = q4;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// kRec: the Gaussians come as prepared records (raster_pair_prep_kernel); batches of 128 are gathered by sorted id
// with cp.async into the two halves of the staging buffer, one batch ahead of the walk (ids two batches ahead),
// one barrier per batch.  !kRec: workspace-free staging of 256 per batch from the raw arrays.
// kMbar (with kRec): no CTA barrier in the loop at all.  Three stages; the cp.async copies of a batch arrive on an
// mbarrier by themselves ("full": the batch has landed, whatever the issuing threads are doing meanwhile), each
// warp arrives on a second mbarrier when it has consumed a stage ("empty"), and a thread waits for "empty" of
// batch b-1 before it gathers batch b+2 into the same stage -- so a warp with little to walk runs up to two batches
// ahead of the slowest one instead of waiting for it at every batch.
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok) __nanosleep(40);  // polite: a waiting warp must not take issue slots from the walking ones
    } while (!ok);
}

constexpr int kMbarStages = 3;
constexpr int kMbarBatch = 96;  // entries per stage (3 x 96 x 80 B = 23 KB: keeps 8-9 CTAs per SM)

template <bool kCull, bool kRec, bool kMbar = false>
__global__ void __launch_bounds__(kPairThreads)
raster_pair_kernel(const int64_t N, const float4* __restrict__ rec, const float* __restrict__ means2d, const float* __restrict__ conics,
                   const float* __restrict__ colors, const float* __restrict__ opacities,
                   const float* __restrict__ background, const int32_t* __restrict__ tile_ranges,
                   const int32_t* __restrict__ tile_order, const int first_tile,
                   const int32_t* __restrict__ sorted_ids, const int W, const int H, const int tiles_w,
                   float* __restrict__ image, const int vec_store,
                   const unsigned long long* __restrict__ m_dev, const PeerImages peers, const LoopConsts consts) {
    __shared__ float4 s_g[(kMbar ? kMbarStages * kMbarBatch : kPairBatch) * kPairRec];
    __shared__ unsigned int s_tmask[kPairBatch / 32];  // long tiles: survivors of the tile-level test
    __shared__ unsigned long long s_full[kMbarStages], s_empty[kMbarStages];
    __shared__ int s_done_warps;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tile = tile_order ? __ldg(tile_order + blockIdx.x) : first_tile + (int)blockIdx.x;
    const int tile_y = tile / tiles_w, tile_x = tile - tile_y * tiles_w;
    // warp -> 8x8 pixel block; lane -> column (lane & 7), rows (lane >> 3) and (lane >> 3) + 4
    const int bx = tile_x * kFastTile + (warp & 1) * 8;
    const int by = tile_y * kFastTile + (warp >> 1) * 8;
    const int j = bx + (lane & 7);
    const int i0 = by + (lane >> 3), i1 = i0 + 4;
    const bool in0 = (i0 < H) && (j < W), in1 = (i1 < H) && (j < W);
    float npx0 = in0 ? -((float)j + 0.5f) : -INFINITY;  // negated x per pixel; -inf = finished
    float npx1 = in1 ? -((float)j + 0.5f) : -INFINITY;
    const f32x2 npy = pk2(-((float)i0 + 0.5f), -((float)i1 + 0.5f));
    const float X0 = (float)bx + 0.5f, X1 = (float)bx + 7.5f;
    const float Y0 = (float)by + 0.5f, Y1 = (float)by + 7.5f;

    const int32_t r0 = tile_ranges[2 * tile], r1 = tile_ranges[2 * tile + 1];
    // Very long lists are mostly Gaussians that cannot touch the tile at all (under the torch binning rules every
    // culled Gaussian is clamped into a border tile: ~10 k entries in each corner tile at config 3, ~350 k at
    // config 5).  For those tiles the four warps first share ONE test of every staged Gaussian against the whole
    // 16x16 tile (64 entries per warp instead of 256), and a warp only runs its own 8x8 test on the survivors:
    // the serial walk of such a list, which bounds the kernel once a frame is split across GPUs, gets ~4x shorter.
    const bool long_tile = kCull && !kMbar && (r1 - r0 > kLongTile);  // (needs a CTA barrier per batch)
    const float TX0 = (float)(tile_x * kFastTile) + 0.5f, TX1 = TX0 + 15.0f;
    const float TY0 = (float)(tile_y * kFastTile) + 0.5f, TY1 = TY0 + 15.0f;
    float T0 = 1.0f, T1 = 1.0f;
    float ar0 = 0.f, ag0 = 0.f, ab0 = 0.f, ar1 = 0.f, ag1 = 0.f, ab1 = 0.f;
    // loop constants come in as kernel parameters (constant bank operands): ptxas otherwise re-materialises
    // literal constants with a MOV in every iteration of the walk, which is bound by issue slots
    const unsigned int bit_one = consts.one_u;
    const f32x2 one2 = pk2(consts.one_f, consts.one_f), mone2 = pk2(consts.mone_f, consts.mone_f);

    constexpr int kBatch = kMbar ? kMbarBatch : (kRec ? kPairThreads : kPairBatch);
    constexpr int kPer = kBatch / kPairThreads;  // staged entries per thread and batch
    auto load_id = [&](int32_t at) { return (at < r1) ? __ldg(sorted_ids + at) : -1; };
    auto gather = [&](int half, int32_t id) {  // kRec: this thread's entry of a batch -> record buffer `half`
        float4* dst = s_g + (half * kPairThreads + tid) * kPairRec;
        if (id >= 0 && (int64_t)id < N) {
            const float4* src = rec + kPairRec * (int64_t)id;
#pragma unroll
            for (int q = 0; q < kPairRec; ++q) cp_async16(dst + q, src + q);
        } else {  // past the end of the list / invalid id (rasterization.mojo:109 guard): can never hit
            dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[1] = dst[0];
            dst[2] = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
            dst[3] = make_float4(0.f, 0.f, 0.f, -INFINITY);
            dst[4] = dst[0];
        }
        cp_async_commit();
    };
    int32_t id_next = -1;
    bool warp_done = false;  // kMbar: this warp's 64 pixels are saturated (it keeps gathering and arriving)
    if (kMbar) {
        if (tid == 0) {
#pragma unroll
            for (int st = 0; st < kMbarStages; ++st) { mbar_init(&s_full[st], kMbarBatch); mbar_init(&s_empty[st], kPairThreads / 32); }
            s_done_warps = 0;
        }
        __syncthreads();
    }
    // kMbar: this thread's entry of batch b -> stage b % 3; the copies (or the sentinel store) arrive on "full"
    auto gather_mbar = [&](int b, int32_t id) {
        const int st = b % kMbarStages;
        if (tid >= kMbarBatch) return;  // the last warp has no entry to fetch
        float4* dst = s_g + (st * kMbarBatch + tid) * kPairRec;
        if (id >= 0 && (int64_t)id < N) {
            const float4* src = rec + kPairRec * (int64_t)id;
#pragma unroll
            for (int q = 0; q < kPairRec; ++q) cp_async16(dst + q, src + q);
            mbar_cp_async_arrive(&s_full[st]);
        } else {
            dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[1] = dst[0];
            dst[2] = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
            dst[3] = make_float4(0.f, 0.f, 0.f, -INFINITY);
            dst[4] = dst[0];
            mbar_arrive(&s_full[st]);
        }
    };
    const int n_batches = (int)((r1 - r0 + kBatch - 1) / kBatch);
    if (kMbar) {
        if (n_batches > 0) gather_mbar(0, load_id(r0 + tid));
        if (n_batches > 1) gather_mbar(1, load_id(r0 + kBatch + tid));
        id_next = load_id(r0 + 2 * kBatch + tid);
    } else if (kRec && r0 < r1) {
        gather(0, load_id(r0 + tid));
        id_next = load_id(r0 + kBatch + tid);
    }
    int batch = 0;
    for (int32_t b0 = r0; b0 < r1; b0 += kBatch, ++batch) {
        const bool fin = !(npx0 > -INFINITY) && !(npx1 > -INFINITY);
        const float4* s_rec = s_g;
        if (kMbar) {
            if (*reinterpret_cast<volatile int*>(&s_done_warps) >= kPairThreads / 32) break;  // every warp is done
            if (batch + 2 < n_batches) {
                // stage (batch + 2) % 3 was last read for batch - 1: wait until all four warps have consumed it
                if (batch >= 1) mbar_wait(&s_empty[(batch - 1) % kMbarStages], ((batch - 1) / kMbarStages) & 1);
                gather_mbar(batch + 2, id_next);
                id_next = load_id(b0 + 3 * kBatch + tid);
            }
            mbar_wait(&s_full[batch % kMbarStages], (batch / kMbarStages) & 1);  // batch has landed
            s_rec = s_g + (batch % kMbarStages) * kMbarBatch * kPairRec;
        } else if (kRec) {
            cp_async_wait_all();  // this thread's part of batch `batch` has landed ...
            if (__syncthreads_count(fin) >= kPairThreads) break;  // ... everyone's has; batch - 1 is fully consumed
            if (b0 + kBatch < r1) gather((batch + 1) & 1, id_next);  // next batch flies during this walk
            id_next = load_id(b0 + 2 * kBatch + tid);
            s_rec = s_g + (batch & 1) * kPairThreads * kPairRec;
        } else {
            if (__syncthreads_count(fin) >= kPairThreads) break;
#pragma unroll
            for (int h = 0; h < kPer; ++h) {
                const int t = tid + h * kPairThreads;
                const int32_t idx = b0 + t;
                if (idx < r1) {
                    const int32_t g = __ldg(sorted_ids + idx);
                    float4 q0, q1, q2, q3, q4;
                    if (g >= 0 && (int64_t)g < N) {
                        pair_record(g, means2d, conics, colors, opacities, q0, q1, q2, q3, q4);
                    } else {
                        q0 = make_float4(0.f, 0.f, 0.f, 0.f);
                        q1 = q0;
                        q2 = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
                        q3 = make_float4(0.f, 0.f, 0.f, -INFINITY);
                        q4 = q0;
                    }
                    float4* dst = s_g + kPairRec * t;
                    dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3; dst[4] = q4;
                }
            }
            __syncthreads();
        }

        const int bs = min(kBatch, (int)(r1 - b0));
        if (long_tile) {
#pragma unroll
            for (int h = 0; h < kPer; ++h) {
                const int e = warp * (32 * kPer) + h * 32 + lane;
                bool hit = false;
                if (e < bs) {
                    const float4* r = s_rec + kPairRec * e;
                    const float4 p0 = r[0], p1 = r[1], p2 = r[2];
                    const float4 hh4 = r[4];
                    hit = pair_cull_hit(p0.x, p0.z, -p1.x, -p1.z, -p2.x, r[3].w, hh4.x, hh4.y, TX0, TX1, TY0, TY1);
                }
                const unsigned int word = __ballot_sync(0xffffffffu, hit);  // bit l <-> entry chunk base + l
                if (lane == 0) s_tmask[warp * kPer + h] = word;
            }
            __syncthreads();
        }
        for (int c0 = 0; c0 < bs; c0 += 32) {
            if (__all_sync(0xffffffffu, !(npx0 > -INFINITY) && !(npx1 > -INFINITY))) break;
            unsigned int tword = 0xffffffffu;
            if (long_tile) {
                tword = s_tmask[c0 >> 5];
                if (tword == 0u) continue;  // nothing of this chunk reaches the tile
            }
            unsigned int mask;
            bool special = false;  // this lane's Gaussian needs the full alpha test (see staging)
            if (kCull) {
                bool hit = false;
                const int gi = c0 + 31 - lane;  // earliest Gaussian = highest ballot bit
                if (gi < bs && ((tword >> (31 - lane)) & 1u)) {
                    const float4* r = s_rec + kPairRec * gi;
                    const float4 p0 = r[0], p1 = r[1], p2 = r[2];
                    const float tau = r[3].w;
                    const float4 hh4 = r[4];
                    const float2 hh = make_float2(hh4.x, hh4.y);
                    special = hh4.z != 0.0f;
                    hit = pair_cull_hit(p0.x, p0.z, -p1.x, -p1.z, -p2.x, tau, hh.x, hh.y, X0, X1, Y0, Y1);
                }
                mask = __ballot_sync(0xffffffffu, hit);
                special = special && hit;
            } else {
                const int rem = bs - c0;
                mask = rem >= 32 ? 0xffffffffu : ~(0xffffffffu >> rem);
                special = true;
            }
            const bool any_special = __any_sync(0xffffffffu, special);
            const float4* rec_hi = s_rec + kPairRec * (c0 + 31);
            // two copies of the walk: chunks whose survivors are all "plain" (the common case) run without the
            // sigma < 0 test and without the 0.999 clamp (4 of 47 instructions)
            auto walk = [&](auto plain_tag) {
                constexpr bool kPlain = decltype(plain_tag)::value;
                while (mask) {
                    unsigned int b_hi;
                    asm("bfind.u32 %0, %1;" : "=r"(b_hi) : "r"(mask));
                    mask ^= bit_one << b_hi;
                    const float4* r = rec_hi - kPairRec * (int)b_hi;
                    const float4 p0 = r[0], p1 = r[1], p2 = r[2];
                    const f32x2 dx = add2(pk2(p0.x, p0.y), pk2(npx0, npx1));
                    const f32x2 dy = add2(pk2(p0.z, p0.w), npy);
                    const f32x2 nbdy = mul2(pk2(p1.z, p1.w), dy);
                    const f32x2 ncdy = mul2(pk2(p2.x, p2.y), dy);
                    const f32x2 L2 = pk2(p2.z, p2.w);
                    const f32x2 lmc = fma2(ncdy, dy, L2);
                    const f32x2 t = fma2(pk2(p1.x, p1.y), dx, nbdy);
                    const f32x2 pw = fma2(t, dx, lmc);
                    float pw0, pw1;
                    upk2(pw, pw0, pw1);
                    bool pass0 = pw0 >= kLog2AlphaThreshold, pass1 = pw1 >= kLog2AlphaThreshold;
                    float a0 = ex2_approx(pw0), a1 = ex2_approx(pw1);
                    if (!kPlain) {
                        const float Lt = r[4].w;  // sigma >= 0  <=>  power <= L (never fails for plain Gaussians)
                        pass0 = pass0 && (pw0 <= Lt);
                        pass1 = pass1 && (pw1 <= Lt);
                        a0 = fminf(0.999f, a0);
                        a1 = fminf(0.999f, a1);
                    }
                    const f32x2 a2 = pk2(a0, a1), T2 = pk2(T0, T1);
                    const f32x2 nT2 = mul2(T2, fma2(a2, mone2, one2));
                    const f32x2 vis2 = mul2(a2, T2);
                    float nT0, nT1, vis0, vis1;
                    upk2(nT2, nT0, nT1);
                    upk2(vis2, vis0, vis1);
                    const bool live0 = nT0 > 1e-4f, live1 = nT1 > 1e-4f;
                    const float4 c = r[3];
                    vis0 = (pass0 && live0) ? vis0 : 0.0f;
                    vis1 = (pass1 && live1) ? vis1 : 0.0f;
                    T0 = (pass0 && live0) ? nT0 : T0;
                    T1 = (pass1 && live1) ? nT1 : T1;
                    npx0 = (pass0 && !live0) ? -INFINITY : npx0;
                    npx1 = (pass1 && !live1) ? -INFINITY : npx1;
                    ar0 = fmaf(c.x, vis0, ar0); ag0 = fmaf(c.y, vis0, ag0); ab0 = fmaf(c.z, vis0, ab0);
                    ar1 = fmaf(c.x, vis1, ar1); ag1 = fmaf(c.y, vis1, ag1); ab1 = fmaf(c.z, vis1, ab1);
                }
            };
            if (any_special) walk(std::false_type{});
            else walk(std::true_type{});
        }
        if (kMbar) {
            // a warp that is done says so BEFORE it releases the stage: whoever sees the release also sees the count
            if (!warp_done && __all_sync(0xffffffffu, !(npx0 > -INFINITY) && !(npx1 > -INFINITY))) {
                warp_done = true;
                if (lane == 0) atomicAdd(&s_done_warps, 1);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                mbar_arrive(&s_empty[batch % kMbarStages]);  // this warp has consumed the stage
            }
        }
    }

    if (kRec) cp_async_wait_all();  // nothing may still be landing in shared memory when it is reused below
    const float bgs = (m_dev != nullptr && *m_dev == 0ull) ? 0.0f : 1.0f;
    const float bgr = bgs * __ldg(background), bgg = bgs * __ldg(background + 1), bgb = bgs * __ldg(background + 2);
    const float o0r = fmaf(T0, bgr, ar0), o0g = fmaf(T0, bgg, ag0), o0b = fmaf(T0, bgb, ab0);
    const float o1r = fmaf(T1, bgr, ar1), o1g = fmaf(T1, bgg, ag1), o1b = fmaf(T1, bgb, ab1);
    const bool full_tile = (tile_x * kFastTile + kFastTile <= W) && (tile_y * kFastTile + kFastTile <= H);
    if (vec_store && full_tile) {
        __syncthreads();  // staging buffers are dead from here on
        float* s_out = reinterpret_cast<float*>(s_g);  // 16 rows x 48 floats = 3 KB
        const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 8 + (lane >> 3);
        float* d0 = s_out + (ly * kFastTile + lx) * 3;
        float* d1 = d0 + 4 * kFastTile * 3;
        d0[0] = o0r; d0[1] = o0g; d0[2] = o0b;
        d1[0] = o1r; d1[1] = o1g; d1[2] = o1b;
        __syncthreads();
        for (int v = tid; v < 16 * 12; v += kPairThreads) {
            const int row = v / 12, c4 = v % 12;
            const float4 val = reinterpret_cast<const float4*>(s_out)[row * 12 + c4];
            const int64_t off = ((int64_t)(tile_y * kFastTile + row) * W + tile_x * kFastTile) * 3;
            reinterpret_cast<float4*>(image + off)[c4] = val;
            // fused band exchange: the same 128-bit store into every peer's image (NVLink posted writes)
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q)  // constant indices: the pointer array stays in the parameter bank
                if (q < peers.n) reinterpret_cast<float4*>(peers.p[q] + off)[c4] = val;
        }
    } else {
#pragma unroll
        for (int q = -1; q < kMaxPeers; ++q) {
            if (q >= peers.n) break;
            float* img = q < 0 ? image : peers.p[q < 0 ? 0 : q];
            if (in0) { float* dst = img + ((int64_t)i0 * W + j) * 3; dst[0] = o0r; dst[1] = o0g; dst[2] = o0b; }
            if (in1) { float* dst = img + ((int64_t)i1 * W + j) * 3; dst[0] = o1r; dst[1] = o1g; dst[2] = o1b; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// warp kernel (tile 16, RGB, needs a 48 B/Gaussian record workspace): the default fast path.
// Same packed-pair arithmetic as raster_pair_kernel, but the four warps of a tile never meet:
//   * raster_prep_kernel turns every Gaussian ONCE per frame into a 48-byte raster record
//     {mx, my, -A, -B | -C, L, hy, hx | r, g, b, tau} (log2-folded conic, culling constants),
//     instead of once per (tile, Gaussian) inside the rasterizer;
//   * each warp walks the tile's list by itself in chunks of 32: the records of chunk k+1 are gathered
//     by sorted id with cp.async (16 B x 3 per lane) into the warp's own staging slot while chunk k is
//     composited; lane l tests Gaussian l of the chunk against the warp's 8x8 block, the survivors are
//     compacted (ballot prefix) into the warp's survivor slot in the duplicated {v, v} layout and walked
//     front to back.  No __syncthreads in the loop: a warp whose 64 pixels are saturated just leaves.
// ------------------------------------------------------------------------------------------
constexpr int kWarpKStages = 3;  // chunk k is consumed while k+1 and k+2 are in flight

__global__ void __launch_bounds__(256)
raster_prep_kernel(const int64_t N, const float* __restrict__ means2d, const float* __restrict__ conics,
                   const float* __restrict__ colors, const float* __restrict__ opacities,
                   float4* __restrict__ rec) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const float mx = __ldg(means2d + 2 * g), my = __ldg(means2d + 2 * g + 1);
    const float ca = __ldg(conics + 3 * g), cb = __ldg(conics + 3 * g + 1), cc = __ldg(conics + 3 * g + 2);
    const float op = __ldg(opacities + g);
    const float A = 0.5f * kLog2e * ca, B = kLog2e * cb, C = 0.5f * kLog2e * cc;
    // accurate log2 keeps alpha = 2^(L - q) within a few ulp of o*exp(-sigma)
    const float L = (op > 0.0f) ? __log2f(op) : -INFINITY;  // same arithmetic as the pair kernel's staging
    const bool pd = (A > 0.0f) && (C > 0.0f) && (4.0f * A * C - B * B > 0.0f);
    float tau = pd ? (L - kLog2AlphaThreshold) : INFINITY;
    if (!(op == op)) tau = INFINITY;  // NaN opacity: evaluate, never cull
    rec[3 * g] = make_float4(mx, my, -A, -B);
    rec[3 * g + 1] = make_float4(-C, L, pd ? __fdividef(-B, 2.0f * C) : 0.0f, pd ? __fdividef(-B, 2.0f * A) : 0.0f);
    rec[3 * g + 2] = make_float4(__ldg(colors + 3 * g), __ldg(colors + 3 * g + 1), __ldg(colors + 3 * g + 2), tau);
}

template <bool kCull>
__global__ void __launch_bounds__(32, 32)
raster_warp_kernel(const int64_t N, const float4* __restrict__ rec, const float* __restrict__ background,
                   const int32_t* __restrict__ tile_ranges, const int32_t* __restrict__ tile_order,
                   const int first_tile, const int32_t* __restrict__ sorted_ids, const int W, const int H,
                   const int tiles_w, float* __restrict__ image, const int vec_store,
                   const unsigned long long* __restrict__ m_dev) {
    __shared__ float4 s_stage1[kWarpKStages][32 * 3];  // gathered records, a ring
    __shared__ float4 s_surv1[32 * kPairRec];          // compacted survivors, {v, v} layout
    float4 (*const s_stage)[32 * 3] = s_stage1;

    // one warp per CTA: the four 8x8 blocks of a tile are scheduled (and retire) independently
    const int lane = threadIdx.x;
    const int warp = (int)blockIdx.x & 3;
    const int tile_slot = (int)blockIdx.x >> 2;
    const int tile = tile_order ? __ldg(tile_order + tile_slot) : first_tile + tile_slot;
    const int tile_y = tile / tiles_w, tile_x = tile - tile_y * tiles_w;
    const int bx = tile_x * kFastTile + (warp & 1) * 8;
    const int by = tile_y * kFastTile + (warp >> 1) * 8;
    const int j = bx + (lane & 7);
    const int i0 = by + (lane >> 3), i1 = i0 + 4;
    const bool in0 = (i0 < H) && (j < W), in1 = (i1 < H) && (j < W);
    float npx0 = in0 ? -((float)j + 0.5f) : -INFINITY;  // negated x per pixel; -inf = finished
    float npx1 = in1 ? -((float)j + 0.5f) : -INFINITY;
    const f32x2 npy = pk2(-((float)i0 + 0.5f), -((float)i1 + 0.5f));
    const float X0 = (float)bx + 0.5f, X1 = (float)bx + 7.5f;
    const float Y0 = (float)by + 0.5f, Y1 = (float)by + 7.5f;

    const int32_t r0 = tile_ranges[2 * tile], r1 = tile_ranges[2 * tile + 1];
    float T0 = 1.0f, T1 = 1.0f;
    float ar0 = 0.f, ag0 = 0.f, ab0 = 0.f, ar1 = 0.f, ag1 = 0.f, ab1 = 0.f;
    const f32x2 one2 = pk2(1.0f, 1.0f), mone2 = pk2(-1.0f, -1.0f);
    const unsigned int lt = (1u << lane) - 1u;
    float4* const surv = s_surv1;

    // gather of one chunk: lane l fetches the record of list entry (chunk base + l) into stage[buf]
    auto gather = [&](int buf, int32_t id) {
        float4* dst = &s_stage[buf][lane * 3];
        if (id >= 0 && (int64_t)id < N) {
            const float4* src = rec + 3 * (int64_t)id;
            cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
        } else {
            // past the end of the list / invalid id (rasterization.mojo:109 guard): can never hit
            dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[1] = make_float4(0.f, -INFINITY, 0.f, 0.f);
            dst[2] = make_float4(0.f, 0.f, 0.f, -INFINITY);
        }
        cp_async_commit();
    };

    // Software pipeline: list ids are fetched 3 chunks ahead of their gather, records 2 chunks ahead of their
    // use, so that a run of chunks without survivors (the culled Gaussians that the torch binning rules pile
    // up in the corner tiles: ~10 k entries, 300 chunks per warp) advances at the speed of the test, not of
    // an L2 round trip per chunk.
    auto load_id = [&](int32_t at) { return (at < r1) ? __ldg(sorted_ids + at) : -1; };
    int32_t id_a = -1, id_b = -1, id_c = -1;  // ids of this lane's entries in chunks k+2, k+3, k+4
    if (r0 < r1) {
        const int32_t id0 = load_id(r0 + lane), id1 = load_id(r0 + 32 + lane);
        id_a = load_id(r0 + 64 + lane); id_b = load_id(r0 + 96 + lane); id_c = load_id(r0 + 128 + lane);
        gather(0, id0);
        gather(1, id1);
    }
    int buf = 0;
    for (int32_t base = r0; base < r1; base += 32, buf = (buf + 1 == kWarpKStages) ? 0 : buf + 1) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // chunk k has landed (k+1 may still fly)
        __syncwarp();
        const float4* st = &s_stage[buf][lane * 3];
        const float4 p0 = st[0], p1 = st[1], p2 = st[2];
        // chunk k+2 goes into the slot chunk k-1 was read from (all lanes passed the vote that ended k-1)
        gather((buf + 2 >= kWarpKStages) ? buf + 2 - kWarpKStages : buf + 2, id_a);
        id_a = id_b; id_b = id_c;
        id_c = load_id(base + 160 + lane);

        bool hit;
        if (kCull) {
            const float A = -p0.z, B = -p0.w, C = -p1.x, tau = p2.w;
            const float u0 = p0.x - X1, u1 = p0.x - X0;
            const float v0 = p0.y - Y1, v1 = p0.y - Y0;
            const bool zu = (u0 <= 0.0f) && (u1 >= 0.0f);
            const bool zv = (v0 <= 0.0f) && (v1 >= 0.0f);
            // min of q over the block = min over the (<= 2) edges facing the mean (see raster_fast_kernel)
            float qmin = 0.0f;
            if (!(zu && zv)) {
                float q1 = INFINITY, q2 = INFINITY;
                if (!zu) {
                    const float ue = (u0 > 0.0f) ? u0 : u1;
                    const float vstar = p1.z * ue;
                    const float dv = vstar - fminf(fmaxf(vstar, v0), v1);
                    q1 = fmaf(fmaf(0.5f * B, p1.z, A) * ue, ue, C * dv * dv);
                }
                if (!zv) {
                    const float ve = (v0 > 0.0f) ? v0 : v1;
                    const float ustar = p1.w * ve;
                    const float du = ustar - fminf(fmaxf(ustar, u0), u1);
                    q2 = fmaf(fmaf(0.5f * B, p1.w, C) * ve, ve, A * du * du);
                }
                qmin = fminf(q1, q2);
            }
            const float um = fmaxf(fabsf(u0), fabsf(u1)), vm = fmaxf(fabsf(v0), fabsf(v1));
            const float slack = 4e-6f * (A * um * um + C * vm * vm) + 1e-3f;
            hit = !(qmin > tau + slack);  // NaN-safe: anything odd counts as a hit
        } else {
            hit = !(p2.w == -INFINITY);  // every real Gaussian
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            float4* d = surv + kPairRec * __popc(mask & lt);
            d[0] = make_float4(p0.x, p0.x, p0.y, p0.y);
            d[1] = make_float4(p0.z, p0.z, p0.w, p0.w);
            d[2] = make_float4(p1.x, p1.x, p1.y, p1.y);
            d[3] = p2;
        }
        __syncwarp();
        const float4* r = surv;
        const float4* const r_end = surv + kPairRec * __popc(mask);
        for (; r != r_end; r += kPairRec) {
            const float4 q0 = r[0], q1 = r[1], q2 = r[2];
            const f32x2 dx = add2(pk2(q0.x, q0.y), pk2(npx0, npx1));
            const f32x2 dy = add2(pk2(q0.z, q0.w), npy);
            const f32x2 nbdy = mul2(pk2(q1.z, q1.w), dy);
            const f32x2 ncdy = mul2(pk2(q2.x, q2.y), dy);
            const f32x2 lmc = fma2(ncdy, dy, pk2(q2.z, q2.w));
            const f32x2 t = fma2(pk2(q1.x, q1.y), dx, nbdy);
            const f32x2 pw = fma2(t, dx, lmc);
            float pw0, pw1;
            upk2(pw, pw0, pw1);
            const float L = q2.z;
            const bool pass0 = (pw0 <= L) && (pw0 >= kLog2AlphaThreshold);
            const bool pass1 = (pw1 <= L) && (pw1 >= kLog2AlphaThreshold);
            const float a0 = fminf(0.999f, ex2_approx(pw0)), a1 = fminf(0.999f, ex2_approx(pw1));
            const f32x2 a2 = pk2(a0, a1), T2 = pk2(T0, T1);
            const f32x2 nT2 = mul2(T2, fma2(a2, mone2, one2));
            const f32x2 vis2 = mul2(a2, T2);
            float nT0, nT1, vis0, vis1;
            upk2(nT2, nT0, nT1);
            upk2(vis2, vis0, vis1);
            const bool live0 = nT0 > 1e-4f, live1 = nT1 > 1e-4f;
            const float4 c = r[3];
            vis0 = (pass0 && live0) ? vis0 : 0.0f;
            vis1 = (pass1 && live1) ? vis1 : 0.0f;
            T0 = (pass0 && live0) ? nT0 : T0;
            T1 = (pass1 && live1) ? nT1 : T1;
            // saturated: this Gaussian is not added (rasterization.mojo:146-150) and the pixel retires
            npx0 = (pass0 && !live0) ? -INFINITY : npx0;
            npx1 = (pass1 && !live1) ? -INFINITY : npx1;
            ar0 = fmaf(c.x, vis0, ar0); ag0 = fmaf(c.y, vis0, ag0); ab0 = fmaf(c.z, vis0, ab0);
            ar1 = fmaf(c.x, vis1, ar1); ag1 = fmaf(c.y, vis1, ag1); ab1 = fmaf(c.z, vis1, ab1);
        }
        if (__all_sync(0xffffffffu, !(npx0 > -INFINITY) && !(npx1 > -INFINITY))) break;  // also orders surv reuse
    }
    cp_async_wait_all();  // nothing may still be landing in shared memory when it is reused below

    const float bgs = (m_dev != nullptr && *m_dev == 0ull) ? 0.0f : 1.0f;
    const float bgr = bgs * __ldg(background), bgg = bgs * __ldg(background + 1), bgb = bgs * __ldg(background + 2);
    const float o0r = fmaf(T0, bgr, ar0), o0g = fmaf(T0, bgg, ag0), o0b = fmaf(T0, bgb, ab0);
    const float o1r = fmaf(T1, bgr, ar1), o1g = fmaf(T1, bgg, ag1), o1b = fmaf(T1, bgb, ab1);
    const bool full_block = (bx + 8 <= W) && (by + 8 <= H);
    if (vec_store && full_block) {
        __syncwarp();  // the survivor slot is dead from here on
        float* s_out = reinterpret_cast<float*>(surv);  // 8 rows x 24 floats
        float* d0 = s_out + ((lane >> 3) * 8 + (lane & 7)) * 3;
        float* d1 = d0 + 4 * 8 * 3;
        d0[0] = o0r; d0[1] = o0g; d0[2] = o0b;
        d1[0] = o1r; d1[1] = o1g; d1[2] = o1b;
        __syncwarp();
        for (int v = lane; v < 8 * 6; v += 32) {  // 8 rows of 96 contiguous bytes
            const int row = v / 6, c4 = v % 6;
            const float4 val = reinterpret_cast<const float4*>(s_out)[row * 6 + c4];
            float* dst = image + ((int64_t)(by + row) * W + bx) * 3;
            reinterpret_cast<float4*>(dst)[c4] = val;
        }
    } else {
        if (in0) { float* dst = image + ((int64_t)i0 * W + j) * 3; dst[0] = o0r; dst[1] = o0g; dst[2] = o0b; }
        if (in1) { float* dst = image + ((int64_t)i1 * W + j) * 3; dst[0] = o1r; dst[1] = o1g; dst[2] = o1b; }
    }
}

// Tiles sorted by list length, longest first (counting sort on len/32 capped to 255 buckets; order
// inside a bucket is arbitrary and does not affect results).  One CTA, n_tiles is small.
__global__ void __launch_bounds__(1024)
tile_order_kernel(const int first_tile, const int n_tiles, const int32_t* __restrict__ tile_ranges_all,
                  int32_t* __restrict__ order) {
    const int32_t* tile_ranges = tile_ranges_all + 2 * (int64_t)first_tile;
    __shared__ int s_cnt[256];
    __shared__ int s_base[256];
    const int tid = threadIdx.x;
    if (tid < 256) s_cnt[tid] = 0;
    __syncthreads();
    for (int t = tid; t < n_tiles; t += blockDim.x) {
        const int len = tile_ranges[2 * t + 1] - tile_ranges[2 * t];
        const int b = 255 - min(255, (len + 31) >> 5);  // bucket 0 = longest
        atomicAdd(&s_cnt[b], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < 256; ++b) { s_base[b] = acc; acc += s_cnt[b]; }
    }
    __syncthreads();
    for (int t = tid; t < n_tiles; t += blockDim.x) {
        const int len = tile_ranges[2 * t + 1] - tile_ranges[2 * t];
        const int b = 255 - min(255, (len + 31) >> 5);
        order[atomicAdd(&s_base[b], 1)] = first_tile + t;
    }
}

}  // namespace bsplat

using namespace bsplat;

template <int CH>
static int launch_faithful(int64_t N, int cdim, int c0, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* background,
                           const int32_t* tile_ranges, const int32_t* sorted_ids, int W, int H, int ts,
                           int row_begin, int row_end, float* image, unsigned long long* stats,
                           const unsigned long long* m_dev, cudaStream_t stream) {
    const int tiles_w = (W + ts - 1) / ts;
    const dim3 grid(tiles_w, row_end - row_begin), block(ts, ts);
    const size_t smem = (size_t)ts * ts * (6 + CH) * sizeof(float);
    raster_faithful_kernel<CH><<<grid, block, smem, stream>>>(N, cdim, c0, means2d, conics, colors, opacities,
                                                              background, tile_ranges, sorted_ids, W, H, ts,
                                                              tiles_w, row_begin, image, stats, m_dev);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

namespace bsplat {
// shared with capi.cu.  mode 0 = fast (pair kernel; with a record workspace: records + cp.async staging),
// 2 = the same without sub-tile culling (exactness A/B), 3 = warp kernel (independent warps, needs the record
// workspace; pair kernel without it), 4 = one pixel per lane (first fast kernel), 1 = faithful.  On config 3 the three fast kernels are within 3 % of each other
// (0.29-0.30 ms): 264 M / 211 M / 187 M warp instructions, but the lighter ones issue less densely
// (profiles/r01_raster_*): the pair kernel is the default.
size_t raster_workspace_bytes(int64_t N) { return (size_t)(N > 0 ? N : 1) * kPairRec * sizeof(float4); }

int rasterize_launch(int64_t N, int channels, const float* means2d, const float* conics, const float* colors,
                     const float* opacities, const float* background_dev,
                     const int32_t* tile_ranges, const int32_t* tile_order, const int32_t* sorted_ids, int W,
                     int H, int tile_size, int row_begin, int row_end, int mode, float* image,
                     unsigned long long* stats, const unsigned long long* m_dev, void* rec_ws,
                     cudaStream_t stream, const PeerImages* peers_in, const int32_t* rec_list,
                     const unsigned long long* rec_list_n) {
    PeerImages peers;
    peers.n = 0;
    if (peers_in) peers = *peers_in;
    if (W <= 0 || H <= 0 || tile_size <= 0 || tile_size > 32 || channels <= 0 || !background_dev)
        return BSPLAT_E_ARG;
    const int tiles_w = (W + tile_size - 1) / tile_size, tiles_h = (H + tile_size - 1) / tile_size;
    if (tiles_h > 65535) return BSPLAT_E_ARG;
    if (row_begin < 0) row_begin = 0;
    if (row_end > tiles_h) row_end = tiles_h;
    if (row_end <= row_begin) return BSPLAT_OK;  // empty band
    if (peers.n > 0 && !(mode == BSPLAT_RASTER_FAST && tile_size == kFastTile && channels == 3 && stats == nullptr &&
                         (reinterpret_cast<uintptr_t>(means2d) & 7u) == 0))
        return BSPLAT_E_ARG;  // the fused exchange exists in the default 16x16 RGB kernel only
    const bool fast_ok = (mode == BSPLAT_RASTER_FAST || (mode >= 2 && mode <= 5)) && tile_size == kFastTile &&
                         channels == 3 && stats == nullptr &&
                         (reinterpret_cast<uintptr_t>(means2d) & 7u) == 0;
    if (fast_ok) {
        const float* bg = background_dev;
        const int vec = ((reinterpret_cast<uintptr_t>(image) & 15u) == 0 && (W % 4) == 0) ? 1 : 0;
        const unsigned grid = (unsigned)(tiles_w * (row_end - row_begin));
        const int first_tile = row_begin * tiles_w;
        const LoopConsts loop_consts = {1u, 1.0f, -1.0f};
        const bool have_rec = rec_ws != nullptr && N > 0 && (reinterpret_cast<uintptr_t>(rec_ws) & 15u) == 0;
        float4* recp = static_cast<float4*>(rec_ws);
        if (mode == 3 && have_rec) {
            raster_prep_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(N, means2d, conics, colors, opacities, recp);
            BSPLAT_LAUNCH_CHECK();
            raster_warp_kernel<true><<<grid * 4, 32, 0, stream>>>(N, recp, bg, tile_ranges, tile_order,
                                                                  first_tile, sorted_ids, W, H, tiles_w,
                                                                  image, vec, m_dev);
        } else if (mode == 4) {
            raster_fast_kernel<true><<<grid, kFastThreads, 0, stream>>>(N, means2d, conics, colors, opacities, bg,
                                                                        tile_ranges, tile_order, first_tile, sorted_ids,
                                                                        W, H, tiles_w, image, vec, m_dev);
        } else if (have_rec) {
            // default: records once per frame, then the cp.async-staged pair kernel
            raster_pair_prep_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(N, means2d, conics, colors,
                                                                                    opacities, recp, rec_list,
                                                                                    rec_list_n);
            BSPLAT_LAUNCH_CHECK();
            if (mode == 5)
                raster_pair_kernel<true, true, true><<<grid, kPairThreads, 0, stream>>>(
                    N, recp, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                    tiles_w, image, vec, m_dev, peers, loop_consts);
            else if (mode == 2)
                raster_pair_kernel<false, true><<<grid, kPairThreads, 0, stream>>>(
                    N, recp, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                    tiles_w, image, vec, m_dev, peers, loop_consts);
            else
                raster_pair_kernel<true, true><<<grid, kPairThreads, 0, stream>>>(
                    N, recp, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                    tiles_w, image, vec, m_dev, peers, loop_consts);
        } else if (mode == 2) {
            raster_pair_kernel<false, false><<<grid, kPairThreads, 0, stream>>>(
                N, nullptr, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                tiles_w, image, vec, m_dev, peers, loop_consts);
        } else {
            raster_pair_kernel<true, false><<<grid, kPairThreads, 0, stream>>>(
                N, nullptr, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                tiles_w, image, vec, m_dev, peers, loop_consts);
        }
        BSPLAT_LAUNCH_CHECK();
        return BSPLAT_OK;
    }
    // faithful path: channels in chunks of <= 4 (alpha is recomputed per chunk; RGB is one chunk)
    for (int c0 = 0; c0 < channels; c0 += 4) {
        const int ch = channels - c0 < 4 ? channels - c0 : 4;
        int rc;
        unsigned long long* st = (c0 == 0) ? stats : nullptr;
        switch (ch) {
            case 1: rc = launch_faithful<1>(N, channels, c0, means2d, conics, colors, opacities, background_dev, tile_ranges, sorted_ids, W, H, tile_size, row_begin, row_end, image, st, m_dev, stream); break;
            case 2: rc = launch_faithful<2>(N, channels, c0, means2d, conics, colors, opacities, background_dev, tile_ranges, sorted_ids, W, H, tile_size, row_begin, row_end, image, st, m_dev, stream); break;
            case 3: rc = launch_faithful<3>(N, channels, c0, means2d, conics, colors, opacities, background_dev, tile_ranges, sorted_ids, W, H, tile_size, row_begin, row_end, image, st, m_dev, stream); break;
            default: rc = launch_faithful<4>(N, channels, c0, means2d, conics, colors, opacities, background_dev, tile_ranges, sorted_ids, W, H, tile_size, row_begin, row_end, image, st, m_dev, stream); break;
        }
        if (rc != BSPLAT_OK) return rc;
    }
    return BSPLAT_OK;
}

int tile_order_launch(int first_tile, int n_tiles, const int32_t* tile_ranges, int32_t* order,
                      cudaStream_t stream) {
    if (n_tiles <= 0) return BSPLAT_OK;
    tile_order_kernel<<<1, 1024, 0, stream>>>(first_tile, n_tiles, tile_ranges, order);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}
}  // namespace bsplat
