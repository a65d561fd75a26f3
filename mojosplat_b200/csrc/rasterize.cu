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
//   raster_pair_kernel      16x16 tiles, RGB: packed FP32 pairs (two pixels per lane), log2-folded
//       conics (a pixel costs a handful of FFMA2/FADD2/FMUL2 + 1 MUFU.EX2), cp.async staging of
//       per-Gaussian records one batch ahead, per-warp culling of 32 staged Gaussians at a time against
//       the warp's 8x8 block with an exact conservative bound, warp / CTA early exit on saturation,
//       128-bit image stores (also into the peers' images in the fused row-band exchange), and a
//       multi-SM compaction pre-pass for very long lists.  Skipping is exact: a skipped Gaussian has
//       alpha < 1/255 on every pixel of the block, so the composited result is unchanged.
// No tensor cores: no stage of this path is a dense contraction.
#include <stdlib.h>
#include <string.h>

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
// pair kernel (tile 16, RGB): the fast path.
// sm_100 executes packed FP32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot, two lanes of work), and the
// rasterizer is bound by instruction issue, not by the FP32 pipe.  So each lane owns TWO pixels -- (x, y) and
// (x, y + 4) of the warp's 8x8 block -- and the whole quadratic, the transmittance update and the colour
// accumulation run as f32x2 instructions on register pairs; the per-Gaussian operands enter as scalar registers
// broadcast to both halves (`R.F32` operands), so records hold every value once.  4 warps (128 threads) per tile.
//   record t (48 B): s[3t] = {mx, my, -A, -B}  s[3t+1] = {-C, L, r, g}  s[3t+2] = {b, tau*, hy, hx}
//   with A = 0.5 a log2e, B = b log2e, C = 0.5 c log2e, L = log2(opacity), hy = -B/(2C), hx = -B/(2A) (edge
//   minimisers of the quadratic), tau = L - log2(1/255) (+inf: never cull, -inf: not a Gaussian; lowest bit = special)
//   power = L - (A dx^2 + B dx dy + C dy^2) as  fma2(fma2(-A, dx, -B dy), dx, fma2(-C dy, dy, L))
// alpha is zeroed when it fails the threshold, so T (1 - alpha) = T and alpha T = 0 need no selects; the walk only
// branches (rarely) when a pixel saturates.  A finished pixel carries -x = -inf (every later power is -inf or NaN
// and fails the alpha test by itself -- no flag in the inner loop).  Each warp first tests 32 staged Gaussians at once (one per lane) against its 8x8
// block with an exact conservative ellipse / rectangle bound and then only walks the survivors (warp ballot);
// a skipped Gaussian has alpha < 1/255 on all 64 pixels, so the composited result is unchanged.
// (Three earlier variants -- one pixel per lane, independent warps with per-warp staging, an mbarrier
// full/empty pipeline instead of the per-batch barrier -- were measured slower and removed: DESIGN.md section 9.)
// ------------------------------------------------------------------------------------------
constexpr int kFastTile = 16;
constexpr int kPairThreads = 128;
constexpr int kPairBatch = 256;   // staged entries per batch without records
#ifndef BSPLAT_RASTER_MINB
#define BSPLAT_RASTER_MINB 0  // 0: no minimum-blocks bound (ptxas picks 56 registers = 9 CTAs per SM by itself: 154 M instructions).  A/B on one box, pipeline frames/s: bound 8 / 9 / 10 (58 / 53 / 48 registers) 2 899 / 2 942 / 2 901; unbound vs 9: 2 971 vs 2 935
#endif
#ifndef BSPLAT_REC_BATCH
#define BSPLAT_REC_BATCH 128
#endif
#ifndef BSPLAT_REC_STAGES
#define BSPLAT_REC_STAGES 2
#endif
constexpr int kRecBatch = BSPLAT_REC_BATCH;    // with records: entries per batch (a multiple of 32) ...
constexpr int kRecStages = BSPLAT_REC_STAGES;  // ... in a ring of this many buffers: the gather runs kRecStages - 1 batches ahead
static_assert(kRecBatch % 32 == 0 && kRecStages >= 2 && kRecStages <= 4, "record staging ring");

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// packed FP32 pairs: the sm_100 intrinsics on float2 (register pairs; FFMA2 / FADD2 / FMUL2)
__device__ __forceinline__ float2 dup2(const float v) { return make_float2(v, v); }  // becomes an `R.F32` operand
__device__ __forceinline__ bool pair_finished(const float2 npx) {  // both halves are -inf
    return !(npx.x > -INFINITY) && !(npx.y > -INFINITY);
}

// Lists longer than kLongTile are mostly Gaussians that cannot touch the tile at all (under the torch binning rules
// every culled Gaussian is clamped into a border tile: ~10 k entries in each corner tile at config 3, ~350 k at
// config 5).  Walking such a list is a serial chain of thousands of batches inside ONE CTA.  Two remedies:
//   * len > kLongTile: the four warps of the tile share ONE test of every staged Gaussian against the whole 16x16 tile
//     (64 entries per warp instead of 256 each) before the per-warp 8x8 tests -- still one CTA, ~4x shorter chain;
//   * len > kPrepassMin (fused frames): the tile-level test, which is embarrassingly parallel, runs as a pre-pass over
//     all SMs (raster_long_compact_kernel) and leaves a compacted private copy of the list; the rasterizer walks only
//     the survivors.  This is what bounds a frame once it is split across GPUs (0.35 ms chains at config 5).
// sorted_ids / tile_ranges, the API-visible lists, are untouched.
constexpr int kLongTile = 2048;
constexpr int kPrepassMin = 16384;     // shorter lists: the in-kernel tile test hides behind the rest of the frame
constexpr int kLongChunk = 4096;       // entries per pre-pass work item
constexpr int kMaxLongChunks = 256;    // longer lists (> 1 M entries) keep the in-kernel tile test
constexpr int kMaxLongSlots = 256;     // at most this many long tiles are compacted (the longest ones)
constexpr int kLongThreads = 1024;     // 4 entries per thread: every gather of a chunk is in flight at once


__device__ __forceinline__ void pair_record(const int64_t g, const float* __restrict__ means2d,
                                            const float* __restrict__ conics, const float* __restrict__ colors,
                                            const float* __restrict__ opacities, float4& q0, float4& q1, float4& q2) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(means2d) + g);
    pair_record_from(m.x, m.y, __ldg(conics + 3 * g), __ldg(conics + 3 * g + 1), __ldg(conics + 3 * g + 2),
                     __ldg(opacities + g), __ldg(colors + 3 * g), __ldg(colors + 3 * g + 1), __ldg(colors + 3 * g + 2),
                     q0, q1, q2);
}

// Records of ALL Gaussians, once per frame (48 B each): with them the rasterizer's staging is a pure gather that
// cp.async can run one batch ahead, and the per-(tile, Gaussian) staging arithmetic disappears.  (Fused frames get
// the records from the projection kernel's epilogue instead; this kernel serves the stage-level entry point.)
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
    float4 q0, q1, q2;
    pair_record(g, means2d, conics, colors, opacities, q0, q1, q2);
    float4* d = rec + kPairRec * g;
    d[0] = q0; d[1] = q1; d[2] = q2;
}

// Scratch of the pre-pass (uint32 words): [0, kMaxLongSlots) survivors per compacted slot (slot = position in
// tile_order), then the per-chunk survivor counts: tiles own disjoint, tile-ordered list ranges, so
// (r0 / chunk) + tile + c never collides between tiles (floor(a) + floor(b) <= floor(a + b)).
__device__ __forceinline__ int64_t long_chunk_index(const int32_t r0, const int tile, const int c) {
    return (int64_t)kMaxLongSlots + (int64_t)(r0 / kLongChunk) + tile + c;
}

// Tile-level test of one chunk: ballots[k] = hits of this warp's 32 entries of round k; ids[k] = this thread's entry.
__device__ __forceinline__ void long_test_chunk(const int64_t N, const float4* __restrict__ rec,
                                                const int32_t* __restrict__ sorted_ids, const int32_t e0,
                                                const int32_t r1, const float TX0, const float TX1, const float TY0,
                                                const float TY1, int32_t (&ids)[kLongChunk / kLongThreads],
                                                unsigned int (&ballots)[kLongChunk / kLongThreads]) {
    constexpr int kRounds = kLongChunk / kLongThreads;
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        const int32_t e = e0 + k * kLongThreads + (int)threadIdx.x;
        ids[k] = (e < r1) ? __ldg(sorted_ids + e) : -1;
    }
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        bool hit = false;
        const int32_t g = ids[k];
        if (g >= 0 && (int64_t)g < N) {
            const float4* r = rec + kPairRec * (int64_t)g;
            const float4 p0 = __ldg(r), p2 = __ldg(r + 2);
            const float nC = __ldg(reinterpret_cast<const float*>(r + 1));
            hit = pair_cull_hit(p0.x, p0.y, -p0.z, -p0.w, -nC, p2.y, p2.z, p2.w, TX0, TX1, TY0, TY1);
        }
        ballots[k] = __ballot_sync(0xffffffffu, hit);
    }
}

// Pre-pass for very long lists (see kPrepassMin).  Work items = (long tile, chunk of kLongChunk entries), found by
// every CTA on its own from the head of tile_order (longest lists first, so the long tiles are a prefix of it).
// Phase 1 counts the survivors of every chunk, a grid-wide barrier follows (the grid is one CTA per SM: co-resident),
// phase 2 repeats the test and writes the survivors, in list order, at the chunk's offset inside the tile's compacted
// list surv[r0 ...); scratch[slot] = number of survivors of the tile.  `barrier` is a zeroed counter.
__global__ void __launch_bounds__(kLongThreads)
raster_long_compact_kernel(const int64_t N, const float4* __restrict__ rec, const int32_t* __restrict__ tile_ranges,
                           const int32_t* __restrict__ tile_order, const int n_order,
                           const int32_t* __restrict__ sorted_ids, const int tiles_w,
                           int32_t* __restrict__ surv, uint32_t* __restrict__ scratch, uint32_t* __restrict__ barrier) {
    pdl_wait();  // (programmatic dependent launch: nothing of the predecessor is read before this)
    __shared__ int s_pref[kMaxLongSlots + 1];   // exclusive prefix of chunk counts over the long tiles
    __shared__ int s_wsum[kLongThreads / 32];
    __shared__ int s_grp[kLongChunk / 32];      // survivors per (round, warp) group of a chunk
    __shared__ int s_red[kLongThreads / 32];
    static_assert(kMaxLongSlots <= kLongThreads && kMaxLongChunks <= kLongThreads, "one thread per slot / chunk");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kRounds = kLongChunk / kLongThreads;
    constexpr int kGroups = kRounds * (kLongThreads / 32);
    static_assert(kGroups % 32 == 0 && kGroups <= kLongThreads, "group prefix layout");
    // ---- work discovery: chunks of slot `tid` ----
    int chunks = 0;
    if (tid < n_order && tid < kMaxLongSlots) {
        const int tile = __ldg(tile_order + tid);
        const int len = __ldg(tile_ranges + 2 * tile + 1) - __ldg(tile_ranges + 2 * tile);
        const int nc = (len + kLongChunk - 1) / kLongChunk;
        if (len > kPrepassMin && nc <= kMaxLongChunks) chunks = nc;
    }
    int incl = chunks;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_wsum[w];
    if (tid < kMaxLongSlots) s_pref[tid] = wbase + incl - chunks;
    if (tid == kMaxLongSlots - 1) s_pref[kMaxLongSlots] = wbase + incl;
    __syncthreads();
    const int total = s_pref[kMaxLongSlots];
    if (total == 0) return;  // (the same decision in every CTA: nobody waits at the barrier below)

    for (int phase = 0; phase < 2; ++phase) {
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            // slot of this item: the largest s with s_pref[s] <= item (zero-chunk slots are skipped by construction)
            int lo = 0, hi = kMaxLongSlots;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_pref[mid] <= item) lo = mid; else hi = mid;
            }
            const int c = item - s_pref[lo];
            const int nc = s_pref[lo + 1] - s_pref[lo];
            const int tile = __ldg(tile_order + lo);
            const int32_t r0 = __ldg(tile_ranges + 2 * tile), r1 = __ldg(tile_ranges + 2 * tile + 1);
            const int tile_y = tile / tiles_w, tile_x = tile - tile_y * tiles_w;
            const float TX0 = (float)(tile_x * kFastTile) + 0.5f, TX1 = TX0 + 15.0f;
            const float TY0 = (float)(tile_y * kFastTile) + 0.5f, TY1 = TY0 + 15.0f;
            const int32_t e0 = r0 + c * kLongChunk;
            int32_t ids[kRounds];
            unsigned int ballots[kRounds];
            long_test_chunk(N, rec, sorted_ids, e0, r1, TX0, TX1, TY0, TY1, ids, ballots);
#pragma unroll
            for (int k = 0; k < kRounds; ++k)
                if (lane == 0) s_grp[k * (kLongThreads / 32) + warp] = __popc(ballots[k]);
            __syncthreads();
            // exclusive prefix over the (round, warp) groups: one warp scan per 32 groups, totals in s_wsum
            if (tid < kGroups) {
                const int v = s_grp[tid];
                int in2 = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, in2, d);
                    if (lane >= d) in2 += u;
                }
                s_grp[tid] = in2 - v;
                if (lane == 31) s_wsum[warp] = in2;
            }
            __syncthreads();
            int gtot[kGroups / 32], total_surv = 0;  // survivors of each run of 32 groups
#pragma unroll
            for (int w = 0; w < kGroups / 32; ++w) { gtot[w] = s_wsum[w]; total_surv += gtot[w]; }
            const int64_t cidx = long_chunk_index(r0, tile, 0);
            if (phase == 0) {
                if (tid == 0) scratch[cidx + c] = (uint32_t)total_surv;
            } else {
                // offset of this chunk inside the tile's compacted list: survivors of the chunks before it
                int before = (tid < c) ? (int)__ldcg(scratch + cidx + tid) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
                if (lane == 0) s_red[warp] = before;
                __syncthreads();
                int base0 = 0;
#pragma unroll
                for (int w = 0; w < kLongThreads / 32; ++w) base0 += s_red[w];
#pragma unroll
                for (int k = 0; k < kRounds; ++k) {
                    const int gi = k * (kLongThreads / 32) + warp;
                    int base = base0 + s_grp[gi];
#pragma unroll
                    for (int w = 0; w < kGroups / 32; ++w)
                        if (w < gi / 32) base += gtot[w];
                    const unsigned int bl = ballots[k];
                    BSPLAT_DASSERT(!((bl >> lane) & 1u) || r0 + base + (int)__popc(bl & ((1u << lane) - 1u)) < r1);
                    if ((bl >> lane) & 1u) surv[r0 + base + __popc(bl & ((1u << lane) - 1u))] = ids[k];
                }
                if (c == nc - 1 && tid == 0) scratch[lo] = (uint32_t)(base0 + total_surv);
            }
            __syncthreads();  // s_grp / s_wsum / s_red are reused by the next item
        }
        if (phase == 0) {
            // grid-wide barrier: every chunk count is visible before any offset is summed
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                atomicAdd(barrier, 1u);
                unsigned ns = 32;
                while (ld_relaxed_u32(barrier) < gridDim.x) {
                    __nanosleep(ns);
                    if (ns < 256) ns <<= 1;
                }
                __threadfence();
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// kRec: the Gaussians come as prepared records; batches of 128 are gathered by sorted id with cp.async into the two
// halves of the staging buffer, one batch ahead of the walk (ids two batches ahead), one barrier per batch.
// !kRec: workspace-free staging of 256 per batch from the raw arrays (same arithmetic, bit-identical image).
// kTrain (training-side forward, bsplat_rasterize_fwd_train_fast): the same walk also yields what the backward pass
// starts from -- per pixel the final transmittance and the list index of the last entry the pixel looked at (the
// entry in front of the one that saturated it, or the end of the list; entries that failed the alpha test in between
// fail it again in the backward pass).
struct TrainOut {
    float* final_T;
    int32_t* last_idx;
};

template <bool kCull, bool kRec, bool kTrain = false>
#if BSPLAT_RASTER_MINB > 0
__global__ void __launch_bounds__(kPairThreads, kTrain ? 8 : BSPLAT_RASTER_MINB)
#else
__global__ void __launch_bounds__(kPairThreads)
#endif
raster_pair_kernel(const int64_t N, const float4* __restrict__ rec, const float* __restrict__ means2d, const float* __restrict__ conics,
                   const float* __restrict__ colors, const float* __restrict__ opacities,
                   const float* __restrict__ background, const int32_t* __restrict__ tile_ranges,
                   const int32_t* __restrict__ tile_order, const int first_tile,
                   const int32_t* __restrict__ sorted_ids, const int W, const int H, const int tiles_w,
                   float* __restrict__ image, const int vec_store,
                   const unsigned long long* __restrict__ m_dev, const PeerImages peers,
                   const int32_t* __restrict__ surv, const uint32_t* __restrict__ chunk_cnt,
                   const TrainOut train) {
    pdl_wait();  // (programmatic dependent launch: nothing of the predecessor is read before this)
    // staging: with records a ring of kRecStages batches, without one batch; the output tile (16 x 48 floats) reuses it.
    // (Kept as small as possible: what shared memory does not take stays L1, which the record gathers live on.)
    constexpr int kStageRecs = kRec ? kRecStages * kRecBatch : kPairBatch;
    static_assert(kStageRecs * kPairRec * sizeof(float4) >= 16 * 48 * sizeof(float), "output tile fits the staging buffer");
    __shared__ float4 s_g[kStageRecs * kPairRec];
    __shared__ unsigned int s_tmask[(kRec ? kRecBatch : kPairBatch) / 32];  // long tiles without a pre-pass: survivors of the tile-level test
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t s_g_addr = (uint32_t)__cvta_generic_to_shared(s_g);
    const int tile = tile_order ? __ldg(tile_order + blockIdx.x) : first_tile + (int)blockIdx.x;
    const int tile_y = tile / tiles_w, tile_x = tile - tile_y * tiles_w;
    // warp -> 8x8 pixel block; lane -> column (lane & 7), rows (lane >> 3) and (lane >> 3) + 4
    const int bx = tile_x * kFastTile + (warp & 1) * 8;
    const int by = tile_y * kFastTile + (warp >> 1) * 8;
    const int j = bx + (lane & 7);
    const int i0 = by + (lane >> 3), i1 = i0 + 4;
    const bool in0 = (i0 < H) && (j < W), in1 = (i1 < H) && (j < W);
    const float2 npy = make_float2(-((float)i0 + 0.5f), -((float)i1 + 0.5f));
    const float X0 = (float)bx + 0.5f, X1 = (float)bx + 7.5f;
    const float Y0 = (float)by + 0.5f, Y1 = (float)by + 7.5f;

    const int32_t r0 = tile_ranges[2 * tile], r1 = tile_ranges[2 * tile + 1];
    // the list as the loop sees it: entries [v0, v1) of `ids` -- sorted_ids, or the compacted private copy the
    // pre-pass left for a very long list (same ids, same order, minus those that cannot reach the tile)
    const int32_t* __restrict__ ids = sorted_ids;
    int32_t v0 = r0, v1 = r1;
    bool compacted = false;
    if constexpr (kRec && kCull) {
        if (surv != nullptr && tile_order != nullptr && (int)blockIdx.x < kMaxLongSlots && r1 - r0 > kPrepassMin &&
            (r1 - r0 + kLongChunk - 1) / kLongChunk <= kMaxLongChunks) {
            compacted = true;
            ids = surv;
            v1 = r0 + (int32_t)__ldg(chunk_cnt + blockIdx.x);
            BSPLAT_DASSERT(v1 >= r0 && v1 <= r1);
        }
    }
    // long list without a pre-pass (stage-level entry point, or a list beyond the pre-pass limits): the four warps
    // share ONE test of every staged Gaussian against the whole tile (64 entries per warp instead of 256 each)
    const bool long_tile = kCull && !compacted && (r1 - r0 > kLongTile);
    const float TX0 = (float)(tile_x * kFastTile) + 0.5f, TX1 = TX0 + 15.0f;
    const float TY0 = (float)(tile_y * kFastTile) + 0.5f, TY1 = TY0 + 15.0f;
    float2 T2 = make_float2(1.0f, 1.0f);
    float2 acc_r = make_float2(0.f, 0.f), acc_g = acc_r, acc_b = acc_r;  // {pixel 0, pixel 1} per channel
    // negated x per pixel; -inf = finished
    float2 npx2 = make_float2(in0 ? -((float)j + 0.5f) : -INFINITY, in1 ? -((float)j + 0.5f) : -INFINITY);
    const float2 one2 = dup2(1.0f);
    int32_t stop0 = r1, stop1 = r1;  // kTrain: list position of the entry that saturated the pixel (r1: none did)

    constexpr int kBatch = kRec ? kRecBatch : kPairBatch;
    constexpr int kPer = (kBatch + kPairThreads - 1) / kPairThreads;  // staged entries per thread and batch
    struct IdSet { int32_t v[kPer]; };
    auto load_ids = [&](int32_t at) {  // this thread's entries of the batch starting at list position `at`
        IdSet o;
#pragma unroll
        for (int h = 0; h < kPer; ++h) {
            const int32_t e = at + h * kPairThreads + tid;
            o.v[h] = (e < v1 && h * kPairThreads + tid < kBatch) ? __ldg(ids + e) : -1;
        }
        return o;
    };
    auto gather = [&](int half, const IdSet& id) {  // kRec: this thread's entries of a batch -> record buffer `half`
#pragma unroll
        for (int h = 0; h < kPer; ++h) {
            if (h * kPairThreads + tid >= kBatch) break;  // (batches smaller than the CTA: the first warps gather)
            float4* dst = s_g + (half * kBatch + h * kPairThreads + tid) * kPairRec;
            BSPLAT_DASSERT(half >= 0 && (half * kBatch + h * kPairThreads + tid + 1) * kPairRec <= kStageRecs * kPairRec);
            if (id.v[h] >= 0 && (int64_t)id.v[h] < N) {
                const float4* src = rec + kPairRec * (int64_t)id.v[h];
#pragma unroll
                for (int q = 0; q < kPairRec; ++q) cp_async16(dst + q, src + q);
            } else {  // past the end of the list / invalid id (rasterization.mojo:109 guard): can never hit
                pair_record_none(dst[0], dst[1], dst[2]);
            }
        }
        cp_async_commit();
    };
    IdSet id_next;
#pragma unroll
    for (int h = 0; h < kPer; ++h) id_next.v[h] = -1;
    if (kRec && v0 < v1) {  // prologue: the first kRecStages - 1 batches are in flight, the ids of the next one loaded
#pragma unroll
        for (int st = 0; st < kRecStages - 1; ++st) gather(st, load_ids(v0 + st * kBatch));
        id_next = load_ids(v0 + (kRecStages - 1) * kBatch);
    }
    int batch = 0;
    for (int32_t b0 = v0; b0 < v1; b0 += kBatch, ++batch) {
        const bool fin = pair_finished(npx2);
        const float4* s_rec = s_g;
        uint32_t a_rec = s_g_addr;  // the same as a 32-bit shared address (see lds_f4)
        if (kRec) {
            cp_async_wait_group<kRecStages - 2>();  // this thread's part of batch `batch` has landed ...
            if (__syncthreads_count(fin) >= kPairThreads) break;  // ... everyone's has; batch - 1 is fully consumed
            // the buffer batch - 1 used is refilled kRecStages - 1 batches ahead of the walk (an empty group past the
            // end of the list keeps the group count in step)
            if (b0 + (kRecStages - 1) * kBatch < v1) gather((batch + kRecStages - 1) % kRecStages, id_next);
            else cp_async_commit();
            id_next = load_ids(b0 + kRecStages * kBatch);
            s_rec = s_g + (batch % kRecStages) * kBatch * kPairRec;
            a_rec = s_g_addr + (uint32_t)((batch % kRecStages) * kBatch * kPairRec * sizeof(float4));
        } else {
            if (__syncthreads_count(fin) >= kPairThreads) break;
#pragma unroll
            for (int h = 0; h < kPer; ++h) {
                const int t = tid + h * kPairThreads;
                const int32_t idx = b0 + t;
                if (idx < v1) {
                    const int32_t g = __ldg(ids + idx);
                    float4 q0, q1, q2;
                    if (g >= 0 && (int64_t)g < N) {
                        pair_record(g, means2d, conics, colors, opacities, q0, q1, q2);
                    } else {
                        pair_record_none(q0, q1, q2);
                    }
                    float4* dst = s_g + kPairRec * t;
                    dst[0] = q0; dst[1] = q1; dst[2] = q2;
                }
            }
            __syncthreads();
        }

        const int bs = min(kBatch, (int)(v1 - b0));
        if (long_tile) {
#pragma unroll
            for (int h = 0; h < kPer; ++h) {
                const int e = warp * (32 * kPer) + h * 32 + lane;
                bool hit = false;
                if (e < bs) hit = pair_record_hit(s_rec + kPairRec * e, TX0, TX1, TY0, TY1, nullptr);
                const unsigned int word = __ballot_sync(0xffffffffu, hit);  // bit l <-> entry chunk base + l
                if (lane == 0) s_tmask[warp * kPer + h] = word;
            }
            __syncthreads();
        }
        for (int c0 = 0; c0 < bs; c0 += 32) {
            if (__all_sync(0xffffffffu, pair_finished(npx2))) break;
            unsigned int tword = 0xffffffffu;
            if (long_tile) {
                tword = s_tmask[c0 >> 5];
                if (tword == 0u) continue;  // nothing of this chunk reaches the tile
            }
            unsigned int mask;
            bool special = false;  // this lane's Gaussian needs the full alpha test (see pair_record_from)
            if (kCull) {
                bool hit = false;
                // lane l tests Gaussian c0 + 31 - l: the earliest Gaussian is the HIGHEST ballot bit, so the
                // walk below needs a single FLO (bfind) per survivor instead of BREV + FLO
                const int gi = c0 + 31 - lane;
                if (gi < bs && ((tword >> (31 - lane)) & 1u))
                    hit = pair_record_hit_at(a_rec + (uint32_t)(gi * (int)(kPairRec * sizeof(float4))), X0, X1, Y0, Y1, &special);
                mask = __ballot_sync(0xffffffffu, hit);
                special = special && hit;
            } else {
                const int rem = bs - c0;
                mask = rem >= 32 ? 0xffffffffu : ~(0xffffffffu >> rem);  // Gaussian c0+k <-> bit 31-k
                special = true;
            }
            const bool any_special = __any_sync(0xffffffffu, special);
            // record of ballot bit 0; bit b is 3 b float4 (48 bytes) earlier
            const uint32_t a_hi = a_rec + (uint32_t)((c0 + 31) * (int)(kPairRec * sizeof(float4)));
            // two copies of the walk: chunks whose survivors are all "plain" (the common case) run without the
            // sigma < 0 test and without the 0.999 clamp
            // alpha of one Gaussian on this lane's two pixels; a failed alpha test gives alpha = 0 (then T (1 - 0) = T and
            // 0 T = 0 exactly: the compositing below needs no select for it)
            auto alpha_of = [&](auto plain_tag, const float tau, const float4 p0, const float4 p1) -> float2 {
                constexpr bool kPlain = decltype(plain_tag)::value;
                const float2 dx = __fadd2_rn(dup2(p0.x), npx2);
                const float2 dy = __fadd2_rn(dup2(p0.y), npy);
                const float2 nbdy = __fmul2_rn(dup2(p0.w), dy);
                const float2 ncdy = __fmul2_rn(dup2(p1.x), dy);
                const float2 lmc = __ffma2_rn(ncdy, dy, dup2(p1.y));
                const float2 t = __ffma2_rn(dup2(p0.z), dx, nbdy);
                const float2 pw = __ffma2_rn(t, dx, lmc);
                bool pass0 = pw.x >= kLog2AlphaThreshold, pass1 = pw.y >= kLog2AlphaThreshold;
                if (!kPlain) {
                    // sigma >= 0  <=>  power <= L: tested for special Gaussians only (plain ones pass by
                    // construction and must be treated exactly as in the plain walk)
                    const float Lt = (__float_as_uint(tau) & 1u) ? p1.y : INFINITY;
                    pass0 = pass0 && (pw.x <= Lt);
                    pass1 = pass1 && (pw.y <= Lt);
                }
                float2 a2 = make_float2(0.0f, 0.0f);
                if (pass0) a2.x = kPlain ? ex2_approx(pw.x) : fminf(0.999f, ex2_approx(pw.x));
                if (pass1) a2.y = kPlain ? ex2_approx(pw.y) : fminf(0.999f, ex2_approx(pw.y));
                return a2;
            };
            // compositing of one Gaussian.  Saturation (T (1 - alpha) <= 1e-4, once per pixel): the Gaussian is not
            // added (rasterization.mojo:146-150) and the pixel retires with its T unchanged -- alpha := 0, -x := -inf.
            // Branch-free: a lane that left the walk for a rare path would make its warp walk the chunk twice.
            auto composite = [&](float2 a2, const float cr, const float cg, const float cb, const int32_t cur) {
                // T (1 - alpha) as T - alpha T: alpha T is needed anyway (one packed instruction instead of two)
                float2 vis2 = __fmul2_rn(a2, T2);
                const float2 nT = __fadd2_rn(T2, make_float2(-vis2.x, -vis2.y));
                const bool dead0 = !(nT.x > 1e-4f), dead1 = !(nT.y > 1e-4f);
                if constexpr (kTrain) {  // (a retired pixel has alpha = 0 from here on: it is never "dead" again)
                    stop0 = dead0 ? cur : stop0;
                    stop1 = dead1 ? cur : stop1;
                }
                vis2.x = dead0 ? 0.0f : vis2.x;
                vis2.y = dead1 ? 0.0f : vis2.y;
                npx2.x = dead0 ? -INFINITY : npx2.x;
                npx2.y = dead1 ? -INFINITY : npx2.y;
                // T - (alpha T or 0): the subtraction again with the selected operand instead of two more selects
                // (the same value as nT where the pixel goes on, T itself where it retired)
                T2 = __fadd2_rn(T2, make_float2(-vis2.x, -vis2.y));
                acc_r = __ffma2_rn(dup2(cr), vis2, acc_r);
                acc_g = __ffma2_rn(dup2(cg), vis2, acc_g);
                acc_b = __ffma2_rn(dup2(cb), vis2, acc_b);
            };
            // two copies of the walk: chunks whose survivors are all "plain" (the common case) run without the
            // sigma < 0 test and without the 0.999 clamp
            auto walk = [&](auto plain_tag) {
                constexpr bool kPlain = decltype(plain_tag)::value;
                while (mask) {
                    // highest set bit = next Gaussian, front to back (bfind -> a single FLO; written in PTX
                    // because nvcc rewrites 31 - clz(x) into a longer clz-based sequence)
                    unsigned int b_hi, below;
                    asm("bfind.u32 %0, %1;" : "=r"(b_hi) : "r"(mask));
                    asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(below) : "r"(b_hi));  // bits below b_hi
                    mask &= below;
                    const uint32_t ra = a_hi - b_hi * (uint32_t)(kPairRec * sizeof(float4));
                    BSPLAT_DASSERT(ra >= s_g_addr && ra + 48u <= s_g_addr + kStageRecs * 48u && c0 + 31 - (int)b_hi < bs);
                    const float4 p0 = lds_f4(ra), p1 = lds_f4(ra + 16u);
                    float cb, tau = 0.0f;
                    if (kPlain) {
                        cb = lds_f1(ra + 32u);
                    } else {
                        const float2 bt = lds_f2(ra + 32u);
                        cb = bt.x; tau = bt.y;
                    }
                    composite(alpha_of(plain_tag, tau, p0, p1), p1.z, p1.w, cb,
                              kTrain ? (int32_t)(b0 + c0 + 31 - (int)b_hi) : 0);
                }
            };
            if (any_special) walk(std::false_type{});
            else walk(std::true_type{});
        }
    }

    pdl_trigger();  // only the output is left: the next kernel of the stream may be staged now
    if (kRec) cp_async_wait_all();  // nothing may still be landing in shared memory when it is reused below
    // sync-free frames: no intersections at all => all-zero image (render.py:73-76), decided on the device
    const float bgs = (m_dev != nullptr && *m_dev == 0ull) ? 0.0f : 1.0f;
    const float bgr = bgs * __ldg(background), bgg = bgs * __ldg(background + 1), bgb = bgs * __ldg(background + 2);
    const float T0 = T2.x, T1 = T2.y, ar0 = acc_r.x, ar1 = acc_r.y, ag0 = acc_g.x, ag1 = acc_g.y, ab0 = acc_b.x,
                ab1 = acc_b.y;
    if constexpr (kTrain) {
        if (in0) { train.final_T[(int64_t)i0 * W + j] = T0; train.last_idx[(int64_t)i0 * W + j] = stop0 - 1; }
        if (in1) { train.final_T[(int64_t)i1 * W + j] = T1; train.last_idx[(int64_t)i1 * W + j] = stop1 - 1; }
    }
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
// 2 = the same without sub-tile culling (exactness A/B), 1 = faithful.
size_t raster_workspace_bytes(int64_t N) { return (size_t)(N > 0 ? N : 1) * kPairRec * sizeof(float4); }
// long-list pre-pass scratch of a fused frame: survivor ids [M_cap] + per-chunk counts
size_t raster_long_surv_bytes(int64_t M_cap) { return (size_t)(M_cap > 0 ? M_cap : 1) * sizeof(int32_t); }
size_t raster_long_cnt_bytes(int64_t M_cap, int64_t n_tiles) {
    return (size_t)(kMaxLongSlots + (M_cap > 0 ? M_cap : 1) / kLongChunk + n_tiles + 2) * sizeof(uint32_t);
}

// rec_ready: the records in rec_ws are already written (projection epilogue); otherwise the record kernel runs here.
// surv / chunk_cnt / long_barrier (optional, fused frames): scratch of the long-list pre-pass; long_barrier is one
// uint32 that is zero when the stream gets here.
int rasterize_launch(int64_t N, int channels, const float* means2d, const float* conics, const float* colors,
                     const float* opacities, const float* background_dev,
                     const int32_t* tile_ranges, const int32_t* tile_order, const int32_t* sorted_ids, int W,
                     int H, int tile_size, int row_begin, int row_end, int mode, float* image,
                     unsigned long long* stats, const unsigned long long* m_dev, void* rec_ws,
                     cudaStream_t stream, const PeerImages* peers_in, const int32_t* rec_list,
                     const unsigned long long* rec_list_n, bool rec_ready, int32_t* surv, uint32_t* chunk_cnt,
                     uint32_t* long_barrier) {
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
    const bool fast_mode = (mode == BSPLAT_RASTER_FAST || mode == BSPLAT_RASTER_FAST_NOCULL);
    if (mode != BSPLAT_RASTER_FAITHFUL && !fast_mode) return BSPLAT_E_ARG;
    // (the raw-array staging reads means2d as float2: an unaligned view falls back to the faithful kernel)
    const bool inputs_ok = (means2d != nullptr && (reinterpret_cast<uintptr_t>(means2d) & 7u) == 0) ||
                           (rec_ready && rec_ws != nullptr);
    const bool fast_ok = fast_mode && tile_size == kFastTile && channels == 3 && stats == nullptr && inputs_ok;
    if (peers.n > 0 && !(fast_ok && mode == BSPLAT_RASTER_FAST))
        return BSPLAT_E_ARG;  // the fused exchange exists in the default 16x16 RGB kernel only
    if (fast_ok) {
        const float* bg = background_dev;
        const int vec = ((reinterpret_cast<uintptr_t>(image) & 15u) == 0 && (W % 4) == 0) ? 1 : 0;
        const unsigned grid = (unsigned)(tiles_w * (row_end - row_begin));
        const int first_tile = row_begin * tiles_w;
        const bool have_rec = rec_ws != nullptr && N > 0 && (reinterpret_cast<uintptr_t>(rec_ws) & 15u) == 0;
        float4* recp = static_cast<float4*>(rec_ws);
        if (have_rec) {
            if (!rec_ready) {
                if (!means2d || !conics || !colors || !opacities || (reinterpret_cast<uintptr_t>(means2d) & 7u) != 0)
                    return BSPLAT_E_ARG;
                raster_pair_prep_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(N, means2d, conics, colors,
                                                                                        opacities, recp, rec_list,
                                                                                        rec_list_n);
                BSPLAT_LAUNCH_CHECK();
            }
            // BSPLAT_DEBUG (A/B measurements only, not part of the interface): "noprepass" / "roworder"
            static const char* dbg = getenv("BSPLAT_DEBUG");
            if (dbg && strstr(dbg, "roworder")) tile_order = nullptr;
            const bool prepass = mode == BSPLAT_RASTER_FAST && surv != nullptr && chunk_cnt != nullptr &&
                                 long_barrier != nullptr && tile_order != nullptr &&
                                 !(dbg && strstr(dbg, "noprepass"));
            if (prepass) {
                // one CTA per SM: the grid-wide barrier inside needs every CTA resident
                BSPLAT_LAUNCH_PDL((raster_long_compact_kernel), 148, kLongThreads, 0, stream, N, recp, tile_ranges, tile_order, (int)grid,
                                                                             sorted_ids, tiles_w, surv, chunk_cnt,
                                                                             long_barrier);
                BSPLAT_LAUNCH_CHECK();
            }
            if (mode == BSPLAT_RASTER_FAST_NOCULL)
                BSPLAT_LAUNCH_PDL((raster_pair_kernel<false, true>), grid, kPairThreads, 0, stream, N, recp, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                    tiles_w, image, vec, m_dev, peers, nullptr, nullptr, TrainOut{nullptr, nullptr});
            else
                BSPLAT_LAUNCH_PDL((raster_pair_kernel<true, true>), grid, kPairThreads, 0, stream, N, recp, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W, H,
                    tiles_w, image, vec, m_dev, peers, prepass ? surv : nullptr,
                    prepass ? chunk_cnt : nullptr, TrainOut{nullptr, nullptr});
        } else {
            if (N > 0 && (!means2d || !conics || !colors || !opacities ||
                          (reinterpret_cast<uintptr_t>(means2d) & 7u) != 0))
                return BSPLAT_E_ARG;
            if (mode == BSPLAT_RASTER_FAST_NOCULL)
                BSPLAT_LAUNCH_PDL((raster_pair_kernel<false, false>), grid, kPairThreads, 0, stream, N, nullptr, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W,
                    H, tiles_w, image, vec, m_dev, peers, nullptr, nullptr, TrainOut{nullptr, nullptr});
            else
                BSPLAT_LAUNCH_PDL((raster_pair_kernel<true, false>), grid, kPairThreads, 0, stream, N, nullptr, means2d, conics, colors, opacities, bg, tile_ranges, tile_order, first_tile, sorted_ids, W,
                    H, tiles_w, image, vec, m_dev, peers, nullptr, nullptr, TrainOut{nullptr, nullptr});
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

// the records of all N Gaussians (stage-level callers outside this file: the fast backward pass)
int raster_records_launch(int64_t N, const float* means2d, const float* conics, const float* colors,
                          const float* opacities, void* rec_ws, cudaStream_t stream) {
    if (N <= 0) return BSPLAT_OK;
    raster_pair_prep_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(N, means2d, conics, colors, opacities,
                                                                            static_cast<float4*>(rec_ws), nullptr,
                                                                            nullptr);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

// Training-side forward through the pair kernel (16x16 tiles, RGB): image bit-identical to the inference kernel's,
// plus final_T / last_idx for rasterize_bwd.cu.  rec_ws (optional, raster_workspace_bytes(N)): records + cp.async
// staging; tile_order (optional): heavy tiles first.
int rasterize_train_fast_launch(int64_t N, const float* means2d, const float* conics, const float* colors,
                                const float* opacities, const float* background_dev, const int32_t* tile_ranges,
                                const int32_t* tile_order, const int32_t* sorted_ids, int W, int H, float* image,
                                float* final_T, int32_t* last_idx, void* rec_ws, cudaStream_t stream) {
    if (W <= 0 || H <= 0 || !background_dev || !image || !final_T || !last_idx || !tile_ranges) return BSPLAT_E_ARG;
    const int tiles_w = (W + kFastTile - 1) / kFastTile, tiles_h = (H + kFastTile - 1) / kFastTile;
    if (N > 0 && (!means2d || !conics || !colors || !opacities || (reinterpret_cast<uintptr_t>(means2d) & 7u) != 0))
        return BSPLAT_E_ARG;
    const int vec = ((reinterpret_cast<uintptr_t>(image) & 15u) == 0 && (W % 4) == 0) ? 1 : 0;
    const unsigned grid = (unsigned)(tiles_w * tiles_h);
    PeerImages peers;
    peers.n = 0;
    const TrainOut train{final_T, last_idx};
    const bool have_rec = rec_ws != nullptr && N > 0 && (reinterpret_cast<uintptr_t>(rec_ws) & 15u) == 0;
    if (have_rec) {
        float4* recp = static_cast<float4*>(rec_ws);
        raster_pair_prep_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(N, means2d, conics, colors, opacities,
                                                                                recp, nullptr, nullptr);
        BSPLAT_LAUNCH_CHECK();
        raster_pair_kernel<true, true, true><<<grid, kPairThreads, 0, stream>>>(
            N, recp, means2d, conics, colors, opacities, background_dev, tile_ranges, tile_order, 0, sorted_ids, W, H,
            tiles_w, image, vec, nullptr, peers, nullptr, nullptr, train);
    } else {
        raster_pair_kernel<true, false, true><<<grid, kPairThreads, 0, stream>>>(
            N, nullptr, means2d, conics, colors, opacities, background_dev, tile_ranges, tile_order, 0, sorted_ids, W,
            H, tiles_w, image, vec, nullptr, peers, nullptr, nullptr, train);
    }
    BSPLAT_LAUNCH_CHECK();
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
