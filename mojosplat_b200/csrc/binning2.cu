// Stage 2 (default path) -- two-level binning.
//
// The reference sorts the (gaussian, tile) pairs twice: argsort by depth, then a STABLE argsort by
// tile id (mojosplat/binning.py:223-231).  This path keeps that structure but moves the depth sort to
// where it is cheap:
//   0. per Gaussian: full-frame tile rectangle + monotone depth key + the 4 digit histograms of the keys
//      (epilogue of the projection kernel in fused frames; bin_prep_kernel for the stage-level entry point)
//   0b. row-band / packed frames only: one stable compaction of the Gaussians that own >= 1 tile
//   1. depth-sort the (remaining) Gaussians  4 onesweep passes over (uint32 depth key, index) -- N items
//   2. count + scan in depth order           gather the rectangle of Gaussian perm[j], exclusive prefix sum, M
//   3. emit in depth order                   (tile id, gaussian id) pairs; tile-digit histograms on the fly
//   4. stable sort by tile id only           ceil(log2(n_tiles)) bits in ceil(bits / 8) onesweep passes over M items
//   5. tile ranges (+ heavy-first order)     from the per-tile counts the last pass yields
// Result = ascending (tile, depth, gaussian index): bit-identical to the single-level sort of
// (tile << depth_bits | depth_key) keys and to the reference's lists (canonical tie order),
// with ~4x less sort traffic: M-scale data is moved by 2 passes of 8 B pairs instead of 6 passes of 12 B.
#include "binning.cuh"
#include "radix_sort.cuh"

namespace bsplat {

// ---- 0. stage-level entry point: rectangles + depth keys + digit histograms -----------------
__global__ void __launch_bounds__(256)
bin_prep_kernel(const int64_t N, const float* __restrict__ means2d, const void* __restrict__ radii,
                const int radii_is_float, const float* __restrict__ depths, const BinParams p,
                const float inv_tile_size, uint2* __restrict__ rects_in, uint32_t* __restrict__ keys,
                uint32_t* __restrict__ hist /* [4][256], nullable */) {
    __shared__ uint32_t s_hist[4][kRadix];
    const bool want_hist = hist != nullptr;
    if (want_hist) {
        for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
        __syncthreads();
    }
    const uint32_t lane = lane_id();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t warp_start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    for (int64_t wi = warp_start; wi < N; wi += stride) {  // warp-uniform trip count
        const int64_t i = wi + lane;
        const bool valid = i < N;
        uint32_t k = 0;
        if (valid) {
            float mx, my, rx, ry;
            load_mean_radii(means2d, radii, radii_is_float, i, mx, my, rx, ry);
            const TileRect r = tile_rect_inv(mx, my, rx, ry, p.W, p.H, p.tile_size_f, inv_tile_size, p.tiles_w,
                                             p.tiles_h, p.semantics);
            rects_in[i] = make_uint2((uint32_t)r.x0 | ((uint32_t)r.y0 << 16),
                                     (uint32_t)(r.x1 - r.x0) | ((uint32_t)(r.y1 - r.y0) << 16));
            k = depth_key(__ldg(depths + i));
            keys[i] = k;
            if (want_hist) {
                atomicAdd(&s_hist[0][k & 0xffu], 1u);
                atomicAdd(&s_hist[1][(k >> 8) & 0xffu], 1u);
                atomicAdd(&s_hist[2][(k >> 16) & 0xffu], 1u);
            }
        }
        if (want_hist) {
            // sign + exponent bits: a handful of distinct values per warp -> aggregate before the atomic
            const uint32_t top = valid ? (k >> 24) : 0x100u;
            const uint32_t peers = __match_any_sync(0xffffffffu, top);
            if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[3][top], (uint32_t)__popc(peers));
        }
    }
    if (want_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) {
            const uint32_t v = (&s_hist[0][0])[i];
            if (v) atomicAdd(hist + i, v);
        }
    }
}

// Decoupled look-back of one warp: exclusive prefix of chunk `chunk` = sum of the aggregates of its predecessors back
// to (and including) the nearest published inclusive prefix.  Four windows of 32 predecessors are fetched per round
// trip: when a whole wave of chunks publishes its aggregates at the same time (persistent grids: every resident CTA),
// a chunk deep in the wave otherwise walks back 32 links per ~1 us.
__device__ __forceinline__ unsigned long long lookback_prefix(const unsigned long long* __restrict__ status,
                                                              const int64_t chunk, const int lane) {
    constexpr int kWin = 4;
    unsigned long long prefix = 0;
    int64_t j = chunk - 1;
    while (true) {
        unsigned long long v[kWin];
#pragma unroll
        for (int u = 0; u < kWin; ++u) {
            const int64_t idx = j - 32 * u - lane;  // lane 0 of window 0 = nearest predecessor
            v[u] = idx >= 0 ? ld_relaxed_u64(status + idx) : kFlagPrefix;
        }
        bool done = false, stalled = false;
#pragma unroll
        for (int u = 0; u < kWin; ++u) {
            if (done || stalled) break;
            const unsigned ready = __ballot_sync(0xffffffffu, (v[u] & ~kValueMask) != 0);
            const unsigned pref = __ballot_sync(0xffffffffu, (v[u] & kFlagPrefix) != 0);
            const unsigned first_pref = pref ? (unsigned)(__ffs(pref) - 1) : 32u;
            const unsigned need = first_pref == 32u ? 0xffffffffu : ((2u << first_pref) - 1u);
            if ((ready & need) != need) { stalled = true; break; }  // not published yet
            unsigned long long contrib = ((unsigned)lane <= first_pref) ? (v[u] & kValueMask) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
            prefix += contrib;
            if (first_pref != 32u) done = true;
            else j -= 32;
        }
        if (done) break;
        if (stalled) __nanosleep(64);  // polite re-poll from the first window that was not ready
    }
    return prefix;
}

// ---- 0b. stable compaction of the Gaussians that own >= 1 tile of the band -------------------------------
// Row-band frames (every rank of a band split would otherwise depth-sort all N Gaussians) and packed frames
// (gsplat rules: culled Gaussians own no tile; projection.mojo:73-87, 213-244).  One pass: persistent CTAs take
// chunks of 2 048 Gaussians in ticket order, flag = rectangle clipped to the band is non-empty, block scan +
// decoupled look-back over the chunk totals, in-band (depth key, index) pairs written in index order, the digit
// histograms of the surviving keys accumulated on the way.  n_out[0] = survivors, n_out[1] = Gaussians with a
// non-empty FULL-FRAME rectangle (the "no intersections at all" rule of render.py:73-76 is a whole-frame rule).
constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 8;
constexpr int kCompactChunk = kCompactThreads * kCompactItems;

// kPre (row-band frames under the torch rules): the same stable compaction, but BEFORE the projection -- the flag is a
// cheap conservative test "can this Gaussian's tile rectangle reach the band's rows at all?" on the camera-space mean
// and the largest scale (BandPretest below), the output the list of candidate indices; only those are projected, and
// the exact compaction afterwards runs over that list (`list` / `list_n`) instead of over all N.
struct BandPretest {
    float r1[3], r2[3], t1, t2;   // rows 1 and 2 of the world -> camera transform
    float fy, cy;
    float jn2;                    // 1 + max(lim_y_neg, lim_y_pos)^2: |J row 1|^2 <= (fy / z)^2 jn2
    float gain2;                  // 9 |Rv|_2^2 (1 + 1e-3): |R(q)|_2 <= 3 for any |q| <= 1 (= 1 for unit q)
    float eps2d;
    float y_lo, y_hi;             // the band in pixels, widened by a pixel; -inf / +inf at the image border
};

// Flags of one thread's kCompactItems consecutive slots (bit k <-> slot base + k); `whole` counts the slots whose
// full-frame rectangle is non-empty (exact mode).
template <bool kPre>
__device__ __forceinline__ unsigned int compact_flags(const int64_t N, const int64_t base,
                                                      const int32_t* __restrict__ list,
                                                      const uint2* __restrict__ rects_in, const int row_begin,
                                                      const int row_end, const float* __restrict__ means3d,
                                                      const float* __restrict__ log_scales, const BandPretest& pt,
                                                      unsigned int& whole) {
    unsigned int flags = 0;
    if (base >= N) return 0u;
    if constexpr (kPre) {
        // a thread's 8 Gaussians are 96 contiguous bytes of each input array: six 128-bit loads instead of 24 scalar
        // ones when the slots are all valid and the arrays are 16-byte aligned
        static_assert((3 * kCompactItems) % 4 == 0, "vector loads of a thread's rows");
        float pre_m[3 * kCompactItems], pre_s[3 * kCompactItems];
        const bool vec = base + kCompactItems <= N &&
                         ((reinterpret_cast<uintptr_t>(means3d) | reinterpret_cast<uintptr_t>(log_scales)) & 15u) == 0;
        if (vec) {
            const float4* vm = reinterpret_cast<const float4*>(means3d + 3 * base);
            const float4* vs = reinterpret_cast<const float4*>(log_scales + 3 * base);
#pragma unroll
            for (int v = 0; v < 3 * kCompactItems / 4; ++v) {
                const float4 a = __ldg(vm + v), b = __ldg(vs + v);
                pre_m[4 * v] = a.x; pre_m[4 * v + 1] = a.y; pre_m[4 * v + 2] = a.z; pre_m[4 * v + 3] = a.w;
                pre_s[4 * v] = b.x; pre_s[4 * v + 1] = b.y; pre_s[4 * v + 2] = b.z; pre_s[4 * v + 3] = b.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 3 * kCompactItems; ++e) {
                const bool in = base + e / 3 < N;
                pre_m[e] = in ? __ldg(means3d + 3 * base + e) : 0.0f;
                pre_s[e] = in ? __ldg(log_scales + 3 * base + e) : 0.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if (base + k >= N) break;
            // conservative: the exact projection computes the same camera-space mean; its radius is 0 or
            // ceil(3.33 sqrt(c11)) with c11 <= |J row 1|^2 |Rv|^2 |R(q)|^2 smax^2 + eps2d.  Anything not provably
            // outside the band stays (NaN / inf compare false).
            const float mx = pre_m[3 * k], my = pre_m[3 * k + 1], mz = pre_m[3 * k + 2];
            const float l0 = pre_s[3 * k], l1 = pre_s[3 * k + 1], l2 = pre_s[3 * k + 2];
            const float lmax = fmaxf(fmaxf(l0, l1), l2);
            const float mcy = __fmaf_rn(pt.r1[2], mz, __fmaf_rn(pt.r1[1], my, __fmul_rn(pt.r1[0], mx))) + pt.t1;
            const float mcz = __fmaf_rn(pt.r2[2], mz, __fmaf_rn(pt.r2[1], my, __fmul_rn(pt.r2[0], mx))) + pt.t2;
            const float ny = __fmaf_rn(pt.cy, mcz, __fmul_rn(pt.fy, mcy));
            // approximate quotients and no square root: everything below carries a relative margin of 1e-4 or
            // more, and the radius enters squared
            const float rz = __frcp_rn(mcz);  // (inf for z = 0: nothing is skipped)
            const float m2y = ny * rz;
            const float smax = __expf(lmax);
            const float fz = pt.fy * rz;
            // (3.33 sqrt(c11) 1.0001)^2 with c11 <= (fy / z)^2 jn2 gain2 smax^2 (1 + 1e-3) + eps2d
            const float r2 = 11.0889f * 1.002f * (fz * fz * pt.jn2 * pt.gain2 * smax * smax * 1.001f + pt.eps2d);
            const float slack = 1e-4f * fabsf(m2y) + 2.0f;  // (the ceil and a pixel of rounding)
            const float da = pt.y_lo - m2y - slack;          // room above the band: skip if radius < da
            const float db = m2y - pt.y_hi - slack;          // room below the band
            const bool above = da > 0.0f && r2 < da * da;
            const bool below = db > 0.0f && r2 < db * db;
            const bool nan_in = !(l0 == l0) || !(l1 == l1) || !(l2 == l2);  // (fmaxf drops NaNs)
            if (!(above || below) || nan_in) flags |= 1u << k;
        }
    } else {
        int32_t gidx[kCompactItems];
        if (list != nullptr && base + kCompactItems <= N) {  // (a thread's slots are contiguous: 128-bit loads)
            static_assert(kCompactItems % 4 == 0, "128-bit loads over a thread's slots");
            const int4* vl = reinterpret_cast<const int4*>(list + base);
#pragma unroll
            for (int v = 0; v < kCompactItems / 4; ++v) {
                const int4 q = __ldg(vl + v);
                gidx[4 * v] = q.x; gidx[4 * v + 1] = q.y; gidx[4 * v + 2] = q.z; gidx[4 * v + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kCompactItems; ++k) {
                const int64_t i = base + k;
                gidx[k] = (i < N) ? (list ? __ldg(list + i) : (int32_t)i) : -1;
            }
        }
        uint2 rc[kCompactItems];
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) rc[k] = gidx[k] >= 0 ? __ldg(rects_in + gidx[k]) : make_uint2(0u, 0u);
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            const uint2 cl = clip_rect_rows(rc[k], row_begin, row_end);
            if ((cl.y & 0xffffu) != 0u && (cl.y >> 16) != 0u) flags |= 1u << k;
            if ((rc[k].y & 0xffffu) != 0u && (rc[k].y >> 16) != 0u) ++whole;
        }
    }
    return flags;
}

// Stable compaction in two launches without any dependency between CTAs (a single-pass chained scan left every CTA
// waiting at a barrier for its look-back -- 46 % of the stall samples, 88 us for 6 M Gaussians):
//   A. flags of every slot (one byte per thread), the survivors per chunk of 2 048; the CTA that finishes last
//      turns the chunk counts into exclusive offsets and publishes the total (n_out[0]);
//   B. scatter: block scan of the flag bytes + the chunk offset; exact mode also moves the depth keys and
//      accumulates their digit histograms.
// n_out[1] (exact mode) = Gaussians with a non-empty FULL-FRAME rectangle (+ whole_extra).
template <bool kPre>
__global__ void __launch_bounds__(kCompactThreads)
bin_compact_flags_kernel(const int64_t N_host, const int32_t* __restrict__ list,
                         const unsigned long long* __restrict__ list_n, const uint2* __restrict__ rects_in,
                         const int row_begin, const int row_end, const float* __restrict__ means3d,
                         const float* __restrict__ log_scales, const BandPretest pt,
                         uint8_t* __restrict__ flag_bytes, uint32_t* __restrict__ counts,
                         uint32_t* __restrict__ offs, uint32_t* __restrict__ done,
                         unsigned long long* __restrict__ n_out, const unsigned long long whole_extra) {
    pdl_wait();  // (programmatic dependent launch: nothing of the predecessor is read before this)
    __shared__ unsigned int s_cnt[kCompactThreads / 32], s_whole[kCompactThreads / 32];
    __shared__ bool s_last;
    const int64_t N = list_n ? (int64_t)(*list_n) : N_host;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kCompactChunk + (int64_t)tid * kCompactItems;
    unsigned int whole = 0;
    const unsigned int flags = compact_flags<kPre>(N, base, list, rects_in, row_begin, row_end, means3d, log_scales, pt,
                                                   whole);
    flag_bytes[(int64_t)blockIdx.x * kCompactThreads + tid] = (uint8_t)flags;
    unsigned int cnt = __popc(flags);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    whole = __reduce_add_sync(0xffffffffu, whole);
    if (lane == 0) { s_cnt[warp] = cnt; s_whole[warp] = whole; }
    __syncthreads();
    if (tid == 0) {
        unsigned int c = 0, wsum = 0;
#pragma unroll
        for (int w = 0; w < kCompactThreads / 32; ++w) { c += s_cnt[w]; wsum += s_whole[w]; }
        counts[blockIdx.x] = c;
        if (!kPre && wsum) atomicAdd(n_out + 1, (unsigned long long)wsum);
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- the last CTA: exclusive scan of the chunk counts ----
    __threadfence();
    const int n_chunks = (int)gridDim.x;
    const int per = (n_chunks + kCompactThreads - 1) / kCompactThreads;
    const int c0 = tid * per;
    unsigned int sum = 0;
    for (int q = 0; q < per; ++q)
        if (c0 + q < n_chunks) sum += __ldcg(counts + c0 + q);
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_cnt[warp] = incl;
    __syncthreads();
    unsigned int run = incl - sum;
    for (int w = 0; w < warp; ++w) run += s_cnt[w];
    for (int q = 0; q < per; ++q)
        if (c0 + q < n_chunks) {
            offs[c0 + q] = run;
            run += __ldcg(counts + c0 + q);
        }
    if (tid == kCompactThreads - 1) {
        n_out[0] = run;
        if (!kPre && whole_extra) atomicAdd(n_out + 1, whole_extra);
    }
}

template <bool kPre>
__global__ void __launch_bounds__(kCompactThreads)
bin_compact_scatter_kernel(const int64_t n_chunks, const int32_t* __restrict__ list,
                           const uint8_t* __restrict__ flag_bytes, const uint32_t* __restrict__ offs,
                           const uint32_t* __restrict__ keys_full, uint32_t* __restrict__ keys,
                           int32_t* __restrict__ perm, uint32_t* __restrict__ hist /* [4][256] */) {
    pdl_wait();  // (programmatic dependent launch: nothing of the predecessor is read before this)
    __shared__ uint32_t s_hist[kPre ? 1 : 4][kRadix];
    __shared__ unsigned int s_warp_sum[kCompactThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!kPre)
        for (int i = tid; i < 4 * kRadix; i += kCompactThreads) (&s_hist[0][0])[i] = 0;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        __syncthreads();  // s_warp_sum of the previous chunk is consumed (and the zeroing above is done)
        const unsigned int flags = flag_bytes[chunk * kCompactThreads + tid];
        const unsigned int cnt = __popc(flags);
        unsigned int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        unsigned int pos = __ldg(offs + chunk) + incl - cnt;
        for (int w = 0; w < warp; ++w) pos += s_warp_sum[w];
        const int64_t base = chunk * kCompactChunk + (int64_t)tid * kCompactItems;
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if ((flags >> k) & 1u) {
                const int32_t gi = list ? __ldg(list + base + k) : (int32_t)(base + k);
                BSPLAT_DASSERT(gi >= 0 && (int64_t)pos < n_chunks * kCompactChunk);
                perm[pos] = gi;
                if (!kPre) {
                    const uint32_t kk = __ldg(keys_full + gi);
                    keys[pos] = kk;
                    atomicAdd(&s_hist[0][kk & 0xffu], 1u);
                    atomicAdd(&s_hist[1][(kk >> 8) & 0xffu], 1u);
                    atomicAdd(&s_hist[2][(kk >> 16) & 0xffu], 1u);
                    atomicAdd(&s_hist[3][kk >> 24], 1u);
                }
                ++pos;
            }
        }
    }
    if constexpr (!kPre) {
        __syncthreads();
        for (int i = tid; i < 4 * kRadix; i += kCompactThreads) {
            const uint32_t v = (&s_hist[0][0])[i];
            if (v) atomicAdd(hist + i, v);
        }
    }
}

// ---- 2. count + scan in depth order ----------------------------------------------------------
// Slot j of the depth order holds Gaussian perm[j]: one 8-byte gather of its full-frame rectangle (the
// rectangle arithmetic -- clamps, IEEE divisions -- happened once, in index order, in step 0), clip to the band,
// count, single-pass decoupled-look-back exclusive scan (warp-parallel look-back, 2 048 Gaussians per CTA).
// The clipped rectangles are kept in depth order for the emitter.
// 1 024 threads x 8 slots per CTA: few chunks keep the look-back chain short (all chunks start together, so the last
// one walks back through every predecessor, 32 per round trip), many threads keep all gathers of a chunk in flight.
#ifdef BSPLAT_PHASES
// A/B instrumentation (make phases; benchmarks/sort_phases.py): SM clock at the phase boundaries of count + scan, per CTA;
// columns 0-7 thread 0 (warp 0: the look-back path), 8-13 the first thread of the last warp (the coverage path)
__device__ unsigned long long g_scan_phases[1024][16];
#define BSPLAT_SPHASE(t, i)                                                                              \
    do {                                                                                                 \
        if (threadIdx.x == (t) && ph_row < 1024) g_scan_phases[ph_row][i] = (unsigned long long)clock64(); \
    } while (0)
extern "C" int bsplat_debug_scan_phases(void* host_out, size_t bytes) {
    return (int)cudaMemcpyFromSymbol(host_out, g_scan_phases, bytes < sizeof(g_scan_phases) ? bytes : sizeof(g_scan_phases));
}
#else
#define BSPLAT_SPHASE(t, i)
#endif
constexpr int kScan2Threads = 1024;
constexpr int kScan2Chunk = kScan2Threads * kScanItems;

__global__ void __launch_bounds__(kScan2Threads)
bin_count_scan2_kernel(const int64_t N_host, const unsigned long long* __restrict__ n_dev,
                       const int32_t* __restrict__ perm, const uint2* __restrict__ rects_in, const int row_begin,
                       const int row_end, uint32_t* __restrict__ offsets, bsplat_bin_info* __restrict__ info,
                       unsigned long long* __restrict__ ws, uint2* __restrict__ rects,
                       uint32_t* __restrict__ hist_xy /* [2][256] or null */, const int tiles_w) {
    __shared__ unsigned int s_chunk;
    // 2-D tile keys (hist_xy != null): the digit histograms of the two tile-sort passes are the column and row
    // coverage counts -- a rectangle adds h to each of its w columns and w to each of its h rows, i.e. four updates of
    // two difference arrays per Gaussian (prefix-summed once per CTA), instead of one update per pair in the emitter
    __shared__ uint32_t s_diff[2][kRadix + 1];
    if (hist_xy != nullptr) {
        for (int i = threadIdx.x; i < 2 * (kRadix + 1); i += kScan2Threads) (&s_diff[0][0])[i] = 0u;
    }
    __shared__ unsigned long long s_warp_sum[kScan2Threads / 32];
    __shared__ unsigned long long s_prefix;

    const int tid = threadIdx.x;
#ifdef BSPLAT_PHASES
    unsigned ph_row = blockIdx.x;
    BSPLAT_SPHASE(0, 0);
#endif
    // (the ticket counter was zeroed at the start of the frame: drawn, like the shared-memory reset above, before the
    // wait for the previous kernel of the stream -- programmatic dependent launch, common.cuh)
    if (tid == 0) s_chunk = atomicAdd(reinterpret_cast<unsigned int*>(ws), 1u);
    pdl_wait();
    BSPLAT_SPHASE(0, 1);
    // n_dev: the number of items lives on the device (compacted depth order); the grid is sized by N_host
    const int64_t N = n_dev ? (int64_t)(*n_dev) : N_host;
    __syncthreads();
    const unsigned int chunk = s_chunk;
    unsigned long long* status = ws + 1;
    const int64_t base = (int64_t)chunk * kScan2Chunk;
    if (base >= N) {  // surplus chunk (tickets are handed out in order: no lower chunk ever waits for this one)
        if (N == 0 && chunk == 0 && tid == 0) { offsets[0] = 0u; info->n_isect = 0ull; }
        return;
    }
    // blocked arrangement: thread t owns slots base + t*kScanItems + k (keeps the scan trivial); all gathers of a
    // thread are issued before the first one is used
    // A thread's kScanItems slots are contiguous in memory: when all of them are valid, the index loads and the
    // offset / rectangle stores go as 128-bit accesses (8 scalar accesses per thread, each touching 32 different
    // sectors per warp, left the kernel waiting on the load / store queue: stall_lg was its top stall reason)
    static_assert(kScanItems % 4 == 0, "128-bit accesses over a thread's slots");
    const int64_t slot0 = base + (int64_t)tid * kScanItems;
    const bool full = slot0 + kScanItems <= N;
    int32_t gi[kScanItems];
    uint2 rc[kScanItems];
    if (full && perm != nullptr) {
        const int4* vp = reinterpret_cast<const int4*>(perm + slot0);
#pragma unroll
        for (int v = 0; v < kScanItems / 4; ++v) {
            const int4 q = __ldg(vp + v);
            gi[4 * v] = q.x; gi[4 * v + 1] = q.y; gi[4 * v + 2] = q.z; gi[4 * v + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int64_t jj = slot0 + k;
            gi[k] = (jj < N) ? (perm ? __ldg(perm + jj) : (int32_t)jj) : -1;
        }
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) rc[k] = gi[k] >= 0 ? __ldg(rects_in + gi[k]) : make_uint2(0u, 0u);
    uint32_t cnt[kScanItems];
    uint32_t thread_sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t jj = base + (int64_t)tid * kScanItems + k;
        const uint2 cl = clip_rect_rows(rc[k], row_begin, row_end);
        cnt[k] = (cl.y & 0xffffu) * (cl.y >> 16);
        rc[k] = cl;
        if (!full && jj < N) rects[jj] = cl;
        thread_sum += cnt[k];
    }
    if (full) {
        uint4* vr = reinterpret_cast<uint4*>(rects + slot0);
#pragma unroll
        for (int v = 0; v < kScanItems / 2; ++v) vr[v] = make_uint4(rc[2 * v].x, rc[2 * v].y, rc[2 * v + 1].x, rc[2 * v + 1].y);
    }
    const int lane = tid & 31, warp = tid >> 5;
    if (thread_sum == 0xffffffffu) return;  // (never: keeps the loads above in front of the clock reads below)
    BSPLAT_SPHASE(0, 2); BSPLAT_SPHASE(kScan2Threads - 32, 8);
    auto add_coverage = [&]() {
        // (s_diff was zeroed before the first barrier above; uint32 wrap-around carries the negative steps.)  Runs of
        // identical rectangles in a thread's depth-consecutive slots -- the culled Gaussians that the torch rules clamp
        // into a corner tile -- are merged first: they would all hit the same four words.
        // Rectangles clipped at an image (or band) border all update the same word: those four words are summed in
        // registers and reduced over the warp instead (a same-address shared-memory atomic serialises its lanes).
        uint32_t px = 0xffffffffu, py = 0u, mult = 0u;
        uint32_t edge_l = 0u, edge_r = 0u, edge_t = 0u, edge_b = 0u;
        const uint32_t x_end = (uint32_t)tiles_w, y_beg = (uint32_t)row_begin, y_end = (uint32_t)row_end;
        auto flush = [&]() {
            if (mult == 0u) return;
            const uint32_t x0 = px & 0xffffu, y0 = px >> 16, w = py & 0xffffu, h = py >> 16;
            BSPLAT_DASSERT(x0 + w <= (uint32_t)kRadix && y0 + h <= (uint32_t)kRadix && x0 + w <= x_end && y0 + h <= y_end);
            if (x0 == 0u) edge_l += h * mult; else atomicAdd(&s_diff[0][x0], h * mult);
            if (x0 + w == x_end) edge_r += h * mult; else atomicAdd(&s_diff[0][x0 + w], 0u - h * mult);
            if (y0 == y_beg) edge_t += w * mult; else atomicAdd(&s_diff[1][y0], w * mult);
            if (y0 + h == y_end) edge_b += w * mult; else atomicAdd(&s_diff[1][y0 + h], 0u - w * mult);
        };
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (cnt[k] == 0u) continue;
            const uint2 cl = clip_rect_rows(rc[k], row_begin, row_end);
            if (cl.x == px && cl.y == py) {
                ++mult;
            } else {
                flush();
                px = cl.x; py = cl.y; mult = 1u;
            }
        }
        flush();
        edge_l = __reduce_add_sync(0xffffffffu, edge_l);
        edge_r = __reduce_add_sync(0xffffffffu, edge_r);
        edge_t = __reduce_add_sync(0xffffffffu, edge_t);
        edge_b = __reduce_add_sync(0xffffffffu, edge_b);
        if ((tid & 31) == 0) {
            if (edge_l) atomicAdd(&s_diff[0][0], edge_l);
            if (edge_r) atomicAdd(&s_diff[0][x_end], 0u - edge_r);
            if (edge_t) atomicAdd(&s_diff[1][y_beg], edge_t);
            if (edge_b) atomicAdd(&s_diff[1][y_end], 0u - edge_b);
        }
    };
    // warp 0 (which goes into the look-back below) adds its coverage now; the other 31 warps do it while warp 0 waits
    // for its predecessors, so the chunk's aggregate is published without waiting for the histogram work
    if (hist_xy != nullptr && warp == 0) add_coverage();
    BSPLAT_SPHASE(0, 3);
    unsigned long long incl = thread_sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScan2Threads / 32; ++w) {
        const unsigned long long s = s_warp_sum[w];
        if (w < warp) warp_excl += s;
        block_total += s;
    }
    const unsigned long long thread_excl = warp_excl + incl - thread_sum;
    BSPLAT_SPHASE(0, 4); BSPLAT_SPHASE(kScan2Threads - 32, 9);
    if (hist_xy != nullptr && warp != 0) {
        add_coverage();
        BSPLAT_SPHASE(kScan2Threads - 32, 10);
        asm volatile("bar.sync 2, %0;" ::"n"(kScan2Threads - 32));  // warps 1 .. 31: every update is in s_diff
        BSPLAT_SPHASE(kScan2Threads - 32, 11);
    }
    if (hist_xy != nullptr && tid >= kScan2Threads - 2 * kRadix) {
        // inclusive prefix of the difference arrays = coverage per column (first 256 of these threads) / row (last
        // 256).  The upper half of the CTA does this (and the global flush) while warp 0 is in the look-back.
        const int t2 = tid - (kScan2Threads - 2 * kRadix);
        const int which = t2 >> 8, b = t2 & (kRadix - 1);
        uint32_t v = s_diff[which][b];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += u;
        }
        __shared__ uint32_t s_wtot[2 * kRadix / 32];
        if (lane == 31) s_wtot[t2 >> 5] = v;
        asm volatile("bar.sync 1, %0;" ::"n"(2 * kRadix));  // (the 512 threads of this branch)
        const int w0 = which * (kRadix / 32);
        for (int q = w0; q < (t2 >> 5); ++q) v += s_wtot[q];
        if (v) atomicAdd(hist_xy + which * kRadix + b, v);
    }
    if (warp == 0) {
        unsigned long long prefix = 0;
        if (chunk == 0) {
            if (lane == 0) st_relaxed_u64(status + 0, kFlagPrefix | block_total);
        } else {
            if (lane == 0) st_relaxed_u64(status + chunk, kFlagAgg | block_total);
            prefix = lookback_prefix(status, (int64_t)chunk, lane);
            if (lane == 0) st_relaxed_u64(status + chunk, kFlagPrefix | (prefix + block_total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    BSPLAT_SPHASE(0, 5); BSPLAT_SPHASE(kScan2Threads - 32, 12);
    __syncthreads();
    BSPLAT_SPHASE(0, 6);
    pdl_trigger();  // only the output is left: the next kernel of the stream may be staged now
    unsigned long long run = s_prefix + thread_excl;
    if (full) {
        uint32_t o[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { o[k] = (uint32_t)run; run += cnt[k]; }
        uint4* vo = reinterpret_cast<uint4*>(offsets + slot0);
#pragma unroll
        for (int v = 0; v < kScanItems / 4; ++v) vo[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int64_t i = slot0 + k;
            if (i < N) offsets[i] = (uint32_t)run;
            run += cnt[k];
        }
    }
    // the chunk that owns the last slot publishes the total
    if (tid == kScan2Threads - 1 && base + kScan2Chunk >= N) {
        offsets[N] = (uint32_t)run;
        info->n_isect = run;
    }
    BSPLAT_SPHASE(0, 7);
}

// ---- 3. emission in depth order ------------------------------------------------------------
// Work is balanced over OUTPUT PAIRS, not over Gaussians: the M pairs are cut into tasks of kEmitTask
// consecutive pairs, a warp takes a task, finds the Gaussian that owns its first pair with a 32-ary search over the
// exclusive offsets (4 probes for 1 M Gaussians) and then walks the task 32 pairs at a time: it keeps a window of
// 32 depth-consecutive Gaussians (rectangles and offsets come from the count + scan kernel, coalesced), the owner
// of pair p is found with a 5-step shuffle binary search over the window's offsets, its rectangle fetched with
// shuffles, the tile id computed and both words stored -- every store instruction writes 128 contiguous bytes,
// no lane idles on a short rectangle, and a Gaussian that covers thousands of tiles (the nearest ones, which the
// depth order puts side by side) is shared by as many warps as it has tasks.  The digit histograms of the
// tile-sort passes are accumulated on the fly (shared-memory atomics; the highest digit is run-length aggregated
// with match.any: neighbouring pairs share it).
constexpr int kEmit2Threads = 256;
constexpr uint32_t kEmitTask = 1024;
constexpr int kMaxTilePasses = 4;

// digit layout of the tile sort: pass q sorts bits [shift[q], shift[q] + bits[q]) of the tile id
struct TilePasses {
    int n;
    int shift[kMaxTilePasses];
    int bits[kMaxTilePasses];
};

// kSep: 2-D tile keys (row << 16 | column; images of at most 256 x 256 tiles).  The digit histograms then come from the
// count + scan kernel (column / row coverage), the key needs no multiplication, and k / w is an exact multiply-high
// with a per-width constant (k w^2 < 2^32 holds for such rectangles).  Otherwise: linear tile ids, histograms on the fly.
template <bool kSep>
__global__ void __launch_bounds__(kEmit2Threads)
bin_emit2_kernel(const int64_t N_host, const unsigned long long* __restrict__ n_dev,
                 const int32_t* __restrict__ perm, const uint2* __restrict__ rects,
                 const int tiles_w, const uint32_t* __restrict__ offsets, const TilePasses tp,
                 uint32_t* __restrict__ tile_keys, int32_t* __restrict__ ids,
                 uint32_t* __restrict__ hist /* [kMaxTilePasses][256] */,
                 bsplat_bin_info* __restrict__ info_dev, const int64_t m_cap) {
    __shared__ uint32_t s_hist[kSep ? 1 : kMaxTilePasses][kRadix];  // kSep: ceil(2^32 / w) for w = 1 .. 256 instead
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    if (kSep) {
        // s_hist[0][w - 1] = ceil(2^32 / w) (w = 1: 2^32 does not fit -- handled where it is used); filled before the
        // wait for the previous kernel (programmatic dependent launch)
        for (int i = tid; i < kRadix; i += kEmit2Threads) s_hist[0][i] = i == 0 ? 0u : 0xffffffffu / (uint32_t)(i + 1) + 1u;
    } else {
        for (int i = tid; i < kMaxTilePasses * kRadix; i += kEmit2Threads) (&s_hist[0][0])[i] = 0;
    }
    pdl_wait();
    // sync-free frames: the pair buffers hold m_cap entries; if this frame produced more, emit nothing,
    // raise the overflow flag (reserved[1]) and let the later passes see it (they skip too)
    if (info_dev != nullptr && (int64_t)info_dev->n_isect > m_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) info_dev->reserved[1] = 1u;
        return;
    }
    const int64_t N = n_dev ? (int64_t)(*n_dev) : N_host;  // compacted depth order: the count is on the device
    __syncthreads();
    const int top = tp.n - 1;
    const int top_shift = tp.n == 1 ? tp.shift[0] : (tp.n == 2 ? tp.shift[1] : (tp.n == 3 ? tp.shift[2] : tp.shift[3]));
    const uint32_t tw = (uint32_t)tiles_w;
    const uint32_t M = N > 0 ? __ldg(offsets + N) : 0u;  // total number of pairs (written by the count + scan kernel)
    // persistent warps: each takes several tasks; the CTA flushes its histograms once
    const int64_t warp_global = (int64_t)blockIdx.x * (kEmit2Threads / 32) + (tid >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * (kEmit2Threads / 32);
    for (int64_t task = warp_global; task * kEmitTask < (int64_t)M; task += warp_stride) {
        const uint32_t S = (uint32_t)(task * kEmitTask);
        const uint32_t E = min(M, S + kEmitTask);
        // owner of pair S: the largest j in [0, N) with offsets[j] <= S -- 32-ary search, all lanes probe
        int64_t lo = 0, hi = N;  // invariant: offsets[lo] <= S, answer in [lo, hi)
        while (hi - lo > 1) {
            const int64_t span = hi - lo;
            const int64_t step = (span + 31) / 32;
            const int64_t probe = lo + (int64_t)(lane + 1) * step;  // lane l probes lo + (l+1) step
            const bool le = probe < hi && __ldg(offsets + probe) <= S;
            const uint32_t m = __ballot_sync(0xffffffffu, le);  // offsets are non-decreasing: a prefix of lanes
            const int k = __popc(m);
            const int64_t new_lo = lo + (int64_t)k * step;
            const int64_t new_hi = min(hi, lo + (int64_t)(k + 1) * step);
            lo = new_lo; hi = new_hi;
        }
        int64_t jg0 = lo;       // first Gaussian of the current window
        uint32_t cur = S;       // next pair to write
        // the window after the current one is fetched while the current one is expanded (its three loads were the
        // kernel's largest stall: one task per warp leaves nothing else to hide them behind)
        int32_t g_nx = 0;
        uint2 rc_nx = make_uint2(0u, 0u);
        uint32_t off_nx = M;    // lanes past N own nothing
        auto fetch_window = [&](const int64_t first) {
            const int64_t jg = first + lane;
            g_nx = 0; rc_nx = make_uint2(0u, 0u); off_nx = M;
            if (jg < N) {
                g_nx = perm ? __ldg(perm + jg) : (int32_t)jg;
                rc_nx = __ldg(rects + jg);
                off_nx = __ldg(offsets + jg);
            }
        };
        fetch_window(jg0);
        while (cur < E) {
            const int32_t g = g_nx;
            const uint2 rc = rc_nx;
            const uint32_t off = off_nx;
            const bool have = jg0 + lane < N;
            fetch_window(jg0 + 32);
            const uint32_t w = rc.y & 0xffffu, cnt = w * (rc.y >> 16);
            // no Gaussian of the window is empty (always true under the torch rules and after a compaction): the owner
            // of a pair follows from ONE warp-wide OR of "my range starts at pair p0 + j" bits instead of a 5-step
            // chain of dependent shuffles
            const bool dense = __ballot_sync(0xffffffffu, have && cnt == 0u) == 0u;
            // kSep: the reciprocal is the integer magic of w (bit pattern carried in the same register)
            const float inv_w = kSep ? __uint_as_float(s_hist[0][(w ? w : 1u) - 1u]) : 1.0f / (float)(w ? w : 1u);
            // pairs covered by this window: [off of lane 0, end of the last live lane)
            const int n_live = (int)min((int64_t)32, N - jg0);
            const uint32_t win_end = __shfl_sync(0xffffffffu, off + cnt, n_live - 1);
            const uint32_t stop = min(E, win_end);
            for (uint32_t p0 = cur; p0 < stop; p0 += 32) {  // warp-uniform trip count
                const uint32_t pidx = p0 + lane;
                const bool valid = pidx < stop;
                // owner = largest lane l with off_l <= pidx (offsets are non-decreasing; off_{l+1} = off_l + cnt_l)
                uint32_t ol = 0;
                if (dense) {
                    // = (number of lanes whose range starts at or before pidx) - 1: those before p0 by ballot, those
                    // inside [p0, p0 + 32) by their start bits (distinct, because no range is empty)
                    const uint32_t rel = off - p0;
                    const uint32_t starts = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
                    const uint32_t before = __popc(__ballot_sync(0xffffffffu, off < p0));
                    ol = before + __popc(starts & (0xffffffffu >> (31u - lane))) - 1u;
                } else {
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const uint32_t cand = ol + step;
                        const uint32_t oc = __shfl_sync(0xffffffffu, off, cand & 31u);
                        if (oc <= pidx) ol = cand;
                    }
                }
                const uint32_t eo = __shfl_sync(0xffffffffu, off, ol);
                const uint32_t xy = __shfl_sync(0xffffffffu, rc.x, ol);
                const uint32_t wo = __shfl_sync(0xffffffffu, w, ol);
                const float iw = __shfl_sync(0xffffffffu, inv_w, ol);
                const int32_t go = __shfl_sync(0xffffffffu, g, ol);
                uint32_t tile = 0;
                if (kSep) {
                    if (valid) {
                        const uint32_t k = pidx - eo;
                        BSPLAT_DASSERT((int64_t)pidx < m_cap && pidx >= eo && wo >= 1u && wo <= (uint32_t)kRadix &&
                                       k / wo == (wo == 1u ? k : __umulhi(k, __float_as_uint(iw))) && (xy >> 16) + k / wo < 256u);
                        // k / w = hi(k * ceil(2^32 / w)), exact while k (w ceil(2^32 / w) - 2^32) < 2^32, i.e. for
                        // k w < 2^32: rectangles of at most 256 x 256 tiles have k < 2^16 and w <= 2^8
                        const uint32_t dy = wo == 1u ? k : __umulhi(k, __float_as_uint(iw));
                        const uint32_t dx = k - dy * wo;
                        tile_keys[pidx] = xy + (dy << 16) + dx;   // (row << 16 | column), the layout of rc.x
                        ids[pidx] = go;
                    }
                } else {
                    if (valid) {
                        const uint32_t k = pidx - eo;
                        // k / w through the reciprocal, corrected to be exact (k < 2^24: a rectangle has < 2^24
                        // tiles whenever the image has, which tile ids as 32-bit keys already require)
                        uint32_t dy = (uint32_t)__float2int_rz(__fmul_rn((float)k, iw));
                        if (dy * wo > k) --dy;
                        else if ((dy + 1u) * wo <= k) ++dy;
                        const uint32_t dx = k - dy * wo;
                        tile = ((xy >> 16) + dy) * tw + (xy & 0xffffu) + dx;
                        tile_keys[pidx] = tile;
                        ids[pidx] = go;
#pragma unroll
                        for (int q = 0; q < kMaxTilePasses - 1; ++q)  // (constant indices keep tp in the parameter bank)
                            if (q < top) atomicAdd(&s_hist[q][(tile >> tp.shift[q]) & ((1u << tp.bits[q]) - 1u)], 1u);
                    }
                    const uint32_t td = valid ? (tile >> top_shift) : 0x100u;
                    const uint32_t peers = __match_any_sync(0xffffffffu, td);
                    if (valid && lane == (uint32_t)(__ffs(peers) - 1))
                        atomicAdd(&s_hist[top][td], (uint32_t)__popc(peers));
                }
            }
            cur = stop;
            jg0 += 32;
            // (a window whose last Gaussian reaches beyond E ends the task; one that ends before E continues)
        }
    }
    pdl_trigger();
    if (kSep) return;
    __syncthreads();
    for (int i = tid; i < tp.n * kRadix; i += kEmit2Threads) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// ---- 5. per-tile counts -> tile ranges (+ heavy-first tile order) ------------------------------
// ranges[t] = [exclusive prefix of counts, + count) -- searchsorted-left semantics for empty tiles
// (binning.py:252-260).  order (optional) lists the tiles [first, first + n_order) by list length,
// longest first (counting sort on len/32, 256 buckets; order inside a bucket is arbitrary).
// Up to 128 CTAs, each owning a contiguous block of tiles, in two phases separated by a grid-wide flag (every
// CTA publishes its block total and its bucket histogram, then waits until all have): the one-CTA version spent
// 18 us issuing ~40 k warp instructions through a single SM.  The grid is small enough to be co-resident; if it
// is not at once (other streams), waiting CTAs only delay the rest, which depend on nobody.
constexpr int kFinThreads = 128;
constexpr int kFinMaxCtas = 128;
// scratch words (zeroed per frame): [0] arrivals, [1 .. 256] bucket sizes, [257 .. 512] bucket cursors,
// [513 .. 640] block totals, [700] spare counter (bin2_spare_counter)
constexpr int kFinScratchWords = 1024;

__device__ __forceinline__ int fin_bucket(const uint32_t cc) { return (int)(255u - min(255u, (cc + 31u) >> 5)); }

__global__ void __launch_bounds__(kFinThreads)
tile_finish2_kernel(const int n_tiles, const int first, const int n_order, const uint32_t* __restrict__ counts,
                    int32_t* __restrict__ ranges, int32_t* __restrict__ order, uint32_t* __restrict__ scratch,
                    const bsplat_bin_info* __restrict__ info_dev, bsplat_bin_info* __restrict__ info_host) {
    pdl_wait();  // (programmatic dependent launch: nothing of the predecessor is read before this)
    __shared__ int s_cnt[256];
    __shared__ int s_res[256];
    __shared__ uint32_t s_w[kFinThreads / 32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int per = (n_tiles + G - 1) / G;
    const int t0 = b * per, t1 = min(n_tiles, t0 + per);
    uint32_t* arrivals = scratch;
    uint32_t* g_bucket = scratch + 1;
    uint32_t* g_cursor = scratch + 257;
    uint32_t* g_total = scratch + 513;
    auto count_at = [&](int t) { return (t < t1 && counts) ? __ldg(counts + t) : 0u; };
    auto in_order = [&](int t) { return order && t < t1 && t >= first && t < first + n_order; };
    auto block_sum = [&](uint32_t v) -> uint32_t {  // every thread gets the sum over the CTA
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        __syncthreads();
        if (lane == 0) s_w[warp] = v;
        __syncthreads();
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kFinThreads / 32; ++w) s += s_w[w];
        return s;
    };
    for (int i = tid; i < 256; i += kFinThreads) s_cnt[i] = 0;
    // The frame's status (M, overflow flag) goes to the caller's pinned host block with a plain store over PCIe --
    // not a cudaMemcpyAsync: a 32-byte copy would queue in the device-to-host copy engine behind a 25 MB image
    // download of an earlier frame and stall the whole binning stream for its duration (measured: a 0.44 ms bubble
    // every third frame of the host-buffer pipeline).
    if (info_host != nullptr && b == 0 && tid < (int)(sizeof(bsplat_bin_info) / sizeof(uint32_t)))
        reinterpret_cast<volatile uint32_t*>(info_host)[tid] = reinterpret_cast<const uint32_t*>(info_dev)[tid];
    __syncthreads();
    // ---- phase 1: block total + bucket histogram of this CTA's tiles ----
    uint32_t mine = 0;
    for (int t = t0 + tid; t < t1; t += kFinThreads) {
        const uint32_t cc = count_at(t);
        mine += cc;
        if (in_order(t)) atomicAdd(&s_cnt[fin_bucket(cc)], 1);
    }
    const uint32_t total = block_sum(mine);
    if (tid == 0) st_relaxed_u32(g_total + b, total);
    if (order)
        for (int i = tid; i < 256; i += kFinThreads)
            if (s_cnt[i]) atomicAdd(g_bucket + i, (uint32_t)s_cnt[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(arrivals, 1u);
        unsigned ns = 32;
        while (ld_relaxed_u32(arrivals) < (uint32_t)G) {
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;
        }
        __threadfence();
    }
    __syncthreads();
    // ---- phase 2: ranges from the block prefix, order from the bucket prefix ----
    const uint32_t before = block_sum(tid < b ? ld_relaxed_u32(g_total + tid) : 0u);  // G <= 128 = kFinThreads
    if (order) {
        // exclusive prefix of the 256 global bucket sizes (two per thread), then this CTA's range in every bucket
        const uint32_t a0 = ld_relaxed_u32(g_bucket + 2 * tid), a1 = ld_relaxed_u32(g_bucket + 2 * tid + 1);
        uint32_t incl = a0 + a1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        __syncthreads();
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t excl = incl - (a0 + a1);
        for (int w = 0; w < warp; ++w) excl += s_w[w];
        const int c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
        s_res[2 * tid] = c0 ? (int)(excl + atomicAdd(g_cursor + 2 * tid, (uint32_t)c0)) : 0;
        s_res[2 * tid + 1] = c1 ? (int)(excl + a0 + atomicAdd(g_cursor + 2 * tid + 1, (uint32_t)c1)) : 0;
    }
    if (tid == 0) s_carry = before;
    __syncthreads();
    for (int tb = t0; tb < t1; tb += kFinThreads) {  // CTA-uniform trip count
        const int t = tb + tid;
        const uint32_t cc = count_at(t);
        uint32_t incl = cc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t excl = s_carry + incl - cc;
        for (int w = 0; w < warp; ++w) excl += s_w[w];
        if (t < t1) {
            ranges[2 * t] = (int32_t)excl;
            ranges[2 * t + 1] = (int32_t)(excl + cc);
            if (in_order(t)) {
                const int bk = fin_bucket(cc);
                order[s_res[bk] + atomicSub(&s_cnt[bk], 1) - 1] = t;
            }
        }
        __syncthreads();
        if (tid == kFinThreads - 1) s_carry = excl + cc;
        __syncthreads();
    }
}

int tile_finish_launch(int n_tiles, int first, int n_order, const uint32_t* counts, int32_t* ranges,
                       int32_t* order, uint32_t* scratch, const bsplat_bin_info* info_dev,
                       bsplat_bin_info* info_host, cudaStream_t stream) {
    const int by_size = (n_tiles + 63) / 64;
    const int G = by_size < 1 ? 1 : (by_size > kFinMaxCtas ? kFinMaxCtas : by_size);
    BSPLAT_LAUNCH_PDL((tile_finish2_kernel), G, kFinThreads, 0, stream, n_tiles, first, n_order, counts, ranges, order, scratch,
                                                       info_dev, info_host);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

// ---- workspace ------------------------------------------------------------------------------
constexpr size_t kBinAlign = 256;
static inline size_t bin_align(size_t v) { return (v + kBinAlign - 1) / kBinAlign * kBinAlign; }

static int tile_bits_of(int64_t n_tiles) {
    int tb = 1;
    while (((int64_t)1 << tb) < n_tiles) ++tb;
    return tb;
}

// ceil(bits / 8) passes of (nearly) equal width: 13 bits -> 7 + 6, 17 bits -> 6 + 6 + 5, never more than 8 per pass
// 2-D tile keys (row << 16 | column): used when the linear tile id needs two passes anyway and both coordinates fit
// a digit.  One pass per coordinate, columns first.
static bool tile_keys_2d(const BinParams& p) {
    return p.tiles_w <= kRadix && p.tiles_h <= kRadix && tile_bits_of((int64_t)p.tiles_w * p.tiles_h) > kRadixBits;
}
static TilePasses tile_passes_2d(const BinParams& p) {
    TilePasses tp;
    tp.n = 2;
    for (int q = 0; q < kMaxTilePasses; ++q) { tp.shift[q] = 0; tp.bits[q] = 0; }
    tp.shift[0] = 0; tp.bits[0] = tile_bits_of(p.tiles_w);
    tp.shift[1] = 16; tp.bits[1] = tile_bits_of(p.tiles_h);
    return tp;
}
static TilePasses tile_passes_of(int64_t n_tiles) {
    TilePasses tp;
    const int tb = tile_bits_of(n_tiles);
    tp.n = (tb + kRadixBits - 1) / kRadixBits;
    int shift = 0;
    for (int q = 0; q < kMaxTilePasses; ++q) {
        const int bits = q < tp.n ? tb / tp.n + (q < tb % tp.n ? 1 : 0) : 0;
        tp.shift[q] = shift;
        tp.bits[q] = bits;
        shift += bits;
    }
    return tp;
}

constexpr int kHistRows = 4 + kMaxTilePasses;   // 4 depth passes + the tile passes
constexpr int kTickets = 16;                    // [0..3] depth, [4..7] tile passes, [8] compaction, [9] band pre-test

struct Bin2Ws {
    // N part (lives from prepare to finish)
    uint32_t* dkeys; uint32_t* dkeys_alt; int32_t* perm; int32_t* perm_alt;
    uint32_t* offsets; uint2* rects; uint2* rects_in; bsplat_bin_info* info; void* scan_ws; size_t scan_bytes;
    uint32_t* hist;      // [kHistRows][256]
    uint32_t* tickets;   // [kTickets]
    unsigned long long* n_band;  // [0] Gaussians with >= 1 tile in the band, [1] with >= 1 tile in the frame
    unsigned long long* compact_status;
    int32_t* cand;                      // row-band pre-test: candidate indices (index order), [N]
    unsigned long long* cand_n;         // [0] number of candidates
    unsigned long long* cand_status;    // pre-test compaction: chunk counts, then chunk offsets (uint32 each)
    uint8_t* cand_flags;                // one flag byte per thread of the pre-test compaction
    uint8_t* compact_flags;             // ... and of the exact compaction
    uint32_t* status_n;  // [4][tilesN][256]
    size_t zero_begin, zero_end_n;  // byte range zeroed by begin
    size_t n_bytes;
    // M part
    uint32_t* tkeys; uint32_t* tkeys_alt; int32_t* ids; int32_t* ids_alt;
    uint32_t* status_m;     // [tile passes][tilesM][256]
    uint32_t* tile_counts;  // [n_tiles], directly behind status_m (zeroed together)
    uint32_t* fin_scratch;  // [kFinScratchWords], directly behind tile_counts
    size_t status_m_off, status_m_bytes, zero_m_bytes;
    size_t total;
};

static Bin2Ws carve_bin2(void* base, int64_t N, int64_t M, int64_t n_tiles) {
    Bin2Ws w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += bin_align(bytes); return r; };
    const size_t n = (size_t)(N > 0 ? N : 1), m = (size_t)(M > 0 ? M : 1);
    const size_t nt = (size_t)(n_tiles > 0 ? n_tiles : 1);
    w.dkeys = (uint32_t*)take(n * 4); w.dkeys_alt = (uint32_t*)take(n * 4);
    w.perm = (int32_t*)take(n * 4); w.perm_alt = (int32_t*)take(n * 4);
    w.offsets = (uint32_t*)take((n + 1) * 4);
    w.rects = (uint2*)take(n * sizeof(uint2));
    w.rects_in = (uint2*)take(n * sizeof(uint2));
    w.cand = (int32_t*)take(n * 4);
    w.cand_flags = (uint8_t*)take((size_t)ceil_div((int64_t)n, kCompactChunk) * kCompactThreads);
    w.compact_flags = (uint8_t*)take((size_t)ceil_div((int64_t)n, kCompactChunk) * kCompactThreads);
    w.zero_begin = off;
    w.info = (bsplat_bin_info*)take(sizeof(bsplat_bin_info));
    w.scan_bytes = bsplat_bin_scan_workspace_bytes(N);
    w.scan_ws = take(w.scan_bytes);
    w.hist = (uint32_t*)take((size_t)kHistRows * kRadix * 4);
    w.tickets = (uint32_t*)take(kTickets * 4);
    w.n_band = (unsigned long long*)take(16);
    w.compact_status = (unsigned long long*)take((size_t)(ceil_div((int64_t)n, kCompactChunk) + 1) * 8);
    w.cand_n = (unsigned long long*)take(16);
    w.cand_status = (unsigned long long*)take((size_t)(ceil_div((int64_t)n, kCompactChunk) + 1) * 8);
    w.status_n = (uint32_t*)take(sort_status_words(sort_tiles_u32(N), 4) * 4);  // (the wide depth passes use fewer)
    w.zero_end_n = off;
    w.n_bytes = off;
    w.tkeys = (uint32_t*)take(m * 4); w.tkeys_alt = (uint32_t*)take(m * 4);
    w.ids = (int32_t*)take(m * 4); w.ids_alt = (int32_t*)take(m * 4);
    w.status_m_off = off;
    w.status_m_bytes = sort_status_words(sort_tiles_u32(M), tile_passes_of(n_tiles).n) * 4;
    // status_m, tile_counts and the finish scratch must be contiguous (single memset): take them as one block
    w.zero_m_bytes = w.status_m_bytes + (nt + kFinScratchWords) * sizeof(uint32_t);
    w.status_m = (uint32_t*)take(w.zero_m_bytes);
    w.tile_counts = w.status_m ? w.status_m + w.status_m_bytes / sizeof(uint32_t) : nullptr;
    w.fin_scratch = w.tile_counts ? w.tile_counts + nt : nullptr;
    w.total = off;
    return w;
}

static inline float inv_tile_of(const BinParams& p) {
    const int ts = (int)p.tile_size_f;
    return ((ts & (ts - 1)) == 0) ? 1.0f / p.tile_size_f : 0.0f;  // exact only for powers of two
}

static inline bool is_band(const BinParams& p) { return p.row_begin > 0 || p.row_end < p.tiles_h; }

// Zeroes the per-frame control words; must precede whatever produces the histograms (the projection epilogue or
// bin_prep_kernel).
int bin2_begin(int64_t N, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    if (!workspace || workspace_bytes < w.n_bytes) return BSPLAT_E_WORKSPACE;
    char* base = static_cast<char*>(workspace);
    BSPLAT_CUDA_TRY(cudaMemsetAsync(base + w.zero_begin, 0, w.zero_end_n - w.zero_begin, stream));
    return BSPLAT_OK;
}

// Where the producer of step 0 writes: rectangles in index order, depth keys (the compaction input when the
// frame is compacted, the sort input otherwise) and -- only when nothing is compacted -- the digit histograms.
void bin2_prep_targets(void* workspace, int64_t N, bool compact, uint2** rects_in, uint32_t** dkeys,
                       uint32_t** hist) {
    Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    *rects_in = w.rects_in;
    *dkeys = compact ? w.dkeys_alt : w.dkeys;
    *hist = compact ? nullptr : w.hist;
}

// have_prep: bin2_begin ran and step 0 is done (fused frames); otherwise both happen here from the stage inputs.
// compact: sort only the Gaussians that own a tile of the band (always for partial bands).
// Row-band frames under the torch rules: list of the Gaussians whose tile rectangle can reach the band at all
// (conservative pre-test on the raw inputs, BandPretest), in index order, count on the device.  Call between bin2_begin
// and the projection; hand the list to the projection (ProjExtra::list) and to bin2_prepare (use_candidates).
int bin2_band_candidates(int64_t N, const float* means3d, const float* log_scales, const bsplat_camera& cam,
                         float eps2d, const BinParams& p, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream, const int32_t** list, const unsigned long long** list_n) {
    Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    if (!workspace || workspace_bytes < w.n_bytes) return BSPLAT_E_WORKSPACE;
    if (p.semantics != BSPLAT_SEM_TORCH || !is_band(p) || N <= 0) return BSPLAT_E_ARG;
    BandPretest pt;
    double a[3][3];  // Rv^T Rv: its largest eigenvalue (<= the largest absolute row sum) bounds |Rv|_2^2
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            a[r][c] = 0.0;
            for (int k = 0; k < 3; ++k) a[r][c] += (double)cam.viewmat[4 * k + r] * (double)cam.viewmat[4 * k + c];
        }
    double rv2 = 0.0;
    for (int r = 0; r < 3; ++r) {
        const double rs = fabs(a[r][0]) + fabs(a[r][1]) + fabs(a[r][2]);
        rv2 = rs > rv2 ? rs : rv2;
    }
    for (int k = 0; k < 3; ++k) { pt.r1[k] = cam.viewmat[4 + k]; pt.r2[k] = cam.viewmat[8 + k]; }
    pt.t1 = cam.viewmat[7]; pt.t2 = cam.viewmat[11];
    pt.fy = cam.fy; pt.cy = cam.cy;
    const float Hf = (float)cam.height, tan_fovy = 0.5f * Hf / cam.fy;
    const float lim_pos = (Hf - cam.cy) / cam.fy + 0.3f * tan_fovy, lim_neg = cam.cy / cam.fy + 0.3f * tan_fovy;
    const float lim = fmaxf(fabsf(lim_pos), fabsf(lim_neg));
    pt.jn2 = (1.0f + lim * lim) * 1.001f;
    pt.gain2 = (float)(9.0 * rv2 * 1.001);
    pt.eps2d = eps2d;
    pt.y_lo = p.row_begin <= 0 ? -INFINITY : (float)p.row_begin * p.tile_size_f - 1.0f;
    pt.y_hi = p.row_end >= p.tiles_h ? INFINITY : (float)p.row_end * p.tile_size_f + 1.0f;
    const int64_t cb = ceil_div(N, kCompactChunk);
    uint32_t* counts = reinterpret_cast<uint32_t*>(w.cand_status);
    BSPLAT_LAUNCH_PDL((bin_compact_flags_kernel<true>), (unsigned)cb, kCompactThreads, 0, stream, N, nullptr, nullptr, nullptr, p.row_begin, p.row_end, means3d, log_scales, pt, w.cand_flags, counts, counts + cb,
        w.tickets + 9, w.cand_n, 0ull);
    BSPLAT_LAUNCH_CHECK();
    BSPLAT_LAUNCH_PDL((bin_compact_scatter_kernel<true>), (unsigned)(cb < 148 * 8 ? cb : 148 * 8), kCompactThreads, 0, stream, cb, nullptr, w.cand_flags, counts + cb, nullptr, nullptr, w.cand, nullptr);
    BSPLAT_LAUNCH_CHECK();
    *list = w.cand;
    *list_n = w.cand_n;
    return BSPLAT_OK;
}

int bin2_prepare(int64_t N, const float* means2d, const void* radii, int radii_is_float, const float* depths,
                 const BinParams& p, void* workspace, size_t workspace_bytes, cudaStream_t stream, bool have_prep,
                 bool compact, bool use_candidates) {
    Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    if (!workspace || workspace_bytes < w.n_bytes) return BSPLAT_E_WORKSPACE;
    compact = compact || is_band(p);
    int rc = BSPLAT_OK;
    if (!have_prep) {
        rc = bin2_begin(N, workspace, workspace_bytes, stream);
        if (rc != BSPLAT_OK) return rc;
    }
    if (N == 0) {
        BSPLAT_CUDA_TRY(cudaMemsetAsync(w.offsets, 0, sizeof(uint32_t), stream));
        return BSPLAT_OK;
    }
    if (!have_prep) {
        const int64_t hb = ceil_div(N, 256 * 8);
        const unsigned grid = (unsigned)(hb < 148 * 4 ? hb : 148 * 4);
        bin_prep_kernel<<<grid, 256, 0, stream>>>(N, means2d, radii, radii_is_float, depths, p, inv_tile_of(p),
                                                  w.rects_in, compact ? w.dkeys_alt : w.dkeys,
                                                  compact ? nullptr : w.hist);
        BSPLAT_LAUNCH_CHECK();
    }
    const int64_t tn = sort_tiles_u32_depth(N);
    const uint64_t* n_dev = nullptr;
    const int32_t* vsrc = nullptr;
    if (compact) {
        const int64_t cb = ceil_div(N, kCompactChunk);
        // use_candidates: only the pre-tested Gaussians were projected; under the torch rules every Gaussian owns a
        // tile somewhere in the frame, so the whole-frame count is N
        uint32_t* counts = reinterpret_cast<uint32_t*>(w.compact_status);
        const int32_t* list = use_candidates ? w.cand : nullptr;
        BSPLAT_LAUNCH_PDL((bin_compact_flags_kernel<false>), (unsigned)cb, kCompactThreads, 0, stream, N, list, use_candidates ? w.cand_n : nullptr, w.rects_in, p.row_begin, p.row_end, nullptr, nullptr,
            BandPretest(), w.compact_flags, counts, counts + cb, w.tickets + 8, w.n_band,
            use_candidates ? (unsigned long long)N : 0ull);
        BSPLAT_LAUNCH_CHECK();
        BSPLAT_LAUNCH_PDL((bin_compact_scatter_kernel<false>), (unsigned)(cb < 148 * 4 ? cb : 148 * 4), kCompactThreads, 0, stream, cb, list, w.compact_flags, counts + cb, w.dkeys_alt, w.dkeys, w.perm, w.hist);
        BSPLAT_LAUNCH_CHECK();
        n_dev = reinterpret_cast<const uint64_t*>(w.n_band);
        vsrc = w.perm;
    }
    const uint32_t* ksrc = w.dkeys; uint32_t* kdst = w.dkeys_alt;
    int32_t* vdst = w.perm_alt;
    for (int pass = 0; pass < 4; ++pass) {
        // (the depth histograms come from the kernel in front of pass 0: passes 1-3 may read theirs early)
        rc = onesweep_pass_u32(N, n_dev, ksrc, pass == 3 ? nullptr : kdst, vsrc, vdst, 8 * pass, 8,
                               w.hist + (size_t)pass * kRadix, 0, w.tickets + pass,
                               w.status_n + (size_t)pass * tn * kRadix, nullptr, stream, 0, pass > 0, 1);
        if (rc != BSPLAT_OK) return rc;
        // ping-pong: pass 0 writes (dkeys_alt, perm_alt), pass 1 (dkeys, perm), ...; pass 3 ends in perm
        ksrc = kdst; kdst = (kdst == w.dkeys_alt) ? w.dkeys : w.dkeys_alt;
        vsrc = vdst; vdst = (vdst == w.perm_alt) ? w.perm : w.perm_alt;
    }
    // after 4 passes the sorted permutation is in w.perm (passes 1 and 3 write w.perm)
    const unsigned grid = (unsigned)ceil_div(N, kScan2Chunk);
    BSPLAT_LAUNCH_PDL((bin_count_scan2_kernel), grid, kScan2Threads, 0, stream, N, reinterpret_cast<const unsigned long long*>(n_dev), w.perm, w.rects_in, p.row_begin, p.row_end, w.offsets,
        w.info, static_cast<unsigned long long*>(w.scan_ws), w.rects, tile_keys_2d(p) ? w.hist + 4 * kRadix : nullptr,
        p.tiles_w);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

// device_m: M is a capacity; the real count is read on the device from the bin info (n_isect) written by
// prepare -- no host read-back between prepare and finish (sync-free / graph-capturable frames).
// info_host (optional): device-accessible pinned host block that receives the frame's bin info from the last kernel.
int bin2_finish(int64_t N, int64_t M, bool device_m, const BinParams& p, void* workspace, size_t workspace_bytes,
                int32_t* sorted_ids, int32_t* tile_ranges, int32_t* tile_order, cudaStream_t stream, bool compact,
                bsplat_bin_info* info_host) {
    const int n_tiles = p.tiles_w * p.tiles_h;
    Bin2Ws w = carve_bin2(workspace, N, M, n_tiles);
    if (!workspace || workspace_bytes < w.total) return BSPLAT_E_WORKSPACE;
    compact = compact || is_band(p);
    bsplat_bin_info* info_dev = device_m ? w.info : nullptr;
    const uint64_t* m_dev = device_m ? reinterpret_cast<const uint64_t*>(w.info) : nullptr;
    // status words of the tile passes, the per-tile counters and the finish scratch are contiguous: one memset
    BSPLAT_CUDA_TRY(cudaMemsetAsync(w.status_m, 0, w.zero_m_bytes, stream));
    const int first = p.row_begin * p.tiles_w, n_order = (p.row_end - p.row_begin) * p.tiles_w;
    if (M == 0) {
        if (!tile_order) {
            BSPLAT_CUDA_TRY(cudaMemsetAsync(tile_ranges, 0, (size_t)n_tiles * 2 * sizeof(int32_t), stream));
            return BSPLAT_OK;
        }
        return tile_finish_launch(n_tiles, first, n_order, nullptr, tile_ranges, tile_order, w.fin_scratch, w.info,
                                  info_host, stream);
    }
    const bool keys2d = tile_keys_2d(p);
    const TilePasses tp = keys2d ? tile_passes_2d(p) : tile_passes_of(n_tiles);
    // one warp per task of kEmitTask pairs; M is the capacity in sync-free frames
    const int64_t emit_ctas = ceil_div(ceil_div(M, (int64_t)kEmitTask), kEmit2Threads / 32);
    const unsigned emit_grid = (unsigned)(emit_ctas < 148 * 8 ? emit_ctas : 148 * 8);
    if (keys2d)
        BSPLAT_LAUNCH_PDL((bin_emit2_kernel<true>), emit_grid, kEmit2Threads, 0, stream, N, compact ? w.n_band : nullptr, w.perm, w.rects, p.tiles_w, w.offsets, tp, w.tkeys, w.ids,
            w.hist + 4 * kRadix, info_dev, M);
    else
        BSPLAT_LAUNCH_PDL((bin_emit2_kernel<false>), emit_grid, kEmit2Threads, 0, stream, N, compact ? w.n_band : nullptr, w.perm, w.rects, p.tiles_w, w.offsets, tp, w.tkeys, w.ids,
            w.hist + 4 * kRadix, info_dev, M);
    BSPLAT_LAUNCH_CHECK();
    const int64_t tm = sort_tiles_u32(M);
    const uint32_t* ksrc = w.tkeys; uint32_t* kdst = w.tkeys_alt;
    const int32_t* vsrc = w.ids; int32_t* vdst = w.ids_alt;
    for (int q = 0; q < tp.n; ++q) {
        const bool last = q == tp.n - 1;
        // last pass: sorted tile ids are not written; per-tile counts come out of the pass instead
        const int rc = onesweep_pass_u32(M, m_dev, ksrc, last ? nullptr : kdst, vsrc, last ? sorted_ids : vdst,
                                         tp.shift[q], tp.bits[q], w.hist + (size_t)(4 + q) * kRadix, 0,
                                         w.tickets + 4 + q, w.status_m + (size_t)q * tm * kRadix,
                                         last ? w.tile_counts : nullptr, stream, keys2d ? p.tiles_w : 0,
                                         // 2-D keys: histograms from count + scan (two kernels back); linear keys:
                                         // from the emitter, the kernel in front of pass 0
                                         (keys2d || q > 0) ? 1 : 0);
        if (rc != BSPLAT_OK) return rc;
        ksrc = kdst; kdst = (kdst == w.tkeys_alt) ? w.tkeys : w.tkeys_alt;
        vsrc = vdst; vdst = (vdst == w.ids_alt) ? w.ids : w.ids_alt;
    }
    return tile_finish_launch(n_tiles, first, n_order, w.tile_counts, tile_ranges, tile_order, w.fin_scratch, w.info,
                              info_host, stream);
}

size_t bin2_workspace_bytes(int64_t N, int64_t M, int64_t n_tiles) { return carve_bin2(nullptr, N, M, n_tiles).total; }
// depth-ordered list of the compacted frame's Gaussians + their count (device)
void bin2_band_list(void* workspace, int64_t N, const int32_t** perm, const unsigned long long** n_band) {
    const Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    *perm = w.perm;
    *n_band = w.n_band;
}
bsplat_bin_info* bin2_info_ptr(void* workspace, int64_t N) { return carve_bin2(workspace, N, 0, 0).info; }
// one spare word of the finish scratch: zero from bin2_finish's memset on (grid barrier of the rasterizer's pre-pass)
uint32_t* bin2_spare_counter(void* workspace, int64_t N, int64_t M, int64_t n_tiles) {
    return carve_bin2(workspace, N, M, n_tiles).fin_scratch + 700;
}

}  // namespace bsplat

using namespace bsplat;

extern "C" size_t bsplat_bin2_workspace_bytes(int64_t N, int64_t M_capacity, int64_t n_tiles) {
    if (N < 0 || M_capacity < 0 || n_tiles < 0) return 0;
    return bin2_workspace_bytes(N, M_capacity, n_tiles);
}

// Phase 1: rectangles + depth keys, depth-sort the Gaussians, count + scan in depth order. Afterwards *info_out
// (device, 32 bytes) holds M; the caller reads it back (the stage's single read-back), sizes
// sorted_ids and calls phase 2 with the same workspace.
extern "C" int bsplat_bin2_prepare(int64_t N, const float* means2d, const void* radii, int32_t radii_is_float,
                                   const float* depths, int32_t width, int32_t height, int32_t tile_size,
                                   int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics,
                                   void* workspace, size_t workspace_bytes, bsplat_bin_info* info_out,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics & 0xff, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || !info_out) return BSPLAT_E_ARG;
    if (N > 0 && (!means2d || !radii || !depths)) return BSPLAT_E_ARG;
    rc = bin2_prepare(N, means2d, radii, radii_is_float, depths, p, workspace, workspace_bytes, stream, false,
                      (semantics & BSPLAT_BIN_PACKED) != 0, false);
    if (rc != BSPLAT_OK) return rc;
    const bsplat_bin_info* src = bin2_info_ptr(workspace, N);
    if (info_out != src)
        BSPLAT_CUDA_TRY(cudaMemcpyAsync(info_out, src, sizeof(bsplat_bin_info), cudaMemcpyDeviceToDevice, stream));
    return BSPLAT_OK;
}

// Phase 2: emit in depth order, stable sort by tile id, tile ranges.
extern "C" int bsplat_bin2_finish(int64_t N, int64_t M, const float* means2d, const void* radii,
                                  int32_t radii_is_float, int32_t width, int32_t height, int32_t tile_size,
                                  int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics,
                                  void* workspace, size_t workspace_bytes, int32_t* sorted_ids,
                                  int32_t* tile_ranges, int32_t* tile_order, void* stream_) {
    (void)means2d; (void)radii; (void)radii_is_float;  // phase 1 kept everything phase 2 needs in the workspace
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics & 0xff, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || M < 0 || !tile_ranges) return BSPLAT_E_ARG;
    if (M >= (int64_t)kStatMask) return BSPLAT_E_OVERFLOW;
    if (M > 0 && !sorted_ids) return BSPLAT_E_ARG;
    return bin2_finish(N, M, false, p, workspace, workspace_bytes, sorted_ids, tile_ranges, tile_order, stream,
                       (semantics & BSPLAT_BIN_PACKED) != 0, nullptr);
}
