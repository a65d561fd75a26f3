// Stage 2 (default path) -- two-level binning.
//
// The reference sorts the (gaussian, tile) pairs twice: argsort by depth, then a STABLE argsort by
// tile id (mojosplat/binning.py:223-231).  This path keeps that structure but moves the depth sort to
// where it is cheap:
//   1. depth-sort the N Gaussians          4 onesweep passes over (uint32 depth key, index) -- N items
//   2. count + scan in depth order         tile rects of Gaussian perm[j], exclusive prefix sum, M
//   3. emit in depth order                 (tile id, gaussian id) pairs; tile-digit histograms on the fly
//   4. stable sort by tile id only         ceil(log2(n_tiles)) bits -> 2 onesweep passes over M items
//   5. tile ranges                         from the sorted tile ids
// Result = ascending (tile, depth, gaussian index): bit-identical to the single-level sort of
// (tile << depth_bits | depth_key) keys and to the reference's lists (canonical tie order),
// with ~4x less sort traffic: M-scale data is moved by 2 passes of 8 B pairs instead of 6 passes of 12 B.
#include "binning.cuh"
#include "radix_sort.cuh"

namespace bsplat {

// ---- 1. depth keys + digit histograms of the 4 depth passes ---------------------------------
__global__ void __launch_bounds__(256)
depth_key_hist_kernel(const int64_t N, const float* __restrict__ depths, uint32_t* __restrict__ keys,
                      uint32_t* __restrict__ hist /* [4][256] */) {
    __shared__ uint32_t s_hist[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t warp_start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    for (int64_t wi = warp_start; wi < N; wi += stride) {  // warp-uniform trip count
        const int64_t i = wi + lane;
        const bool valid = i < N;
        uint32_t k = 0;
        if (valid) {
            k = depth_key(__ldg(depths + i));
            keys[i] = k;
            atomicAdd(&s_hist[0][k & 0xffu], 1u);
            atomicAdd(&s_hist[1][(k >> 8) & 0xffu], 1u);
            atomicAdd(&s_hist[2][(k >> 16) & 0xffu], 1u);
        }
        // sign + exponent bits: a handful of distinct values per warp -> aggregate before the atomic
        const uint32_t top = valid ? (k >> 24) : 0x100u;
        const uint32_t peers = __match_any_sync(0xffffffffu, top);
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[3][top], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// ---- 1b. row-band frames: only the Gaussians that reach the band are depth-sorted -----------------------
// flag = 0 for a Gaussian whose tile rectangle intersects the band, 1 otherwise; a 1-bit onesweep pass on the
// flags is a stable partition of the indices (in-band first, ascending index), band_gather_hist_kernel then
// collects the depth keys of the in-band prefix.  Every rank of a row-band split used to sort all N Gaussians.
__global__ void __launch_bounds__(256)
band_flag_kernel(const int64_t N, const float* __restrict__ means2d, const void* __restrict__ radii,
                 const int radii_is_float, const float* __restrict__ depths, const BinParams p,
                 uint32_t* __restrict__ flags, uint32_t* __restrict__ keys_full, uint32_t* __restrict__ hist2,
                 unsigned long long* __restrict__ n_band) {
    __shared__ unsigned int s_in;
    if (threadIdx.x == 0) s_in = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int mine = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        float mx, my, rx, ry;
        load_mean_radii(means2d, radii, radii_is_float, i, mx, my, rx, ry);
        const TileRect r = tile_rect(mx, my, rx, ry, p.W, p.H, p.tile_size_f, p.tiles_w, p.tiles_h, p.semantics,
                                     p.row_begin, p.row_end);
        const bool in = (r.x1 > r.x0) && (r.y1 > r.y0);
        flags[i] = in ? 0u : 1u;
        keys_full[i] = depth_key(__ldg(depths + i));
        mine += in ? 1u : 0u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_in, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_in) {
            atomicAdd(hist2 + 0, s_in);
            atomicAdd(n_band, (unsigned long long)s_in);
        }
    }
}

// hist2[1] = N - hist2[0]; one thread, stream-ordered behind band_flag_kernel.
__global__ void band_fix_hist_kernel(const int64_t N, uint32_t* __restrict__ hist2) {
    hist2[1] = (uint32_t)N - hist2[0];
}

__global__ void __launch_bounds__(256)
band_gather_hist_kernel(const unsigned long long* __restrict__ n_band, const int32_t* __restrict__ perm0,
                        const uint32_t* __restrict__ keys_full, uint32_t* __restrict__ keys,
                        uint32_t* __restrict__ hist /* [4][256] */) {
    __shared__ uint32_t s_hist[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int64_t n = (int64_t)(*n_band);
    const uint32_t lane = lane_id();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t warp_start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    for (int64_t wi = warp_start; wi < n; wi += stride) {  // warp-uniform trip count
        const int64_t j = wi + lane;
        const bool valid = j < n;
        uint32_t k = 0;
        if (valid) {
            k = __ldg(keys_full + __ldg(perm0 + j));
            keys[j] = k;
            atomicAdd(&s_hist[0][k & 0xffu], 1u);
            atomicAdd(&s_hist[1][(k >> 8) & 0xffu], 1u);
            atomicAdd(&s_hist[2][(k >> 16) & 0xffu], 1u);
        }
        const uint32_t top = valid ? (k >> 24) : 0x100u;
        const uint32_t peers = __match_any_sync(0xffffffffu, top);
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[3][top], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * kRadix; i += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// ---- 3. emission in depth order ------------------------------------------------------------
// Work is balanced over OUTPUT PAIRS, not over Gaussians: the M pairs are cut into tasks of kEmitTask
// consecutive pairs, a warp takes a task, finds the Gaussian that owns its first pair with a 32-ary search over the
// exclusive offsets (4 probes for 1 M Gaussians) and then walks the task 32 pairs at a time: it keeps a window of
// 32 depth-consecutive Gaussians (rectangles and offsets come from the count + scan kernel, coalesced), the owner
// of pair p is found with a 5-step shuffle binary search over the window's offsets, its rectangle fetched with
// shuffles, the tile id computed and both words stored -- every store instruction writes 128 contiguous bytes,
// no lane idles on a short rectangle, and a Gaussian that covers thousands of tiles (the nearest ones, which the
// depth order puts side by side) is shared by as many warps as it has tasks.  The digit histograms of the two
// tile-sort passes are accumulated on the fly (shared-memory atomics; the high digit is run-length aggregated with
// match.any: neighbouring pairs share it).
constexpr int kEmit2Threads = 256;
constexpr uint32_t kEmitTask = 1024;

__global__ void __launch_bounds__(kEmit2Threads)
bin_emit2_kernel(const int64_t N_host, const unsigned long long* __restrict__ n_dev,
                 const int32_t* __restrict__ perm, const uint2* __restrict__ rects,
                 const int tiles_w, const uint32_t* __restrict__ offsets, const int lo_bits,
                 uint32_t* __restrict__ tile_keys, int32_t* __restrict__ ids, uint32_t* __restrict__ hist /* [2][256] */,
                 bsplat_bin_info* __restrict__ info_dev, const int64_t m_cap) {
    // sync-free frames: the pair buffers hold m_cap entries; if this frame produced more, emit nothing,
    // raise the overflow flag (reserved[1]) and let the later passes see it (they skip too)
    if (info_dev != nullptr && (int64_t)info_dev->n_isect > m_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) info_dev->reserved[1] = 1u;
        return;
    }
    __shared__ uint32_t s_hist[2][kRadix];
    const int64_t N = n_dev ? (int64_t)(*n_dev) : N_host;  // band-compacted order: the count is on the device
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    for (int i = tid; i < 2 * kRadix; i += kEmit2Threads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t lo_mask = (1u << lo_bits) - 1u;
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
        while (cur < E) {
            const int64_t jg = jg0 + lane;
            const bool have = jg < N;
            int32_t g = 0;
            uint2 rc = make_uint2(0u, 0u);
            uint32_t off = M;   // lanes past N own nothing
            if (have) {
                g = __ldg(perm + jg);
                rc = __ldg(rects + jg);
                off = __ldg(offsets + jg);
            }
            const uint32_t w = rc.y & 0xffffu, cnt = w * (rc.y >> 16);
            const float inv_w = 1.0f / (float)(w ? w : 1u);
            // pairs covered by this window: [off of lane 0, end of the last live lane)
            const int n_live = (int)min((int64_t)32, N - jg0);
            const uint32_t win_end = __shfl_sync(0xffffffffu, off + cnt, n_live - 1);
            const uint32_t stop = min(E, win_end);
            for (uint32_t p0 = cur; p0 < stop; p0 += 32) {  // warp-uniform trip count
                const uint32_t pidx = p0 + lane;
                const bool valid = pidx < stop;
                // owner = largest lane l with off_l <= pidx (offsets are non-decreasing; off_{l+1} = off_l + cnt_l)
                uint32_t ol = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const uint32_t cand = ol + step;
                    const uint32_t oc = __shfl_sync(0xffffffffu, off, cand & 31u);
                    if (oc <= pidx) ol = cand;
                }
                const uint32_t eo = __shfl_sync(0xffffffffu, off, ol);
                const uint32_t xy = __shfl_sync(0xffffffffu, rc.x, ol);
                const uint32_t wo = __shfl_sync(0xffffffffu, w, ol);
                const float iw = __shfl_sync(0xffffffffu, inv_w, ol);
                const int32_t go = __shfl_sync(0xffffffffu, g, ol);
                uint32_t tile = 0;
                if (valid) {
                    const uint32_t k = pidx - eo;
                    // k / w through the reciprocal, corrected to be exact (k < 2^24: a rectangle has < 2^24 tiles
                    // whenever the image has, which tile ids as 32-bit keys already require)
                    uint32_t dy = (uint32_t)__float2int_rz(__fmul_rn((float)k, iw));
                    if (dy * wo > k) --dy;
                    else if ((dy + 1u) * wo <= k) ++dy;
                    const uint32_t dx = k - dy * wo;
                    tile = ((xy >> 16) + dy) * tw + (xy & 0xffffu) + dx;
                    tile_keys[pidx] = tile;
                    ids[pidx] = go;
                    atomicAdd(&s_hist[0][tile & lo_mask], 1u);
                }
                const uint32_t top = valid ? (tile >> lo_bits) : 0x100u;
                const uint32_t peers = __match_any_sync(0xffffffffu, top);
                if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[1][top], (uint32_t)__popc(peers));
            }
            cur = stop;
            jg0 += 32;
            // (a window whose last Gaussian reaches beyond E ends the task; one that ends before E continues)
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * kRadix; i += kEmit2Threads) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// ---- 5. per-tile counts -> tile ranges (+ heavy-first tile order), one CTA ---------------------
// ranges[t] = [exclusive prefix of counts, + count) -- searchsorted-left semantics for empty tiles
// (binning.py:252-260).  order (optional) lists the tiles [first, first + n_order) by list length,
// longest first (counting sort on len/32, 256 buckets; order inside a bucket is arbitrary).
constexpr int kFinishThreads = 1024;
constexpr int kFinishMaxPer = 32;  // tiles per thread held in registers (covers 32 768 tiles = 4K at 16 px)

// One CTA.  Warp w owns the consecutive tiles [w * 32 per, (w + 1) * 32 per) and walks them 32 at a time (lane =
// consecutive tile): every load and store of the kernel is coalesced.  (The first version gave each THREAD
// consecutive tiles: 32 sectors per warp instruction through the one SM's LSU -- 18 us for 8 160 tiles, all of it
// lg-throttle stalls.)  Counts are read once into registers, scanned per row of 32 (shuffles) with a running
// carry, warp totals are scanned across the block, ranges and bucket positions come from the registers.
__global__ void __launch_bounds__(kFinishThreads)
tile_finish_kernel(const int n_tiles, const int first, const int n_order, const uint32_t* __restrict__ counts,
                   int32_t* __restrict__ ranges, int32_t* __restrict__ order) {
    __shared__ uint32_t s_warp[kFinishThreads / 32];
    __shared__ int s_cnt[256];
    __shared__ int s_base[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 256) s_cnt[tid] = 0;
    const int per = (n_tiles + kFinishThreads - 1) / kFinishThreads;  // rows of 32 tiles per warp
    const int w0 = warp * per * 32;
    const bool in_regs = per <= kFinishMaxPer;
    auto count_at = [&](int t) { return (t < n_tiles && counts) ? __ldg(counts + t) : 0u; };
    uint32_t c[kFinishMaxPer];   // this lane's count in row k
    uint32_t ex[kFinishMaxPer];  // exclusive prefix inside the warp's tiles
    uint32_t carry = 0;
    auto row_scan = [&](uint32_t v, uint32_t& excl) {
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        excl = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    };
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < kFinishMaxPer; ++k) c[k] = (k < per) ? count_at(w0 + k * 32 + lane) : 0u;
#pragma unroll
        for (int k = 0; k < kFinishMaxPer; ++k)
            if (k < per) row_scan(c[k], ex[k]);
    } else {
        for (int k = 0; k < per; ++k) { uint32_t e; row_scan(count_at(w0 + k * 32 + lane), e); }
    }
    if (lane == 0) s_warp[warp] = carry;  // total of the warp's tiles
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const uint32_t warp_base = s_warp[warp];
    // bucket counters are hit by thousands of tiles with a handful of distinct lengths: aggregate per warp
    // (match.any on the bucket) so that one lane per distinct bucket issues the shared-memory atomic
    auto bucket_add = [&](bool take, uint32_t cc, int* table) -> int {
        const int bkt = take ? (int)(255u - min(255u, (cc + 31u) >> 5)) : 256;
        const unsigned peers = __match_any_sync(0xffffffffu, bkt);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (take && lane == leader) base = atomicAdd(&table[bkt], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        return base + __popc(peers & ((1u << lane) - 1u));
    };
    auto in_order = [&](int t) { return order && t < n_tiles && t >= first && t < first + n_order; };
    auto put_range = [&](int k, uint32_t cc, uint32_t e) {  // warp-uniform call sites (match.any inside)
        const int t = w0 + k * 32 + lane;
        if (t < n_tiles) {
            ranges[2 * t] = (int32_t)(warp_base + e);
            ranges[2 * t + 1] = (int32_t)(warp_base + e + cc);
        }
        if (order) bucket_add(in_order(t), cc, s_cnt);
    };
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < kFinishMaxPer; ++k)
            if (k < per) put_range(k, c[k], ex[k]);
    } else {
        carry = 0;
        for (int k = 0; k < per; ++k) {
            const uint32_t cc = count_at(w0 + k * 32 + lane);
            uint32_t e;
            row_scan(cc, e);
            put_range(k, cc, e);
        }
    }
    if (!order) return;
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 256 bucket sizes, 8 per lane
        int local[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { local[k] = tot; tot += s_cnt[lane * 8 + k]; }
        int wi = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s_base[lane * 8 + k] = wi - tot + local[k];
    }
    __syncthreads();
    auto put_order = [&](int k, uint32_t cc) {
        const int t = w0 + k * 32 + lane;
        const bool take = in_order(t);
        const int pos = bucket_add(take, cc, s_base);
        if (take) order[pos] = t;
    };
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < kFinishMaxPer; ++k)
            if (k < per) put_order(k, c[k]);
    } else {
        for (int k = 0; k < per; ++k) put_order(k, count_at(w0 + k * 32 + lane));
    }
}

int tile_finish_launch(int n_tiles, int first, int n_order, const uint32_t* counts, int32_t* ranges,
                       int32_t* order, cudaStream_t stream) {
    tile_finish_kernel<<<1, kFinishThreads, 0, stream>>>(n_tiles, first, n_order, counts, ranges, order);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

// ---- workspace ------------------------------------------------------------------------------
constexpr size_t kBinAlign = 256;
static inline size_t bin_align(size_t v) { return (v + kBinAlign - 1) / kBinAlign * kBinAlign; }

struct Bin2Ws {
    // N part (lives from prepare to finish)
    uint32_t* dkeys; uint32_t* dkeys_alt; int32_t* perm; int32_t* perm_alt;
    uint32_t* offsets; uint2* rects; bsplat_bin_info* info; void* scan_ws; size_t scan_bytes;
    uint32_t* hist;      // [7][256]: 4 depth passes + 2 tile passes + band partition
    uint32_t* tickets;   // [8]
    unsigned long long* n_band;  // in-band Gaussians (row-band frames)
    uint32_t* status_n;  // [5][tilesN][256]: 4 depth passes + band partition
    size_t zero_begin, zero_end_n;  // byte range zeroed by prepare
    size_t n_bytes;
    // M part
    uint32_t* tkeys; uint32_t* tkeys_alt; int32_t* ids; int32_t* ids_alt;
    uint32_t* status_m;  // [2][tilesM][256]
    uint32_t* tile_counts;  // [n_tiles], directly behind status_m (zeroed together)
    size_t status_m_off, status_m_bytes;
    size_t total;
};

static Bin2Ws carve_bin2(void* base, int64_t N, int64_t M, int64_t n_tiles) {
    Bin2Ws w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += bin_align(bytes); return r; };
    const size_t n = (size_t)(N > 0 ? N : 1), m = (size_t)(M > 0 ? M : 1);
    w.dkeys = (uint32_t*)take(n * 4); w.dkeys_alt = (uint32_t*)take(n * 4);
    w.perm = (int32_t*)take(n * 4); w.perm_alt = (int32_t*)take(n * 4);
    w.offsets = (uint32_t*)take((n + 1) * 4);
    w.rects = (uint2*)take(n * sizeof(uint2));
    w.zero_begin = off;
    w.info = (bsplat_bin_info*)take(sizeof(bsplat_bin_info));
    w.scan_bytes = bsplat_bin_scan_workspace_bytes(N);
    w.scan_ws = take(w.scan_bytes);
    w.hist = (uint32_t*)take(7 * kRadix * 4);
    w.tickets = (uint32_t*)take(8 * 4);
    w.n_band = (unsigned long long*)take(16);
    w.status_n = (uint32_t*)take(sort_status_words(sort_tiles_u32(N), 5) * 4);
    w.zero_end_n = off;
    w.n_bytes = off;
    w.tkeys = (uint32_t*)take(m * 4); w.tkeys_alt = (uint32_t*)take(m * 4);
    w.ids = (int32_t*)take(m * 4); w.ids_alt = (int32_t*)take(m * 4);
    w.status_m_off = off;
    w.status_m_bytes = sort_status_words(sort_tiles_u32(M), 2) * 4;
    // status_m and tile_counts must be contiguous (single memset): take them as one block
    w.status_m = (uint32_t*)take(w.status_m_bytes + (size_t)(n_tiles > 0 ? n_tiles : 1) * sizeof(uint32_t));
    w.tile_counts = w.status_m ? w.status_m + w.status_m_bytes / sizeof(uint32_t) : nullptr;
    w.total = off;
    return w;
}

static int tile_bits_of(const BinParams& p) {
    const int64_t n_tiles = (int64_t)p.tiles_w * p.tiles_h;
    int tb = 1;
    while (((int64_t)1 << tb) < n_tiles) ++tb;
    return tb;
}

int bin2_prepare(int64_t N, const float* means2d, const void* radii, int radii_is_float, const float* depths,
                 const BinParams& p, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    if (!workspace || workspace_bytes < w.n_bytes) return BSPLAT_E_WORKSPACE;
    char* base = static_cast<char*>(workspace);
    BSPLAT_CUDA_TRY(cudaMemsetAsync(base + w.zero_begin, 0, w.zero_end_n - w.zero_begin, stream));
    if (N == 0) {
        BSPLAT_CUDA_TRY(cudaMemsetAsync(w.offsets, 0, sizeof(uint32_t), stream));
        return BSPLAT_OK;
    }
    int rc = BSPLAT_OK;
    const int64_t tn = sort_tiles_u32(N);
    const int64_t hb = ceil_div(N, 256 * 8);
    const unsigned hist_grid = (unsigned)(hb < 148 * 2 ? hb : 148 * 2);
    // A band that is not the whole image: partition the indices first, sort only the in-band prefix (device count)
    const bool band = p.row_begin > 0 || p.row_end < p.tiles_h;
    const uint64_t* n_dev = nullptr;
    const int32_t* vsrc = nullptr;
    if (band) {
        uint32_t* flags = w.offsets;                               // dead until the count + scan kernel
        uint32_t* keys_full = reinterpret_cast<uint32_t*>(w.rects);  // dead until the count + scan kernel
        uint32_t* hist2 = w.hist + 6 * kRadix;
        band_flag_kernel<<<hist_grid * 2, 256, 0, stream>>>(N, means2d, radii, radii_is_float, depths, p, flags,
                                                            keys_full, hist2, w.n_band);
        BSPLAT_LAUNCH_CHECK();
        band_fix_hist_kernel<<<1, 1, 0, stream>>>(N, hist2);
        BSPLAT_LAUNCH_CHECK();
        rc = onesweep_pass_u32(N, nullptr, flags, nullptr, nullptr, w.perm, 0, 1, hist2, 0, w.tickets + 6,
                               w.status_n + (size_t)4 * tn * kRadix, nullptr, stream);
        if (rc != BSPLAT_OK) return rc;
        band_gather_hist_kernel<<<hist_grid, 256, 0, stream>>>(w.n_band, w.perm, keys_full, w.dkeys, w.hist);
        BSPLAT_LAUNCH_CHECK();
        n_dev = reinterpret_cast<const uint64_t*>(w.n_band);
        vsrc = w.perm;
    } else {
        depth_key_hist_kernel<<<hist_grid, 256, 0, stream>>>(N, depths, w.dkeys, w.hist);
        BSPLAT_LAUNCH_CHECK();
    }
    const uint32_t* ksrc = w.dkeys; uint32_t* kdst = w.dkeys_alt;
    int32_t* vdst = w.perm_alt;
    for (int pass = 0; pass < 4; ++pass) {
        rc = onesweep_pass_u32(N, n_dev, ksrc, pass == 3 ? nullptr : kdst, vsrc, vdst, 8 * pass, 8,
                               w.hist + (size_t)pass * kRadix, 0, w.tickets + pass,
                               w.status_n + (size_t)pass * tn * kRadix, nullptr, stream);
        if (rc != BSPLAT_OK) return rc;
        // ping-pong: pass 0 writes (dkeys_alt, perm_alt), pass 1 (dkeys, perm), ...; pass 3 ends in perm
        ksrc = kdst; kdst = (kdst == w.dkeys_alt) ? w.dkeys : w.dkeys_alt;
        vsrc = vdst; vdst = (vdst == w.perm_alt) ? w.perm : w.perm_alt;
    }
    // after 4 passes the sorted permutation is in w.perm (passes 1 and 3 write w.perm)
    return bin_count_scan_launch(N, w.perm, means2d, radii, radii_is_float, depths, p, w.offsets, w.info,
                                 w.scan_ws, /*finalize_key_range=*/false, stream, w.rects,
                                 reinterpret_cast<const unsigned long long*>(n_dev));
}

// device_m: M is a capacity; the real count is read on the device from the bin info (n_isect) written by
// prepare -- no host read-back between prepare and finish (sync-free / graph-capturable frames).
int bin2_finish(int64_t N, int64_t M, bool device_m, const float* means2d, const void* radii, int radii_is_float,
                const BinParams& p, void* workspace, size_t workspace_bytes, int32_t* sorted_ids,
                int32_t* tile_ranges, int32_t* tile_order, cudaStream_t stream) {
    const int n_tiles = p.tiles_w * p.tiles_h;
    Bin2Ws w = carve_bin2(workspace, N, M, n_tiles);
    if (!workspace || workspace_bytes < w.total) return BSPLAT_E_WORKSPACE;
    bsplat_bin_info* info_dev = device_m ? w.info : nullptr;
    const uint64_t* m_dev = device_m ? reinterpret_cast<const uint64_t*>(w.info) : nullptr;
    if (M == 0) {
        BSPLAT_CUDA_TRY(cudaMemsetAsync(tile_ranges, 0, (size_t)n_tiles * 2 * sizeof(int32_t), stream));
        if (tile_order)
            return tile_finish_launch(n_tiles, p.row_begin * p.tiles_w, (p.row_end - p.row_begin) * p.tiles_w,
                                      nullptr, tile_ranges, tile_order, stream);
        return BSPLAT_OK;
    }
    // status words of the two passes and the per-tile counters are contiguous: one memset
    BSPLAT_CUDA_TRY(cudaMemsetAsync(w.status_m, 0, w.status_m_bytes + (size_t)n_tiles * sizeof(uint32_t), stream));
    const int tb = tile_bits_of(p);
    const int lo_bits = tb > 8 ? (tb + 1) / 2 : tb;  // split the tile id evenly over <= 2 passes
    const int hi_bits = tb - lo_bits;
    // one warp per task of kEmitTask pairs; M is the capacity in sync-free frames
    const int64_t emit_ctas = ceil_div(ceil_div(M, (int64_t)kEmitTask), kEmit2Threads / 32);
    const bool band = p.row_begin > 0 || p.row_end < p.tiles_h;  // prepare compacted the depth order to the band
    bin_emit2_kernel<<<(unsigned)(emit_ctas < 148 * 8 ? emit_ctas : 148 * 8), kEmit2Threads, 0, stream>>>(
        N, band ? w.n_band : nullptr, w.perm, w.rects, p.tiles_w, w.offsets, lo_bits, w.tkeys, w.ids,
        w.hist + 4 * kRadix, info_dev, M);
    BSPLAT_LAUNCH_CHECK();
    int rc = BSPLAT_OK;
    const int64_t tm = sort_tiles_u32(M);
    if (hi_bits > 0) {
        rc = onesweep_pass_u32(M, m_dev, w.tkeys, w.tkeys_alt, w.ids, w.ids_alt, 0, lo_bits, w.hist + 4 * kRadix,
                               0, w.tickets + 4, w.status_m, nullptr, stream);
        if (rc != BSPLAT_OK) return rc;
        // last pass: sorted tile ids are not written; per-tile counts come out of the pass instead
        rc = onesweep_pass_u32(M, m_dev, w.tkeys_alt, nullptr, w.ids_alt, sorted_ids, lo_bits, hi_bits,
                               w.hist + 5 * kRadix, 0, w.tickets + 5, w.status_m + (size_t)tm * kRadix,
                               w.tile_counts, stream);
    } else {
        rc = onesweep_pass_u32(M, m_dev, w.tkeys, nullptr, w.ids, sorted_ids, 0, lo_bits, w.hist + 4 * kRadix,
                               0, w.tickets + 4, w.status_m, w.tile_counts, stream);
    }
    if (rc != BSPLAT_OK) return rc;
    return tile_finish_launch(n_tiles, p.row_begin * p.tiles_w, (p.row_end - p.row_begin) * p.tiles_w,
                              w.tile_counts, tile_ranges, tile_order, stream);
}

size_t bin2_workspace_bytes(int64_t N, int64_t M, int64_t n_tiles) { return carve_bin2(nullptr, N, M, n_tiles).total; }
void bin2_band_list(void* workspace, int64_t N, const int32_t** perm, const unsigned long long** n_band) {
    const Bin2Ws w = carve_bin2(workspace, N, 0, 0);
    *perm = w.perm;
    *n_band = w.n_band;
}
bsplat_bin_info* bin2_info_ptr(void* workspace, int64_t N) { return carve_bin2(workspace, N, 0, 0).info; }

}  // namespace bsplat

using namespace bsplat;

extern "C" size_t bsplat_bin2_workspace_bytes(int64_t N, int64_t M_capacity, int64_t n_tiles) {
    if (N < 0 || M_capacity < 0 || n_tiles < 0) return 0;
    return bin2_workspace_bytes(N, M_capacity, n_tiles);
}

// Phase 1: depth-sort the Gaussians, tile rects + prefix sum in depth order. Afterwards *info_out
// (device, 32 bytes) holds M; the caller reads it back (the stage's single read-back), sizes
// sorted_ids and calls phase 2 with the same workspace.
extern "C" int bsplat_bin2_prepare(int64_t N, const float* means2d, const void* radii, int32_t radii_is_float,
                                   const float* depths, int32_t width, int32_t height, int32_t tile_size,
                                   int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics,
                                   void* workspace, size_t workspace_bytes, bsplat_bin_info* info_out,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || !info_out) return BSPLAT_E_ARG;
    if (N > 0 && (!means2d || !radii || !depths)) return BSPLAT_E_ARG;
    rc = bin2_prepare(N, means2d, radii, radii_is_float, depths, p, workspace, workspace_bytes, stream);
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
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || M < 0 || !tile_ranges) return BSPLAT_E_ARG;
    if (M >= (int64_t)kStatMask) return BSPLAT_E_OVERFLOW;
    if (M > 0 && (!means2d || !radii || !sorted_ids)) return BSPLAT_E_ARG;
    return bin2_finish(N, M, false, means2d, radii, radii_is_float, p, workspace, workspace_bytes, sorted_ids,
                       tile_ranges, tile_order, stream);
}
