// Stage 2 -- binning: per-Gaussian tile rectangles, device-wide exclusive scan, key/id emission
// and tile-range extraction.  (The sort itself is radix_sort.cu.)
//
// Replaces mojosplat/binning.py:108-262 (torch backend; bit-exact target) and the gsplat pair
// isect_tiles / isect_offset_encode (binning.py:41-102).  The reference walks the Gaussians in a
// Python loop and sorts twice; here
//   1. bin_count_scan_kernel  fuses the rect computation with a single-pass decoupled-look-back
//      prefix sum (reads 20 B / Gaussian, writes 4 B),
//   2. bin_emit_kernel        writes (key, id) pairs with block-cooperative, perfectly coalesced
//      stores (a binary search over the block's 256 offsets finds the owner of each output slot),
//   3. tile_ranges_kernel     turns the sorted keys into [start, end) per tile, with
//      searchsorted-left semantics for empty tiles.
// Key = (tile_id << depth_bits) | (depth_key - depth_bias): only live bits are ever sorted.
#include "binning.cuh"

namespace bsplat {

__global__ void __launch_bounds__(kScanThreads)
bin_count_scan_kernel(const int64_t N_host, const unsigned long long* __restrict__ n_dev,
                      const int32_t* __restrict__ perm, const float* __restrict__ means2d,
                      const void* __restrict__ radii, const int radii_is_float,
                      const float* __restrict__ depths, const BinParams p,
                      uint32_t* __restrict__ offsets, bsplat_bin_info* __restrict__ info,
                      unsigned long long* __restrict__ ws, uint2* __restrict__ rects) {
    __shared__ unsigned int s_chunk;
    __shared__ unsigned long long s_warp_sum[kScanThreads / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_red[2][kScanThreads / 32];

    const int tid = threadIdx.x;
    // n_dev: the number of items lives on the device (band-compacted depth order); the grid is sized by N_host
    const int64_t N = n_dev ? (int64_t)(*n_dev) : N_host;
    if (tid == 0) s_chunk = atomicAdd(reinterpret_cast<unsigned int*>(ws), 1u);
    __syncthreads();
    const unsigned int chunk = s_chunk;
    unsigned long long* status = ws + 1;
    const int64_t base = (int64_t)chunk * kScanChunk;
    if (base >= N) {  // surplus chunk (tickets are handed out in order: no lower chunk ever waits for this one)
        if (N == 0 && chunk == 0 && tid == 0) { offsets[0] = 0u; info->n_isect = 0ull; }
        return;
    }

    // blocked arrangement: thread t owns items base + t*kScanItems + k (keeps the scan trivial)
    uint32_t cnt[kScanItems];
    uint32_t kmin_inv = 0u, kmax = 0u;  // max(~key) and max(key); both start at 0
    uint32_t thread_sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t jj = base + (int64_t)tid * kScanItems + k;
        uint32_t c = 0;
        if (jj < N) {
            // two-level path: slot jj of the depth-sorted order holds Gaussian perm[jj]
            const int64_t i = perm ? (int64_t)__ldg(perm + jj) : jj;
            float mx, my, rx, ry;
            load_mean_radii(means2d, radii, radii_is_float, i, mx, my, rx, ry);
            const TileRect r = tile_rect(mx, my, rx, ry, p.W, p.H, p.tile_size_f, p.tiles_w, p.tiles_h,
                                         p.semantics, p.row_begin, p.row_end);
            c = (uint32_t)((r.x1 - r.x0) * (r.y1 - r.y0));
            // two-level path: keep the rectangle for the emitter (x0 | y0 << 16, w | h << 16), in depth order
            if (rects != nullptr)
                rects[jj] = make_uint2((uint32_t)r.x0 | ((uint32_t)r.y0 << 16),
                                       (uint32_t)(r.x1 - r.x0) | ((uint32_t)(r.y1 - r.y0) << 16));
            if (c > 0 && depths != nullptr) {  // depth-key range: only the single-level key layout needs it
                const uint32_t dk = depth_key(__ldg(depths + i));
                kmax = max(kmax, dk);
                kmin_inv = max(kmin_inv, ~dk);
            }
        }
        cnt[k] = c;
        thread_sum += c;
    }

    // block-wide exclusive scan of thread sums (64-bit: a block can exceed 2^32 only in theory,
    // but the running total across chunks can)
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long incl = thread_sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    // min/max key reduction
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        kmin_inv = max(kmin_inv, __shfl_xor_sync(0xffffffffu, kmin_inv, d));
    }
    if (lane == 0) { s_red[0][warp] = kmax; s_red[1][warp] = kmin_inv; }
    __syncthreads();
    unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const unsigned long long s = s_warp_sum[w];
        if (w < warp) warp_excl += s;
        block_total += s;
    }
    const unsigned long long thread_excl = warp_excl + incl - thread_sum;

    // decoupled look-back over previous chunks: warp 0 inspects 32 predecessors per round trip
    if (warp == 0) {
        unsigned long long prefix = 0;
        if (chunk == 0) {
            if (lane == 0) st_relaxed_u64(status + 0, kFlagPrefix | block_total);
        } else {
            if (lane == 0) st_relaxed_u64(status + chunk, kFlagAgg | block_total);
            int64_t j = (int64_t)chunk - 1;
            while (true) {
                const int64_t idx = j - lane;  // lane 0 = nearest predecessor
                const unsigned long long v = idx >= 0 ? ld_relaxed_u64(status + idx) : kFlagPrefix;
                const unsigned ready = __ballot_sync(0xffffffffu, (v & ~kValueMask) != 0);
                const unsigned pref = __ballot_sync(0xffffffffu, (v & kFlagPrefix) != 0);
                const unsigned first_pref = pref ? (unsigned)(__ffs(pref) - 1) : 32u;
                const unsigned need = first_pref == 32u ? 0xffffffffu : ((2u << first_pref) - 1u);
                if ((ready & need) != need) { __nanosleep(64); continue; }  // not published yet: polite re-poll
                unsigned long long contrib = ((unsigned)lane <= first_pref) ? (v & kValueMask) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                prefix += contrib;
                if (first_pref != 32u) break;
                j -= 32;
            }
            if (lane == 0) st_relaxed_u64(status + chunk, kFlagPrefix | (prefix + block_total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            uint32_t bmax = 0, bmin_inv = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; ++w) {
                bmax = max(bmax, s_red[0][w]);
                bmin_inv = max(bmin_inv, s_red[1][w]);
            }
            if (block_total > 0) {
                atomicMax(&info->max_depth_key, bmax);
                atomicMax(&info->reserved[0], bmin_inv);  // ~min, finalised after the scan
            }
        }
    }
    __syncthreads();
    unsigned long long run = s_prefix + thread_excl;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)tid * kScanItems + k;
        if (i < N) offsets[i] = (uint32_t)run;
        run += cnt[k];
    }
    // the chunk that owns the last Gaussian publishes the total
    if (tid == kScanThreads - 1 && base + kScanChunk >= N) {
        offsets[N] = (uint32_t)run;
        info->n_isect = run;
    }
}

// min_depth_key = ~max(~key); run after the scan (stream order) so every atomicMax has landed.
__global__ void bin_finalize_info_kernel(bsplat_bin_info* info) {
    info->min_depth_key = ~info->reserved[0];
    if (info->n_isect == 0) { info->min_depth_key = 0; info->max_depth_key = 0; }
}

// ------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
bin_emit_kernel(const int64_t N, const float* __restrict__ means2d, const void* __restrict__ radii,
                const int radii_is_float, const float* __restrict__ depths, const BinParams p,
                const uint32_t* __restrict__ offsets, const uint32_t depth_bias, const int depth_bits,
                uint64_t* __restrict__ keys, int32_t* __restrict__ ids) {
    __shared__ uint32_t s_off[kEmitThreads + 1];
    __shared__ uint32_t s_xy[kEmitThreads];   // x0 | y0 << 16
    __shared__ uint32_t s_w[kEmitThreads];    // rect width in tiles
    __shared__ uint32_t s_dk[kEmitThreads];   // biased depth key

    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kEmitThreads;
    const int n_here = (int)min((int64_t)kEmitThreads, N - base);
    if (tid < n_here) {
        const int64_t i = base + tid;
        float mx, my, rx, ry;
        load_mean_radii(means2d, radii, radii_is_float, i, mx, my, rx, ry);
        const TileRect r = tile_rect(mx, my, rx, ry, p.W, p.H, p.tile_size_f, p.tiles_w, p.tiles_h,
                                     p.semantics, p.row_begin, p.row_end);
        s_xy[tid] = (uint32_t)r.x0 | ((uint32_t)r.y0 << 16);
        s_w[tid] = (uint32_t)(r.x1 - r.x0);
        s_dk[tid] = depth_key(__ldg(depths + i)) - depth_bias;
        s_off[tid] = __ldg(offsets + i);
    }
    if (tid == 0) s_off[n_here] = __ldg(offsets + base + n_here);
    __syncthreads();

    const uint32_t begin = s_off[0], end = s_off[n_here];
    for (uint32_t pos = begin + tid; pos < end; pos += kEmitThreads) {
        // largest j with s_off[j] <= pos  (upper_bound - 1); zero-count Gaussians are skipped
        int lo = 0, hi = n_here;  // invariant: s_off[lo] <= pos < s_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= pos) lo = mid; else hi = mid;
        }
        const uint32_t k = pos - s_off[lo];
        const uint32_t w = s_w[lo];
        const uint32_t dy = k / w, dx = k - dy * w;
        const uint32_t xy = s_xy[lo];
        const uint32_t tile = ((xy >> 16) + dy) * (uint32_t)p.tiles_w + (xy & 0xffffu) + dx;
        keys[pos] = ((uint64_t)tile << depth_bits) | (uint64_t)s_dk[lo];
        ids[pos] = (int32_t)(base + lo);
    }
}

// ------------------------------------------------------------------------------------------
// tile_ranges[t] = [first index with tile >= t, first index with tile >= t+1)
template <typename KeyT>
__global__ void tile_ranges_kernel(const int64_t M, const KeyT* __restrict__ keys, const int tile_shift,
                                   const int n_tiles, int32_t* __restrict__ ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > M) return;
    const int64_t cur = (i < M) ? (int64_t)(keys[i] >> tile_shift) : (int64_t)n_tiles;
    const int64_t prev = (i > 0) ? (int64_t)(keys[i - 1] >> tile_shift) : -1;
    for (int64_t t = prev + 1; t <= cur; ++t) {
        if (t < n_tiles) ranges[2 * t] = (int32_t)i;
        if (t >= 1) ranges[2 * (t - 1) + 1] = (int32_t)i;
    }
}

// workspace and info must be zeroed by the caller (stream-ordered memset)
int bin_count_scan_launch(int64_t N, const int32_t* perm, const float* means2d, const void* radii,
                          int radii_is_float, const float* depths, const BinParams& p, uint32_t* offsets,
                          bsplat_bin_info* info, void* workspace, bool finalize_key_range, cudaStream_t stream,
                          uint2* rects, const unsigned long long* n_dev) {
    const unsigned grid = (unsigned)ceil_div(N, kScanChunk);
    bin_count_scan_kernel<<<grid, kScanThreads, 0, stream>>>(N, n_dev, perm, means2d, radii, radii_is_float,
                                                             finalize_key_range ? depths : nullptr, p, offsets, info,
                                                             static_cast<unsigned long long*>(workspace), rects);
    BSPLAT_LAUNCH_CHECK();
    if (finalize_key_range) {  // only the single-level path needs min/max depth keys
        bin_finalize_info_kernel<<<1, 1, 0, stream>>>(info);
        BSPLAT_LAUNCH_CHECK();
    }
    return BSPLAT_OK;
}

int tile_ranges_u32_launch(int64_t M, const uint32_t* sorted_tile_ids, int n_tiles, int32_t* tile_ranges,
                           cudaStream_t stream) {
    const int threads = 256;
    const unsigned grid = (unsigned)ceil_div(M + 1, threads);
    tile_ranges_kernel<uint32_t><<<grid, threads, 0, stream>>>(M, sorted_tile_ids, 0, n_tiles, tile_ranges);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

int make_bin_params(int32_t width, int32_t height, int32_t tile_size, int32_t row_begin,
                       int32_t row_end, int32_t semantics, BinParams* p) {
    if (width <= 0 || height <= 0 || tile_size <= 0) return BSPLAT_E_ARG;
    if (semantics != BSPLAT_SEM_TORCH && semantics != BSPLAT_SEM_GSPLAT) return BSPLAT_E_ARG;
    p->W = width; p->H = height;
    p->tiles_w = (width + tile_size - 1) / tile_size;
    p->tiles_h = (height + tile_size - 1) / tile_size;
    if (p->tiles_w > 65535 || p->tiles_h > 65535) return BSPLAT_E_ARG;
    p->semantics = semantics;
    p->row_begin = row_begin < 0 ? 0 : row_begin;
    p->row_end = row_end > p->tiles_h ? p->tiles_h : row_end;
    if (p->row_end < p->row_begin) return BSPLAT_E_ARG;
    p->tile_size_f = (float)tile_size;
    return BSPLAT_OK;
}

}  // namespace bsplat

using namespace bsplat;

extern "C" size_t bsplat_bin_scan_workspace_bytes(int64_t N) {
    const int64_t chunks = ceil_div(N > 0 ? N : 1, kScanChunk);
    return (size_t)(chunks + 1) * sizeof(unsigned long long);
}

extern "C" int bsplat_bin_count_scan(int64_t N, const float* means2d, const void* radii,
                                     int32_t radii_is_float, const float* depths, int32_t width,
                                     int32_t height, int32_t tile_size, int32_t tile_row_begin,
                                     int32_t tile_row_end, int32_t semantics, uint32_t* offsets,
                                     bsplat_bin_info* info, void* workspace, size_t workspace_bytes,
                                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || !offsets || !info) return BSPLAT_E_ARG;
    if (N > 0 && (!means2d || !radii || !depths)) return BSPLAT_E_ARG;
    const size_t need = bsplat_bin_scan_workspace_bytes(N);
    if (!workspace || workspace_bytes < need) return BSPLAT_E_WORKSPACE;
    BSPLAT_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
    BSPLAT_CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(bsplat_bin_info), stream));
    if (N == 0) {
        BSPLAT_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(uint32_t), stream));
        return BSPLAT_OK;
    }
    const unsigned grid = (unsigned)ceil_div(N, kScanChunk);
    return bin_count_scan_launch(N, nullptr, means2d, radii, radii_is_float, depths, p, offsets, info, workspace,
                                 true, stream, nullptr, nullptr);
}

extern "C" bsplat_key_layout bsplat_make_key_layout(const bsplat_bin_info* info_host, int32_t width,
                                                    int32_t height, int32_t tile_size) {
    bsplat_key_layout L;
    const int64_t tiles_w = (width + tile_size - 1) / tile_size;
    const int64_t tiles_h = (height + tile_size - 1) / tile_size;
    const int64_t n_tiles = tiles_w * tiles_h;
    int tb = 1;
    while (((int64_t)1 << tb) < n_tiles) ++tb;
    L.tile_bits = tb;
    L.depth_bias = info_host->min_depth_key;
    const uint32_t span = info_host->max_depth_key - info_host->min_depth_key;
    int db = 1;
    while (db < 32 && (span >> db) != 0) ++db;
    L.depth_bits = db;
    return L;
}

extern "C" int bsplat_bin_emit(int64_t N, const float* means2d, const void* radii,
                               int32_t radii_is_float, const float* depths, int32_t width,
                               int32_t height, int32_t tile_size, int32_t tile_row_begin,
                               int32_t tile_row_end, int32_t semantics, const uint32_t* offsets,
                               bsplat_key_layout layout, uint64_t* keys, int32_t* ids, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    BinParams p;
    int rc = make_bin_params(width, height, tile_size, tile_row_begin, tile_row_end, semantics, &p);
    if (rc != BSPLAT_OK) return rc;
    if (N < 0 || !offsets) return BSPLAT_E_ARG;
    if (layout.depth_bits < 1 || layout.depth_bits > 32 || layout.tile_bits < 1 ||
        layout.depth_bits + layout.tile_bits > 64)
        return BSPLAT_E_ARG;
    if (N == 0) return BSPLAT_OK;
    if (!means2d || !radii || !depths || !keys || !ids) return BSPLAT_E_ARG;
    const unsigned grid = (unsigned)ceil_div(N, kEmitThreads);
    bin_emit_kernel<<<grid, kEmitThreads, 0, stream>>>(N, means2d, radii, radii_is_float, depths, p,
                                                       offsets, layout.depth_bias, layout.depth_bits,
                                                       keys, ids);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

extern "C" int bsplat_tile_ranges(int64_t M, const uint64_t* sorted_keys, int32_t tile_shift,
                                  int32_t n_tiles, int32_t* tile_ranges, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M < 0 || n_tiles <= 0 || !tile_ranges || tile_shift < 0 || tile_shift > 63) return BSPLAT_E_ARG;
    if (M > 0 && !sorted_keys) return BSPLAT_E_ARG;
    const int threads = 256;
    const unsigned grid = (unsigned)ceil_div(M + 1, threads);
    tile_ranges_kernel<uint64_t><<<grid, threads, 0, stream>>>(M, sorted_keys, tile_shift, n_tiles, tile_ranges);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}
