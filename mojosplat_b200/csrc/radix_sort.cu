// Stage 2b -- onesweep least-significant-digit radix sort of (uint64 key, int32 id) pairs.
//
// Replaces the two sorts of the reference's binning (mojosplat/binning.py:223-231: argsort by
// depth, then stable argsort by tile id) -- and gsplat's cub::DeviceRadixSort inside isect_tiles
// (binning.py:73-82) -- with one stable sort over the packed key
// (tile_id << depth_bits) | depth_key, restricted to the live bits [begin_bit, end_bit).
//
// Structure (Adinets & Merrill, "Onesweep"):
//   radix_histogram_kernel  one read of the keys builds the digit histograms of ALL passes
//   radix_scan_kernel       exclusive scan of each 256-bin histogram
//   onesweep_kernel         per pass: each CTA takes a tile of 3072 pairs in ticket order, ranks
//                           them stably by digit (warp match-any multisplit), resolves its global
//                           digit offsets with a decoupled look-back chain (one chain per digit,
//                           one thread per digit) and scatters through shared memory so that
//                           global stores are contiguous runs per digit.
// Per pass the kernel moves 12 B in + 12 B out per pair; nothing else touches HBM.
#include "common.cuh"

namespace bsplat {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 12;
constexpr int kSortTile = kSortThreads * kSortItems;  // 3072 pairs per CTA
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;

constexpr uint32_t kStatAgg = 1u << 30;
constexpr uint32_t kStatPrefix = 2u << 30;
constexpr uint32_t kStatMask = (1u << 30) - 1;

struct SortWs {
    uint32_t* hist;     // [P][256]
    uint32_t* tickets;  // [256] (P used)
    uint32_t* status;   // [P][n_tiles][256]
    size_t bytes;
};

static SortWs carve_ws(void* ws, int64_t M, int passes) {
    const int64_t n_tiles = ceil_div(M > 0 ? M : 1, kSortTile);
    SortWs w;
    uint32_t* p = static_cast<uint32_t*>(ws);
    w.hist = p;
    w.tickets = p + (size_t)kMaxPasses * kRadix;
    w.status = w.tickets + kRadix;
    w.bytes = ((size_t)kMaxPasses * kRadix + kRadix + (size_t)passes * n_tiles * kRadix) * sizeof(uint32_t);
    return w;
}

__global__ void __launch_bounds__(256)
radix_histogram_kernel(const int64_t M, const uint64_t* __restrict__ keys, const int begin_bit,
                       const int end_bit, const int passes, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    // warp-uniform trip count so that match.any sees full warps
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t warp_start = start - lane;
    for (int64_t wi = warp_start; wi < M; wi += stride) {
        const int64_t i = wi + lane;
        const bool valid = i < M;
        const uint64_t key = valid ? __ldg(keys + i) : 0ull;
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + p * kRadixBits;
            const int bits = min(kRadixBits, end_bit - shift);
            const uint32_t d = (uint32_t)(key >> shift) & ((1u << bits) - 1u);
            const uint32_t dm = valid ? d : 0x100u;
            const uint32_t peers = __match_any_sync(0xffffffffu, dm);
            if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[p][d], __popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

__global__ void __launch_bounds__(kRadix) radix_scan_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_warp[kRadix / 32];
    uint32_t* h = hist + (size_t)blockIdx.x * kRadix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t v = h[tid];
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    h[tid] = off + incl - v;
}

__global__ void __launch_bounds__(kSortThreads)
onesweep_kernel(const int64_t M, const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                const int32_t* __restrict__ vals_in, int32_t* __restrict__ vals_out, const int shift,
                const int bits, const uint32_t* __restrict__ bins, uint32_t* __restrict__ ticket,
                uint32_t* __restrict__ status) {
    __shared__ uint32_t s_whist[kSortWarps][kRadix];
    __shared__ union {
        uint64_t keys[kSortTile];
        uint32_t vals[kSortTile];
    } s_x;
    __shared__ uint32_t s_local_off[kRadix];
    __shared__ uint32_t s_gbase[kRadix];
    __shared__ uint32_t s_warp_tot[kSortWarps];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const int warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * kSortTile;
    const int n_valid = (int)min((int64_t)kSortTile, M - tile_base);
    const uint32_t digit_mask = (1u << bits) - 1u;

    // ---- load (warp-striped: memory order == (warp, item, lane) order) ----
    uint64_t key[kSortItems];
    int32_t val[kSortItems];
    const int warp_off = warp * (kSortItems * 32);
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int local = warp_off + it * 32 + (int)lane;
        if (local < n_valid) {
            key[it] = __ldg(keys_in + tile_base + local);
            val[it] = __ldg(vals_in + tile_base + local);
        } else {
            key[it] = ~0ull;
            val[it] = 0;
        }
    }

    // ---- stable rank inside the warp ----
    uint32_t rank[kSortItems];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int local = warp_off + it * 32 + (int)lane;
        const bool valid = local < n_valid;
        const uint32_t d = (uint32_t)(key[it] >> shift) & digit_mask;
        const uint32_t dm = valid ? d : 0x100u;
        const uint32_t peers = __match_any_sync(0xffffffffu, dm);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && (int)lane == leader) {
            pre = s_whist[warp][d];
            s_whist[warp][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[it] = pre + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: exclusive scan over warps, tile count, look-back ----
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_whist[w][tid];
            s_whist[w][tid] = run;
            run += c;
        }
        const uint32_t count = run;
        // block exclusive scan of `count` over the 256 digit-threads
        uint32_t incl = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        uint32_t excl = incl - count;
        for (int w = 0; w < warp; ++w) excl += s_warp_tot[w];
        s_local_off[tid] = excl;

        // decoupled look-back for digit `tid`
        uint32_t prev = 0;
        uint32_t* my = status + (size_t)tile * kRadix + tid;
        if (tile == 0) {
            st_relaxed_u32(my, kStatPrefix | count);
        } else {
            st_relaxed_u32(my, kStatAgg | count);
            int64_t j = (int64_t)tile - 1;
            while (true) {
                const uint32_t v = ld_relaxed_u32(status + (size_t)j * kRadix + tid);
                if ((v & ~kStatMask) == 0) continue;
                prev += v & kStatMask;
                if (v & kStatPrefix) break;
                --j;
            }
            st_relaxed_u32(my, kStatPrefix | (prev + count));
        }
        s_gbase[tid] = bins[tid] + prev - excl;
    }
    __syncthreads();

    // ---- keys: scatter to shared in tile-sorted order, then contiguous runs to global ----
    uint32_t pos[kSortItems];
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int local = warp_off + it * 32 + (int)lane;
        const uint32_t d = (uint32_t)(key[it] >> shift) & digit_mask;
        pos[it] = s_local_off[d] + s_whist[warp][d] + rank[it];
        if (local < n_valid) s_x.keys[pos[it]] = key[it];
    }
    __syncthreads();
    uint32_t dst[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int j = tid + k * kSortThreads;
        if (j < n_valid) {
            const uint64_t kk = s_x.keys[j];
            const uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
            dst[k] = s_gbase[d] + (uint32_t)j;
            keys_out[dst[k]] = kk;
        }
    }
    __syncthreads();
    // ---- values ride the same permutation ----
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int local = warp_off + it * 32 + (int)lane;
        if (local < n_valid) s_x.vals[pos[it]] = (uint32_t)val[it];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int j = tid + k * kSortThreads;
        if (j < n_valid) vals_out[dst[k]] = (int32_t)s_x.vals[j];
    }
}

}  // namespace bsplat

using namespace bsplat;

static int sort_passes(int begin_bit, int end_bit) {
    return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
}

extern "C" size_t bsplat_radix_sort_workspace_bytes(int64_t M, int32_t begin_bit, int32_t end_bit) {
    if (M < 0 || end_bit < begin_bit) return 0;
    return carve_ws(nullptr, M, sort_passes(begin_bit, end_bit)).bytes;
}

extern "C" int bsplat_radix_sort_pairs(int64_t M, uint64_t* keys, uint64_t* keys_alt, int32_t* vals,
                                       int32_t* vals_alt, int32_t begin_bit, int32_t end_bit,
                                       void* workspace, size_t workspace_bytes, int32_t* result_in_alt,
                                       void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (result_in_alt) *result_in_alt = 0;
    if (M < 0 || begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) return BSPLAT_E_ARG;
    if (M >= (int64_t)kStatMask) return BSPLAT_E_OVERFLOW;
    const int passes = sort_passes(begin_bit, end_bit);
    if (M == 0 || passes == 0) return BSPLAT_OK;
    if (!keys || !keys_alt || !vals || !vals_alt) return BSPLAT_E_ARG;
    SortWs w = carve_ws(workspace, M, passes);
    if (!workspace || workspace_bytes < w.bytes) return BSPLAT_E_WORKSPACE;
    const int64_t n_tiles = ceil_div(M, kSortTile);

    BSPLAT_CUDA_TRY(cudaMemsetAsync(workspace, 0, w.bytes, stream));
    int64_t hist_blocks = ceil_div(M, 256 * 8);
    const int hist_grid = (int)(hist_blocks < 148 * 8 ? hist_blocks : 148 * 8);
    radix_histogram_kernel<<<hist_grid, 256, 0, stream>>>(M, keys, begin_bit, end_bit, passes, w.hist);
    BSPLAT_LAUNCH_CHECK();
    radix_scan_kernel<<<passes, kRadix, 0, stream>>>(w.hist);
    BSPLAT_LAUNCH_CHECK();

    uint64_t* ksrc = keys; uint64_t* kdst = keys_alt;
    int32_t* vsrc = vals; int32_t* vdst = vals_alt;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = (end_bit - shift) < kRadixBits ? (end_bit - shift) : kRadixBits;
        onesweep_kernel<<<(unsigned)n_tiles, kSortThreads, 0, stream>>>(
            M, ksrc, kdst, vsrc, vdst, shift, bits, w.hist + (size_t)p * kRadix, w.tickets + p,
            w.status + (size_t)p * n_tiles * kRadix);
        BSPLAT_LAUNCH_CHECK();
        uint64_t* tk = ksrc; ksrc = kdst; kdst = tk;
        int32_t* tv = vsrc; vsrc = vdst; vdst = tv;
    }
    if (result_in_alt) *result_in_alt = (ksrc == keys_alt) ? 1 : 0;
    return BSPLAT_OK;
}
