// Stage 2b -- onesweep least-significant-digit radix sort (stable), uint32 or uint64 keys with
// int32 payloads.
//
// Replaces the two sorts of the reference's binning (mojosplat/binning.py:223-231: argsort by
// depth, then stable argsort by tile id) -- and gsplat's cub::DeviceRadixSort inside isect_tiles
// (binning.py:73-82).  Only live key bits [begin_bit, end_bit) are ever sorted.
//
// Structure (Adinets & Merrill, "Onesweep"):
//   radix_histogram_kernel  one read of the keys builds the digit histograms of ALL passes
//                           (the two-level binning path gets them from its producers instead)
//   radix_scan_kernel       exclusive scan of each 256-bin histogram
//   onesweep_kernel         per pass: each CTA takes a tile of pairs in ticket order, ranks them
//                           stably by digit (warp match-any multisplit), resolves its global digit
//                           offsets with a decoupled look-back chain (one chain per digit, one
//                           thread per digit, 8 predecessors fetched per round trip) and scatters
//                           through shared memory so that global stores are contiguous runs per digit.
// Per pass the kernel moves key+payload in and out once; nothing else touches HBM.
#include "radix_sort.cuh"

#ifdef BSPLAT_PHASES
// A/B instrumentation (not in the product build): per CTA, SM clock at the phase boundaries of a pass + the global
// timer at entry and exit.  Slot = pass (from the ticket address), row = tile.
__device__ unsigned long long g_phases[8][4096][10];
#define BSPLAT_PHASE(i)                                                                                  \
    do {                                                                                                 \
        if (threadIdx.x == 0 && ph_row < 4096) g_phases[ph_slot][ph_row][i] = (unsigned long long)clock64(); \
    } while (0)
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
extern "C" int bsplat_debug_phases(void* host_out, size_t bytes) {
    return (int)cudaMemcpyFromSymbol(host_out, g_phases, bytes < sizeof(g_phases) ? bytes : sizeof(g_phases));
}
#else
#define BSPLAT_PHASE(i)
#endif

namespace bsplat {

template <typename KeyT>
__global__ void __launch_bounds__(256)
radix_histogram_kernel(const int64_t M, const KeyT* __restrict__ keys, const int begin_bit,
                       const int end_bit, const int passes, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
        const KeyT key = keys[i];
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + p * kRadixBits;
            const int bits = min(kRadixBits, end_bit - shift);
            atomicAdd(&s_hist[p][(uint32_t)(key >> shift) & ((1u << bits) - 1u)], 1u);
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

// keys_out may be null on a final pass whose keys are not needed; vals_in may be null: the payload
// is then the element's own index (first pass of an argsort).
// m_dev (optional) overrides M with a device-side count (launches sized by a capacity).
//
// Phases: load -> stable rank (warp match-any multisplit) -> per-digit counts, published at once as
// the tile's aggregate -> scatter key/payload into shared memory in tile-sorted order (registers are
// free from here on) -> decoupled look-back, kLookback predecessors in flight per round trip ->
// contiguous per-digit runs to global.
// LOOKBACK: predecessors fetched per look-back round trip.  When every tile of a pass is resident at once (the
// N-scale depth passes: ~245 tiles), all tiles publish their aggregates at the same time and the last tile has to
// walk back through ALL of them -- tiles / LOOKBACK dependent round trips of ~1 us each bound the pass, so those
// launches use a 64-wide window (4 round trips instead of 15); the M-scale passes run in waves, find inclusive
// prefixes a few tiles back and keep the 16-wide window.
// THREADS: 256 (3 CTAs per SM), or 512 for the depth passes of frames whose tiles are all resident at once anyway
// (twice the tile, half the tiles: the look-back of such a pass is quadratic in the number of tiles, see above).
template <typename KeyT, int ITEMS, int BITS, int LOOKBACK = kLookback, int THREADS = kSortThreads>
__global__ void __launch_bounds__(THREADS, THREADS == kSortThreads ? BSPLAT_SORT_MINB : 1)
onesweep_kernel(const int64_t M_host, const uint64_t* __restrict__ m_dev, const KeyT* __restrict__ keys_in,
                KeyT* __restrict__ keys_out, const int32_t* __restrict__ vals_in,
                int32_t* __restrict__ vals_out, const int shift, const int bits_rt,
                const uint32_t* __restrict__ hist, const int hist_is_scanned, uint32_t* __restrict__ ticket,
                uint32_t* __restrict__ status, uint32_t* __restrict__ key_counts, const int key_row_stride,
                const int hist_early) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    static_assert(THREADS >= kRadix && THREADS % 32 == 0, "one thread per digit");
    // BITS > 0: digit width known at compile time (branch-free ballot loop with constant masks)
    const int bits = BITS > 0 ? BITS : bits_rt;
    __shared__ uint32_t s_whist[WARPS][kRadix];
    // key / payload exchange buffers: dynamic shared memory (TILE * (sizeof(KeyT) + 4) bytes; > 48 KB for 512 threads)
    extern __shared__ __align__(16) unsigned char s_dyn[];
    KeyT* s_keys = reinterpret_cast<KeyT*>(s_dyn);
    int32_t* s_vals = reinterpret_cast<int32_t*>(s_dyn + (size_t)TILE * sizeof(KeyT));
    __shared__ uint32_t s_local_off[kRadix];
    __shared__ uint32_t s_gbase[kRadix];
    __shared__ uint32_t s_bins[kRadix];
    __shared__ uint32_t s_warp_tot[kRadix / 32];
    __shared__ uint32_t s_tile;

    // Programmatic dependent launch: this CTA may be resident while the previous kernel of the stream is still
    // running.  What does not depend on that kernel happens before the wait: the ticket (its counter was zeroed at
    // the start of the frame), the shared-memory reset and -- when the caller says the histogram is older than the
    // previous kernel (hist_early) -- the histogram load.
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const int warp = tid >> 5;
#ifdef BSPLAT_PHASES
    const int ph_slot = (int)(((uintptr_t)ticket >> 2) & 7);
    uint32_t ph_row = blockIdx.x;  // until the ticket is known
    const unsigned long long ph_g0 = gtimer();
    const unsigned long long ph_c0 = clock64();
#endif
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * kRadix; i += THREADS) (&s_whist[0][0])[i] = 0;
    const bool dig = tid < kRadix;  // threads that own a digit (whole warps: 0 .. 7)
    const uint32_t h_early = (hist_early && dig) ? hist[tid] : 0u;
    pdl_wait();
    // device-side count (sync-free frames): M_host is then the capacity the launch and the buffers were
    // sized for; a count beyond it disables the pass (the emitter has flagged the overflow)
    int64_t M = M_host;
    if (m_dev) {
        const int64_t md = (int64_t)(*m_dev);
        M = md <= M_host ? md : 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
#ifdef BSPLAT_PHASES
    ph_row = tile;
    if (tid == 0 && ph_row < 4096) { g_phases[ph_slot][ph_row][8] = ph_g0; g_phases[ph_slot][ph_row][0] = ph_c0; }
    BSPLAT_PHASE(1);  // after the dependency wait
#endif
    if (tile_base >= M) return;  // launches sized by capacity: surplus CTAs retire (uniform per CTA)
    const int n_valid = (int)min((int64_t)TILE, M - tile_base);
    const uint32_t digit_mask = (1u << bits) - 1u;
    const int n_digits = 1 << bits;
    uint32_t* my_status = status + (size_t)tile * kRadix + (dig ? tid : 0);
    // global digit offsets: either already exclusive-scanned, or a raw histogram scanned here
    // (256-wide block scan; saves a kernel launch per sort)
    {
        const uint32_t h = dig ? (hist_early ? h_early : hist[tid]) : 0u;
        uint32_t incl = h;
        if (!hist_is_scanned) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            if (lane == 31 && dig) s_warp_tot[warp] = incl;
        }
        __syncthreads();
        if (dig) {
            uint32_t excl = h;
            if (!hist_is_scanned) {
                excl = incl - h;
                for (int w = 0; w < warp; ++w) excl += s_warp_tot[w];
            }
            s_bins[tid] = excl;
        }
        __syncthreads();  // s_warp_tot is reused below
    }

    {
        // ---- load (warp-striped: memory order == (warp, item, lane) order) ----
        KeyT key[ITEMS];
        int32_t val[ITEMS];
        const int warp_off = warp * (ITEMS * 32);
        const bool has_vals = vals_in != nullptr;
        const KeyT* kin = keys_in + tile_base + warp_off + lane;
        const int32_t* vin = has_vals ? vals_in + tile_base + warp_off + lane : nullptr;
        const int32_t idx0 = (int32_t)(tile_base + warp_off + lane);
        if (n_valid == TILE) {  // every tile but the last: no bounds checks
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                key[it] = kin[it * 32];
                val[it] = has_vals ? vin[it * 32] : idx0 + it * 32;
            }
        } else {
            // slots past the end hold all-ones keys: they rank behind every real element of the last digit (same
            // digit, later position), land at tile positions >= n_valid and are simply never written out
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int local = warp_off + it * 32 + (int)lane;
                if (local < n_valid) {
                    key[it] = kin[it * 32];
                    val[it] = has_vals ? vin[it * 32] : idx0 + it * 32;
                } else {
                    key[it] = (KeyT)~(KeyT)0;
                    val[it] = 0;
                }
            }
        }

        // ---- stable rank inside the warp ----
        // peers[it] = lanes of this warp holding the same digit.  Built from one ballot per digit bit
        // (fixed cost); match.any was measured ~3x slower here: its cost grows with the number of
        // distinct values in the warp, which is ~30 for high-entropy digits.
        uint32_t rank[ITEMS];
        uint32_t peers[ITEMS];
        const uint32_t lt = lanemask_lt();
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & digit_mask;
            uint32_t p = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < kRadixBits; ++b) {
                if (b < bits) {
                    // lanes sharing bit b with this lane: m ^ (bit ? 0 : ~0).  Written in PTX so that it stays at
                    // 4 instructions (LOP3 -> predicate, VOTE, SEL, LOP3); nvcc's own lowering took 6.
                    uint32_t m, x;
                    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\t"
                        "and.b32 t, %2, %3;\n\t"
                        "setp.ne.b32 q, t, 0;\n\t"
                        "vote.sync.ballot.b32 %0, q, 0xffffffff;\n\t"
                        "selp.b32 %1, 0, 0xffffffff, q;\n\t}"
                        : "=r"(m), "=r"(x) : "r"(d), "r"(1u << b));
                    p &= (m ^ x);
                }
            }
            peers[it] = p;
        }
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & digit_mask;
            const int leader = __ffs(peers[it]) - 1;
            uint32_t pre = 0;
            if ((int)lane == leader) {
                pre = s_whist[warp][d];
                s_whist[warp][d] = pre + __popc(peers[it]);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rank[it] = pre + __popc(peers[it] & lt);
            __syncwarp();
        }
        __syncthreads();
        BSPLAT_PHASE(2);  // loaded + ranked

        // ---- per-digit: exclusive scan over warps, tile count -> aggregate published immediately ----
        uint32_t run = 0;
        if (dig) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = s_whist[w][tid];
                s_whist[w][tid] = run;
                run += c;
            }
        }
        // the padding slots of the last tile were counted under the last digit: take them out again
        const uint32_t count = run - ((uint32_t)tid == digit_mask ? (uint32_t)(TILE - n_valid) : 0u);
        if (tid < n_digits) st_relaxed_u32(my_status, (tile == 0 ? kStatPrefix : kStatAgg) | count);
        uint32_t incl = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31 && dig) s_warp_tot[warp] = incl;
        __syncthreads();
        if (dig) {
            uint32_t excl = incl - count;
            for (int w = 0; w < warp; ++w) excl += s_warp_tot[w];
            s_local_off[tid] = excl;
            // s_gbase holds (count, excl) needs: keep count in s_gbase until the look-back rewrites it
            s_gbase[tid] = count;
        }
        __syncthreads();
        BSPLAT_PHASE(3);  // aggregates published, local scan

        // ---- scatter key / payload into shared memory in tile-sorted order ----
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & digit_mask;
            const uint32_t pos = s_local_off[d] + s_whist[warp][d] + rank[it];  // padding slots: pos >= n_valid
            BSPLAT_DASSERT(d < (uint32_t)kRadix && pos < (uint32_t)TILE);
            s_keys[pos] = key[it];
            s_vals[pos] = val[it];
        }
    }

    BSPLAT_PHASE(4);  // thread 0 has scattered its items
    // ---- decoupled look-back for digit `tid` (registers are free: wide window) ----
    if (tid < n_digits) {
        const uint32_t count = s_gbase[tid];
        uint32_t prev = 0;
        if (tile != 0) {
            int64_t j = (int64_t)tile - 1;
            bool found = false;
            while (!found) {
                // polite wait on the nearest unread predecessor: one word, exponential back-off, so that
                // waiting warps do not take issue slots from CTAs that still have to rank and publish
                const uint32_t* near = status + (size_t)j * kRadix + tid;
                uint32_t v0 = ld_relaxed_u32(near);
                unsigned ns = 32;
                while ((v0 & ~kStatMask) == 0) {
                    __nanosleep(ns);
                    if (ns < 512) ns <<= 1;
                    v0 = ld_relaxed_u32(near);
                }
                prev += v0 & kStatMask;
                if (v0 & kStatPrefix) break;
                --j;
                if (j < 0) break;
                // the ones behind it are older and almost surely published: fetch a window at once
                uint32_t v[LOOKBACK];
#pragma unroll
                for (int w = 0; w < LOOKBACK; ++w)
                    v[w] = (j - w >= 0) ? ld_relaxed_u32(status + (size_t)(j - w) * kRadix + tid) : kStatPrefix;
                int consumed = 0;
#pragma unroll
                for (int w = 0; w < LOOKBACK; ++w) {
                    if ((v[w] & ~kStatMask) == 0) break;  // not published yet: go back to the polite wait
                    prev += v[w] & kStatMask;
                    ++consumed;
                    if (v[w] & kStatPrefix) { found = true; break; }
                }
                j -= consumed;
                if (j < 0) break;
            }
            st_relaxed_u32(my_status, kStatPrefix | (prev + count));
        }
        s_gbase[tid] = s_bins[tid] + prev - s_local_off[tid];
    }
    __syncthreads();
    BSPLAT_PHASE(5);  // look-back of every digit done
    pdl_trigger();  // only the output is left: the next kernel of the stream may be staged now

    // ---- contiguous per-digit runs to global ----
    const bool write_keys = keys_out != nullptr, count_keys = key_counts != nullptr;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int j = tid + k * THREADS;
        if (j < n_valid) {
            const KeyT kk = s_keys[j];
            const uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
            const uint32_t dst = s_gbase[d] + (uint32_t)j;
            BSPLAT_DASSERT((int64_t)dst < M);
            if (write_keys) keys_out[dst] = kk;
            vals_out[dst] = s_vals[j];
            if (count_keys) {
                // Final pass of the tile sort: the CTA's elements are now fully sorted by key (globally
                // sorted by the low digit on input, stably re-sorted by the high digit here), so equal keys
                // form runs.  Run start j adds -j, run end j adds j+1: two atomics per run give
                // key_counts[key] += run length (uint32 wrap-around arithmetic).
                const bool first = (j == 0) || (s_keys[j - 1] != kk);
                const bool last = (j == n_valid - 1) || (s_keys[j + 1] != kk);
                if (first || last) {
                    // key_row_stride > 0: 2-D keys (row << 16 | column) counted under row * stride + column
                    const uint32_t ki = key_row_stride > 0
                                            ? (uint32_t)(kk >> 16) * (uint32_t)key_row_stride + ((uint32_t)kk & 0xffffu)
                                            : (uint32_t)kk;
                    if (first && last) atomicAdd(key_counts + ki, 1u);
                    else if (first) atomicAdd(key_counts + ki, (uint32_t)(0 - j));
                    else atomicAdd(key_counts + ki, (uint32_t)(j + 1));
                }
            }
        }
    }
#ifdef BSPLAT_PHASES
    __syncthreads();
    BSPLAT_PHASE(6);
    if (tid == 0 && ph_row < 4096) g_phases[ph_slot][ph_row][9] = gtimer();
#endif
}

// ------------------------------------------------------------------------------------------
int64_t sort_tiles_u32(int64_t M) { return ceil_div(M > 0 ? M : 1, kSortThreads * kSortItems32); }
int64_t sort_tiles_u64(int64_t M) { return ceil_div(M > 0 ? M : 1, kSortThreads * kSortItems64); }
// Depth passes of frames with at most kSortWideMax Gaussians run 512-thread CTAs over tiles of 8192 pairs, one CTA per
// SM: every tile of such a pass is resident at once, all tiles publish their aggregates at the same time and each
// one sums ALL its predecessors (nobody has an inclusive prefix yet), so the look-back costs tiles^2 / 2 status rows
// in total -- 30 % of the pass's instructions with 245 tiles of 4096, a quarter of that with 123 tiles of 8192.
// Only for frames that own the GPU (one stream, pdl_call_switch()): such a CTA takes the whole register file of an SM,
// and in the overlapped pipeline it would have to wait until every rasterizer CTA of another frame has left one
// (measured: 2 790 -> 2 650 frames/s), while the pipeline hides the look-back latency anyway.
// (BSPLAT_DEBUG=nowide: A/B switch.)
static bool sort_wide_enabled() {
    static const bool on = [] { const char* d = getenv("BSPLAT_DEBUG"); return !(d && strstr(d, "nowide")); }();
    return on;
}
bool sort_depth_wide(int64_t N) { return sort_wide_enabled() && pdl_call_switch() && N > 8192 && N <= kSortWideMax; }
int64_t sort_tiles_u32_depth(int64_t N) {
    return sort_depth_wide(N) ? ceil_div(N, (int64_t)kSortWideThreads * kSortItems32) : sort_tiles_u32(N);
}

size_t sort_status_words(int64_t n_tiles, int passes) { return (size_t)passes * n_tiles * kRadix; }

#define BSPLAT_ONESWEEP32(B)                                                                              \
    BSPLAT_LAUNCH_PDL((onesweep_kernel<uint32_t, kSortItems32, B>), (unsigned)n_tiles, kSortThreads, kSmem32, stream, \
        M, m_dev, keys_in, keys_out, vals_in, vals_out, shift, bits, hist, hist_is_scanned, ticket, status,    \
        key_counts, key_row_stride, hist_early)

int onesweep_pass_u32(int64_t M, const uint64_t* m_dev, const uint32_t* keys_in, uint32_t* keys_out,
                      const int32_t* vals_in, int32_t* vals_out, int shift, int bits, const uint32_t* hist,
                      int hist_is_scanned, uint32_t* ticket, uint32_t* status, uint32_t* key_counts,
                      cudaStream_t stream, int key_row_stride, int hist_early, int depth_pass) {
    constexpr size_t kSmem32 = (size_t)kSortThreads * kSortItems32 * (sizeof(uint32_t) + sizeof(int32_t));
    if (depth_pass && bits == 8 && sort_depth_wide(M)) {
        constexpr size_t kSmemWide = (size_t)kSortWideThreads * kSortItems32 * (sizeof(uint32_t) + sizeof(int32_t));
        static const cudaError_t attr = cudaFuncSetAttribute(
            onesweep_kernel<uint32_t, kSortItems32, 8, 64, kSortWideThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)kSmemWide);
        if (attr != cudaSuccess) return (int)attr;
        BSPLAT_LAUNCH_PDL((onesweep_kernel<uint32_t, kSortItems32, 8, 64, kSortWideThreads>),
                          (unsigned)sort_tiles_u32_depth(M), kSortWideThreads, kSmemWide, stream, M, m_dev, keys_in, keys_out,
                          vals_in, vals_out, shift, bits, hist, hist_is_scanned, ticket, status, key_counts, key_row_stride,
                          hist_early);
        BSPLAT_LAUNCH_CHECK();
        return BSPLAT_OK;
    }
    const int64_t n_tiles = sort_tiles_u32(M);
    if (bits == 8 && n_tiles <= 3 * 148) {  // every tile resident at once (3 CTAs per SM): wide look-back window
        BSPLAT_LAUNCH_PDL((onesweep_kernel<uint32_t, kSortItems32, 8, 64>), (unsigned)n_tiles, kSortThreads, kSmem32, stream, M, m_dev, keys_in, keys_out, vals_in, vals_out, shift, bits, hist, hist_is_scanned, ticket, status,
            key_counts, key_row_stride, hist_early);
        BSPLAT_LAUNCH_CHECK();
        return BSPLAT_OK;
    }
    switch (bits) {
        case 8: BSPLAT_ONESWEEP32(8); break;
        case 7: BSPLAT_ONESWEEP32(7); break;
        case 6: BSPLAT_ONESWEEP32(6); break;
        default: BSPLAT_ONESWEEP32(0); break;
    }
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

int radix_scan_launch(uint32_t* hist, int passes, cudaStream_t stream) {
    radix_scan_kernel<<<passes, kRadix, 0, stream>>>(hist);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}

}  // namespace bsplat

using namespace bsplat;

namespace {
struct SortWs {
    uint32_t* hist;     // [kMaxPasses][256]
    uint32_t* tickets;  // [256]
    uint32_t* status;   // [P][n_tiles][256]
    size_t bytes;
};

SortWs carve_ws(void* ws, int64_t M, int passes) {
    const int64_t n_tiles = sort_tiles_u64(M);
    SortWs w;
    uint32_t* p = static_cast<uint32_t*>(ws);
    w.hist = p;
    w.tickets = p + (size_t)kMaxPasses * kRadix;
    w.status = w.tickets + kRadix;
    w.bytes = ((size_t)kMaxPasses * kRadix + kRadix + sort_status_words(n_tiles, passes)) * sizeof(uint32_t);
    return w;
}

int sort_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits; }
}  // namespace

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
    const int64_t n_tiles = sort_tiles_u64(M);

    BSPLAT_CUDA_TRY(cudaMemsetAsync(workspace, 0, w.bytes, stream));
    int64_t hist_blocks = ceil_div(M, 256 * 4);
    const int hist_grid = (int)(hist_blocks < 148 * 8 ? hist_blocks : 148 * 8);
    radix_histogram_kernel<uint64_t><<<hist_grid, 256, 0, stream>>>(M, keys, begin_bit, end_bit, passes, w.hist);
    BSPLAT_LAUNCH_CHECK();
    int rc = radix_scan_launch(w.hist, passes, stream);
    if (rc != BSPLAT_OK) return rc;

    uint64_t* ksrc = keys; uint64_t* kdst = keys_alt;
    int32_t* vsrc = vals; int32_t* vdst = vals_alt;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = (end_bit - shift) < kRadixBits ? (end_bit - shift) : kRadixBits;
        BSPLAT_LAUNCH_PDL((onesweep_kernel<uint64_t, kSortItems64, 0>), (unsigned)n_tiles, kSortThreads,
                          (size_t)kSortThreads * kSortItems64 * (sizeof(uint64_t) + sizeof(int32_t)), stream, M, nullptr, ksrc, kdst, vsrc, vdst, shift, bits, w.hist + (size_t)p * kRadix, 1, w.tickets + p,
            w.status + (size_t)p * n_tiles * kRadix, nullptr, 0, 0);
        BSPLAT_LAUNCH_CHECK();
        uint64_t* tk = ksrc; ksrc = kdst; kdst = tk;
        int32_t* tv = vsrc; vsrc = vdst; vdst = tv;
    }
    if (result_in_alt) *result_in_alt = (ksrc == keys_alt) ? 1 : 0;
    return BSPLAT_OK;
}
