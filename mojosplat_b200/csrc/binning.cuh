// Shared declarations of the binning stage (binning.cu, binning2.cu).
#pragma once
#include "common.cuh"

namespace bsplat {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // Gaussians per block

// workspace layout of the scan: [0] chunk ticket, [1..] one 64-bit status word per chunk
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

struct BinParams {
    int W, H, tiles_w, tiles_h, semantics, row_begin, row_end;
    float tile_size_f;
};

__device__ inline void load_mean_radii(const float* __restrict__ means2d, const void* __restrict__ radii,
                                       int radii_is_float, int64_t i, float& mx, float& my, float& rx,
                                       float& ry) {
    mx = __ldg(means2d + 2 * i);
    my = __ldg(means2d + 2 * i + 1);
    if (radii_is_float) {
        const float* r = static_cast<const float*>(radii);
        rx = __ldg(r + 2 * i);
        ry = __ldg(r + 2 * i + 1);
    } else {
        const int32_t* r = static_cast<const int32_t*>(radii);
        rx = (float)__ldg(r + 2 * i);  // int32 -> float promotion of `means2d - radii`
        ry = (float)__ldg(r + 2 * i + 1);
    }
}


int bin_count_scan_launch(int64_t N, const int32_t* perm, const float* means2d, const void* radii,
                          int radii_is_float, const float* depths, const BinParams& p, uint32_t* offsets,
                          bsplat_bin_info* info, void* workspace, bool finalize_key_range, cudaStream_t stream,
                          uint2* rects, const unsigned long long* n_dev);
int tile_ranges_u32_launch(int64_t M, const uint32_t* sorted_tile_ids, int n_tiles, int32_t* tile_ranges,
                           cudaStream_t stream);
int make_bin_params(int32_t width, int32_t height, int32_t tile_size, int32_t row_begin, int32_t row_end,
                    int32_t semantics, BinParams* p);

}  // namespace bsplat
