// Shared declarations of the onesweep radix sort (radix_sort.cu) used by the binning stage.
#pragma once
#include "common.cuh"

namespace bsplat {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems64 = 10;  // 2560 (uint64 key, int32) pairs per CTA
#ifndef BSPLAT_SORT_ITEMS32
#define BSPLAT_SORT_ITEMS32 16
#endif
#ifndef BSPLAT_SORT_MINB
#define BSPLAT_SORT_MINB 3
#endif
constexpr int kSortItems32 = BSPLAT_SORT_ITEMS32;  // 4096 (uint32 key, int32) pairs per CTA (80 registers, 3 CTAs per SM)
constexpr int kSortWideThreads = 512;  // depth passes of small frames: 8192 pairs per CTA, one CTA per SM
constexpr int64_t kSortWideMax = 148LL * kSortWideThreads * kSortItems32;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;
constexpr int kLookback = 16;  // predecessors fetched per look-back round trip

constexpr uint32_t kStatAgg = 1u << 30;
constexpr uint32_t kStatPrefix = 2u << 30;
constexpr uint32_t kStatMask = (1u << 30) - 1;

int64_t sort_tiles_u32(int64_t M);
int64_t sort_tiles_u64(int64_t M);
int64_t sort_tiles_u32_depth(int64_t N);  // tiles of a depth pass (onesweep_pass_u32 with depth_pass != 0)
size_t sort_status_words(int64_t n_tiles, int passes);

// One pass over (uint32 key, int32 payload) pairs on key bits [shift, shift+bits).
//   hist    256 digit counts of this pass: exclusive-scanned (hist_is_scanned != 0) or raw (the kernel
//           then scans them itself -- saves the separate scan launch)
//   ticket  one zeroed uint32; status: zeroed [sort_tiles_u32(M)][256] uint32
//   m_dev   optional device-side element count (grid still sized by M)
//   vals_in == nullptr: payload = element index; keys_out == nullptr: keys are not written
//   key_counts != nullptr (only valid on the LAST pass of a sort whose keys are fully covered by the
//   passes): key_counts[key] += number of elements with that key (zeroed by the caller); with key_row_stride > 0
//   the keys are 2-D (row << 16 | column) and are counted under row * key_row_stride + column
//   hist_early: the histogram was complete before the PREVIOUS kernel of the stream started (it may then be read
//   while that kernel is still running: programmatic dependent launch, common.cuh)
int onesweep_pass_u32(int64_t M, const uint64_t* m_dev, const uint32_t* keys_in, uint32_t* keys_out,
                      const int32_t* vals_in, int32_t* vals_out, int shift, int bits, const uint32_t* hist,
                      int hist_is_scanned, uint32_t* ticket, uint32_t* status, uint32_t* key_counts,
                      cudaStream_t stream, int key_row_stride = 0, int hist_early = 0, int depth_pass = 0);
// In-place exclusive scan of `passes` consecutive 256-bin histograms.
int radix_scan_launch(uint32_t* hist, int passes, cudaStream_t stream);

}  // namespace bsplat
