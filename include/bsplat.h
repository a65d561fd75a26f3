/*
 * bsplat.h -- C ABI of libbsplat.so, the B200 (sm_100a) backend of the MojoSplat forward path.
 *
 * Plain C: raw device pointers, sizes as runtime integers, a cudaStream_t passed as void*.
 * No torch / C++ types cross this boundary, nothing is allocated inside (the caller owns every
 * buffer and passes workspace), no global mutable state: calls are re-entrant and thread-safe,
 * one stream per call.  Every entry point returns 0 (BSPLAT_OK) or a negative bsplat error /
 * a positive cudaError_t; bsplat_error_string() explains either.
 *
 * Each entry point replaces one reference interface (paths relative to the reference checkout):
 *
 *   bsplat_project_fwd        mojosplat/projection.py:15-48 project_gaussians
 *                             (= MAX op `project_gaussians`, kernels/projection.mojo:260-328,
 *                              destination-passing call at projection.py:438-454; gsplat
 *                              fully_fused_projection call at projection.py:381-397)
 *   bsplat_bin_count_scan     mojosplat/binning.py:139-168   (tile rects, counts, total)
 *   bsplat_bin_emit           mojosplat/binning.py:172-209   (emission loop; gsplat isect_tiles,
 *                              binning.py:73-82)
 *   bsplat_radix_sort_pairs   mojosplat/binning.py:223-231   (argsort depth + stable argsort tile)
 *   bsplat_tile_ranges        mojosplat/binning.py:252-260   (searchsorted -> tile_ranges;
 *                              gsplat isect_offset_encode, binning.py:84-100)
 *   bsplat_bin2_prepare/finish mojosplat/binning.py:108-262 as a whole (two-level sort, default path)
 *   bsplat_rasterize_fwd      mojosplat/rasterization.py:13-57 rasterize_gaussians
 *                             (= MAX op `rasterize_to_pixels_3dgs_fwd`,
 *                              kernels/rasterization.mojo:169-240, call at rasterization.py:167-183)
 *   bsplat_rasterize_fwd_train / bsplat_rasterize_bwd   no reference counterpart (the reference is
 *                             forward-only: render.py:11 @torch.no_grad, README.md:145); derivative of
 *                             kernels/rasterization.mojo:138-162, checked against torch autograd
 *   bsplat_render_fwd         mojosplat/render.py:12-103 render_gaussians (the three stages chained
 *                              on one stream with a single 16-byte read-back)
 *   bsplat_render_enqueue     same without the read-back (device-side M, two-stream overlap, graph capture)
 *   bsplat_render_fwd_host    same, host buffers in / host image out (end-to-end path)
 *
 * Array layouts are the reference's: row-major AoS, fp32 / int32, single camera per launch.
 */
#ifndef BSPLAT_H_
#define BSPLAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSPLAT_VERSION 1

/* error codes (negative; positive values are cudaError_t) */
#define BSPLAT_OK 0
#define BSPLAT_E_ARG (-1)        /* bad argument (null pointer, negative size, unsupported tile size ...) */
#define BSPLAT_E_WORKSPACE (-2)  /* workspace too small; *needed_bytes tells how much is required */
#define BSPLAT_E_OVERFLOW (-3)   /* more than 2^30-1 intersections */
#define BSPLAT_E_NODEVICE (-4)   /* no CUDA device / wrong architecture */

/* binning / projection rule sets */
#define BSPLAT_SEM_TORCH 0   /* reference torch backend: projection.py:199-283, binning.py:139-162 */
#define BSPLAT_SEM_GSPLAT 1  /* gsplat / Mojo rules: projection.mojo:59-87,213-244; isect_tiles */

/* or-ed into the `semantics` argument of the stage-level entry points */
#define BSPLAT_PROJ_ALLOW_FMA 0x100  /* bsplat_project_fwd: let the compiler contract a*b+c into FMAs (A/B build of
                                        the kernel; faster, radii may then differ from the reference by 1 for a few
                                        Gaussians per million -- the default rounds like the eager torch ops) */
#define BSPLAT_PROJ_FAST_MATH 0x200  /* bsplat_project_fwd: MUFU-based exp / reciprocal / square root and free FMA
                                        contraction (third build of the kernel).  Values within 1e-4 + 1e-4 |ref|;
                                        a radius next to an integer boundary may come out one off, so tile lists
                                        can differ from the reference's -- an option for callers that want the
                                        HBM-bound kernel, never the default */
#define BSPLAT_BIN_PACKED 0x100      /* bsplat_bin2_prepare / bsplat_bin2_finish (pass it to both): compact the
                                        Gaussians that own no tile away before the depth sort (gsplat's packed
                                        layout; identical lists) */

/* rasterizer arithmetic */
#define BSPLAT_RASTER_FAST 0      /* folded exp2 form, packed FP32 pairs, sub-tile culling (default) */
#define BSPLAT_RASTER_FAITHFUL 1  /* operation order of kernels/rasterization.mojo:138-157 */
#define BSPLAT_RASTER_FAST_NOCULL 2  /* fast arithmetic without sub-tile culling (exactness A/B) */

/* bsplat_render_fwd `flags`: low byte = rasterizer mode, plus */
#define BSPLAT_FLAG_BIN_SINGLE_LEVEL 0x100  /* one sort of packed 64-bit keys instead of the two-level sort */
#define BSPLAT_FLAG_CAMERA_INDIRECT 0x200   /* bsplat_render_enqueue: the camera pose / intrinsics are read
                                               from `cam` (pinned host memory) when the stream gets there, so a
                                               captured CUDA graph can be replayed with a new pose; width and
                                               height are still read on the host at enqueue time */
#define BSPLAT_FLAG_PACKED 0x400            /* packed (culled-compacted) layout: Gaussians without a tile leave the
                                               frame right after projection (one stable compaction), so the depth
                                               sort, count + scan and the raster records only see the visible ones
                                               (projection.mojo:73-87,213-244; gsplat packed=True).  Same image and
                                               lists.  A no-op under BSPLAT_SEM_TORCH, where every Gaussian owns a tile */
#define BSPLAT_FLAG_PROJ_FMA 0x800          /* the BSPLAT_PROJ_ALLOW_FMA projection inside a fused frame */
#define BSPLAT_FLAG_NO_BAND_PRETEST 0x1000  /* row-band frames: project every Gaussian instead of only those a
                                             * conservative pre-test cannot rule out of the band (A/B and tests:
                                             * both ways give bit-identical lists and images) */
#define BSPLAT_FLAG_PROJ_FAST 0x2000        /* the BSPLAT_PROJ_FAST_MATH projection inside a fused frame */

/* Pinhole camera, world->camera. Mirrors mojosplat/utils.py:5-31 (Camera.view_matrix, Ks, H, W,
 * near, far) as a POD. viewmat is row-major 4x4. */
typedef struct bsplat_camera {
    float viewmat[16];
    float fx, fy, cx, cy;
    int32_t width, height;
    float near_plane, far_plane;
} bsplat_camera;

/* Device-side summary written by bsplat_bin_count_scan (32 bytes). */
typedef struct bsplat_bin_info {
    uint64_t n_isect;        /* M: number of (gaussian, tile) intersections */
    uint32_t min_depth_key;  /* min / max monotone depth key over Gaussians with >= 1 tile */
    uint32_t max_depth_key;
    uint32_t reserved[4];    /* [0]: scratch; [1]: 1 = M exceeded the capacity of a sync-free frame (bsplat_render_enqueue) */
} bsplat_bin_info;

/* Key layout chosen on the host from a bsplat_bin_info. */
typedef struct bsplat_key_layout {
    uint32_t depth_bias;  /* subtracted from every depth key (= min_depth_key) */
    int32_t depth_bits;   /* live depth bits after the bias */
    int32_t tile_bits;    /* ceil(log2(n_tiles)) */
} bsplat_key_layout;

int bsplat_version(void);
const char* bsplat_error_string(int code);
/* 0 when device `device` (or the current one if < 0) is an sm_100 part. */
int bsplat_check_device(int device);

/* ---- stage 1: projection ------------------------------------------------------------------- */
/* Projects N Gaussians through n_cams cameras; outputs are [n_cams, N, ...] contiguous.
 * opacities may be NULL for BSPLAT_SEM_TORCH (the torch backend ignores them). */
int bsplat_project_fwd(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                       const float* opacities, const bsplat_camera* cams_host, int32_t n_cams,
                       float eps2d, int32_t semantics, float* means2d, float* conics, float* depths,
                       int32_t* radii, void* stream);

/* ---- stage 2: binning ---------------------------------------------------------------------- */
size_t bsplat_bin_scan_workspace_bytes(int64_t N);
/* Tile rect per Gaussian -> exclusive prefix sum offsets[N+1] (uint32) and *info (device).
 * radii is int32 [N,2] (projection output) or float [N,2] when radii_is_float != 0
 * (the reference's tests pass float radii, tests/test_binning.py:23).
 * tile_row_begin/end restrict emission to a band of tile rows (row-band multi-GPU split);
 * pass 0 and tiles_h for the whole image. */
int bsplat_bin_count_scan(int64_t N, const float* means2d, const void* radii, int32_t radii_is_float,
                          const float* depths, int32_t width, int32_t height, int32_t tile_size,
                          int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics,
                          uint32_t* offsets, bsplat_bin_info* info, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Host helper: key layout (depth bias / live bits) from a read-back info. */
bsplat_key_layout bsplat_make_key_layout(const bsplat_bin_info* info_host, int32_t width,
                                         int32_t height, int32_t tile_size);
/* Writes keys[M] = (tile_id << depth_bits) | (depth_key - depth_bias) and gaussian ids[M] in
 * emission order (gaussian, tile row, tile column ascending). */
int bsplat_bin_emit(int64_t N, const float* means2d, const void* radii, int32_t radii_is_float,
                    const float* depths, int32_t width, int32_t height, int32_t tile_size,
                    int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics,
                    const uint32_t* offsets, bsplat_key_layout layout, uint64_t* keys, int32_t* ids,
                    void* stream);
size_t bsplat_radix_sort_workspace_bytes(int64_t M, int32_t begin_bit, int32_t end_bit);
/* Stable LSD onesweep radix sort of (uint64 key, int32 value) pairs on key bits
 * [begin_bit, end_bit). Ping-pongs between the two buffer pairs; *result_in_alt tells where the
 * sorted data ended up (0: keys/vals, 1: keys_alt/vals_alt). */
int bsplat_radix_sort_pairs(int64_t M, uint64_t* keys, uint64_t* keys_alt, int32_t* vals,
                            int32_t* vals_alt, int32_t begin_bit, int32_t end_bit, void* workspace,
                            size_t workspace_bytes, int32_t* result_in_alt, void* stream);
/* tile_ranges[n_tiles, 2] int32 = [first, past-last) index of each tile in the sorted keys
 * (searchsorted-left semantics for empty tiles, binning.py:252-260). */
int bsplat_tile_ranges(int64_t M, const uint64_t* sorted_keys, int32_t tile_shift, int32_t n_tiles,
                       int32_t* tile_ranges, void* stream);

/* Two-level binning (default path; same results, ~4x less sort traffic): depth-sort the N
 * Gaussians (4 onesweep passes over uint32 depth keys), count + scan and emit in depth order, then a
 * stable sort by tile id only (ceil(log2 n_tiles) bits in ceil(bits / 8) passes over the M pairs: 2 for up to
 * 65 536 tiles, 3-4 beyond) -- the structure of the reference's argsort(depth) + stable argsort(tile)
 * (binning.py:223-231).  means2d / radii / depths are read by prepare only (finish ignores them).
 * prepare: writes M to *info_out (device); the caller reads it back, sizes sorted_ids[M] and calls
 * finish with the same workspace (>= bsplat_bin2_workspace_bytes(N, M, n_tiles)). The last sort pass
 * also yields the per-tile counts, so tile_ranges (and the rasterizer's tile_order) cost one tiny kernel. */
size_t bsplat_bin2_workspace_bytes(int64_t N, int64_t M_capacity, int64_t n_tiles);
int bsplat_bin2_prepare(int64_t N, const float* means2d, const void* radii, int32_t radii_is_float,
                        const float* depths, int32_t width, int32_t height, int32_t tile_size,
                        int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics, void* workspace,
                        size_t workspace_bytes, bsplat_bin_info* info_out, void* stream);
int bsplat_bin2_finish(int64_t N, int64_t M, const float* means2d, const void* radii,
                       int32_t radii_is_float, int32_t width, int32_t height, int32_t tile_size,
                       int32_t tile_row_begin, int32_t tile_row_end, int32_t semantics, void* workspace,
                       size_t workspace_bytes, int32_t* sorted_ids, int32_t* tile_ranges,
                       int32_t* tile_order /* nullable: [tiles of the band], longest list first */,
                       void* stream);

/* ---- stage 3: rasterization ---------------------------------------------------------------- */
/* tile_order[n_tiles]: tile ids sorted by list length, longest first (scheduling hint for the
 * rasterizer: under the torch binning rules the culled Gaussians pile up in the corner tiles). */
int bsplat_tile_order(int32_t first_tile, int32_t n_tiles, const int32_t* tile_ranges,
                      int32_t* tile_order, void* stream);
/* image[height, width, channels]. opacities are used raw (no sigmoid), like the reference.
 * tile_order may be NULL (row-major tile order). Only tile rows [tile_row_begin, tile_row_end) are
 * rasterized and written (pass 0 and tiles_h for the whole frame; a row band for the multi-GPU split --
 * tile_order, if given, must then list exactly the band's tiles).
 * workspace (optional, bsplat_rasterize_workspace_bytes(N) = 48 B per Gaussian, 16-byte aligned): holds the
 * per-Gaussian raster records, written once per call; with it the fast kernel stages its batches with
 * cp.async one batch ahead (faster, identical results); without it it stages from the raw arrays.
 * mode: BSPLAT_RASTER_FAST / _FAITHFUL / _FAST_NOCULL (anything else: BSPLAT_E_ARG). */
size_t bsplat_rasterize_workspace_bytes(int64_t N);
int bsplat_rasterize_fwd(int64_t N, int32_t channels, const float* means2d, const float* conics,
                         const float* colors, const float* opacities, const float* background,
                         const int32_t* tile_ranges, const int32_t* tile_order,
                         const int32_t* sorted_ids, int64_t M, int32_t width, int32_t height,
                         int32_t tile_size, int32_t tile_row_begin, int32_t tile_row_end, int32_t mode,
                         float* image, void* workspace, size_t workspace_bytes, void* stream);

/* Faithful kernel + counters: stats[0] += evaluated (pixel, Gaussian) pairs, stats[1] += contributing
 * pairs (device uint64[2], caller-zeroed) -- the algorithmic work E_all / E_pass of SURVEY.md 8d. */
int bsplat_rasterize_stats(int64_t N, int32_t channels, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* background,
                           const int32_t* tile_ranges, const int32_t* sorted_ids, int64_t M,
                           int32_t width, int32_t height, int32_t tile_size, float* image,
                           uint64_t* stats, void* stream);

/* ---- colours: spherical harmonics (placeholder in the reference, render.py:82-87) -- */
/* colors[N,3] = max(sum_k Y_k(dir) coeffs[n,k,:] + 0.5, 0), dir = normalize(means3d[n] - campos), real SH of
 * `degree` (0..3), coeffs [N, K_stored, 3] with K_stored >= (degree+1)^2 (extra bands ignored).
 * campos_host: camera position in world space, 3 floats on the host. */
int bsplat_sh_eval(int64_t N, int32_t degree, int32_t K_stored, const float* coeffs, const float* means3d,
                   const float* campos_host, float* colors, void* stream);

/* ---- stage 3, training side (the reference is forward-only, render.py:11) --------- */
/* Forward with the faithful arithmetic (kernels/rasterization.mojo:138-162) that also stores what the
 * backward pass needs: final_T[height, width] (transmittance left at each pixel) and
 * last_idx[height, width] (index INTO sorted_ids of the last Gaussian composited there, -1 if none).
 * channels in 1..4; any tile size <= 32. */
int bsplat_rasterize_fwd_train(int64_t N, int32_t channels, const float* means2d, const float* conics,
                               const float* colors, const float* opacities, const float* background,
                               const int32_t* tile_ranges, const int32_t* sorted_ids, int64_t M,
                               int32_t width, int32_t height, int32_t tile_size, float* image,
                               float* final_T, int32_t* last_idx, void* stream);
/* The same through the fast forward kernel (16x16 tiles, RGB): image bit-identical to bsplat_rasterize_fwd's default
 * mode (within 1e-4 of the faithful arithmetic, like every fast frame); last_idx holds the last list entry the
 * backward pass has to look at (the entry in front of the one that saturated the pixel, or the end of the tile's
 * list). tile_order (optional, bsplat_tile_order): heavy tiles first. workspace (optional,
 * bsplat_rasterize_workspace_bytes(N)): per-Gaussian records + cp.async staging. */
int bsplat_rasterize_fwd_train_fast(int64_t N, const float* means2d, const float* conics, const float* colors,
                                    const float* opacities, const float* background, const int32_t* tile_ranges,
                                    const int32_t* tile_order, const int32_t* sorted_ids, int64_t M, int32_t width,
                                    int32_t height, float* image, float* final_T, int32_t* last_idx,
                                    void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the compositing: ACCUMULATES d loss / d (means2d[N,2], conics[N,3], colors[N,C],
 * opacities[N]) into the caller-zeroed gradient arrays, given grad_image[height, width, C] and the
 * final_T / last_idx of bsplat_rasterize_fwd_train[_fast] on the same inputs. Threshold tests are piecewise
 * constant; no gradient flows through the 0.999 alpha clamp. (d loss / d background =
 * sum_pixels final_T * grad_image is left to the caller.) */
int bsplat_rasterize_bwd(int64_t N, int32_t channels, const float* means2d, const float* conics,
                         const float* colors, const float* opacities, const float* background,
                         const int32_t* tile_ranges, const int32_t* sorted_ids, int64_t M,
                         int32_t width, int32_t height, int32_t tile_size, const float* final_T,
                         const int32_t* last_idx, const float* grad_image, float* grad_means2d,
                         float* grad_conics, float* grad_colors, float* grad_opacities, void* stream);

/* ---- fused forward ------------------------------------------------------------------------- */
/* Optional outputs of the fused path (any pointer may be NULL). */
typedef struct bsplat_render_aux {
    float* means2d;        /* [N,2] */
    float* conics;         /* [N,3] */
    float* depths;         /* [N]   */
    int32_t* radii;        /* [N,2] */
    int32_t* tile_ranges;  /* [tiles_h, tiles_w, 2] */
    int32_t* sorted_ids;   /* [sorted_ids_capacity]; filled when M <= capacity */
    int64_t sorted_ids_capacity;
    int64_t n_isect;       /* out: M */
    int32_t timing;        /* in: record per-stage CUDA-event times (adds a final sync) */
    int32_t n_launches;    /* out: kernels launched by this call */
    int32_t sort_passes;   /* out: onesweep passes (= ceil(live key bits / 8)) */
    int32_t key_bits;      /* out: live key bits (depth_bits + tile_bits) */
    float stage_ms[4];     /* out when timing != 0: projection, count+scan+emit, sort+ranges, raster */
} bsplat_render_aux;

size_t bsplat_render_workspace_bytes(int64_t N, int64_t M_capacity, int32_t width, int32_t height,
                                     int32_t tile_size);
/* render.py:63-101 on one stream. Returns BSPLAT_E_WORKSPACE and sets *needed_bytes when the
 * workspace cannot hold the M this frame produced (call again with a larger one).
 * M == 0 gives an all-zero image (render.py:73-76), not the background. */
int bsplat_render_fwd(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                      const float* opacities, const float* colors, int32_t channels,
                      const bsplat_camera* cam_host, const float* background, int32_t tile_size,
                      int32_t semantics, int32_t flags, float* image, void* workspace,
                      size_t workspace_bytes, size_t* needed_bytes, bsplat_render_aux* aux,
                      void* stream);
/* Split-phase frame for pipelined multi-view rendering. begin: projection + depth sort + count/scan
 * and an ASYNC copy of the bin info (M) into pinned host memory -- no sync. The caller waits for the
 * stream (event), reads M, then calls end (emit + tile sort + ranges + raster) with the same workspace,
 * which must hold bsplat_render_workspace_bytes(N, M, ...). Two streams with one workspace each overlap
 * begin(k+1) with end(k). Two-level binning only. */
int bsplat_render_begin(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                        const float* opacities, const bsplat_camera* cam_host, int32_t tile_size,
                        int32_t semantics, void* workspace, size_t workspace_bytes,
                        bsplat_bin_info* info_host_pinned, void* stream);
int bsplat_render_end(int64_t N, int64_t M, const float* colors, const float* opacities, int32_t channels,
                      const bsplat_camera* cam_host, const float* background, int32_t tile_size,
                      int32_t semantics, int32_t flags, float* image, void* workspace,
                      size_t workspace_bytes, size_t* needed_bytes, void* stream);
/* Sync-free frame (render.py:63-101 without the M read-back): nothing in this call waits for the GPU.
 * The workspace is sized for M_capacity pairs; the real M stays on the device. If a frame produces more
 * pairs than that, info.reserved[1] is set and the image is invalid: the caller inspects
 * *info_host_pinned (copied asynchronously on stream_bin; nullable) after synchronising and re-renders
 * with a larger capacity. stream_raster may be NULL / equal to stream_bin (single stream) or a second,
 * lower-priority stream: event_bin_done (a cudaEvent_t) then orders rasterization behind binning, and
 * binning of the next frame overlaps rasterization of this one. Safe to capture into a CUDA graph. */
int bsplat_render_enqueue(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                          const float* opacities, const float* colors, int32_t channels,
                          const bsplat_camera* cam, const float* background, int32_t tile_size,
                          int32_t semantics, int32_t flags, float* image, void* workspace,
                          size_t workspace_bytes, int64_t M_capacity, size_t* needed_bytes,
                          bsplat_bin_info* info_host_pinned, void* stream_bin, void* stream_raster,
                          void* event_bin_done);
/* The sync-free frame restricted to tile rows [tile_row_begin, tile_row_end) (row-band multi-GPU split):
 * all N Gaussians are projected, only the band is binned / rasterized / written into `image`. */
int bsplat_render_enqueue_band(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                               const float* opacities, const float* colors, int32_t channels,
                               const bsplat_camera* cam, const float* background, int32_t tile_size,
                               int32_t semantics, int32_t flags, int32_t tile_row_begin, int32_t tile_row_end,
                               float* image, void* workspace, size_t workspace_bytes, int64_t M_capacity,
                               size_t* needed_bytes, bsplat_bin_info* info_host_pinned, void* stream_bin,
                               void* stream_raster, void* event_bin_done);
/* ... with the band exchange fused into the rasterizer: every finished tile is also stored (128-bit stores)
 * into the image buffers of the n_peers (<= 7) other ranks -- peer-mapped device pointers of the same layout,
 * e.g. torch symmetric memory over NVLink / NVSwitch. Follow with a cross-device barrier on the stream; no
 * all-gather is needed. 16x16 tiles, RGB, BSPLAT_RASTER_FAST only. */
int bsplat_render_enqueue_band_p2p(int64_t N, const float* means3d, const float* log_scales, const float* quats,
                                   const float* opacities, const float* colors, int32_t channels,
                                   const bsplat_camera* cam, const float* background, int32_t tile_size,
                                   int32_t semantics, int32_t flags, int32_t tile_row_begin,
                                   int32_t tile_row_end, float* image, float* const* peer_images,
                                   int32_t n_peers, void* workspace, size_t workspace_bytes,
                                   int64_t M_capacity, size_t* needed_bytes, bsplat_bin_info* info_host_pinned,
                                   void* stream_bin, void* stream_raster, void* event_bin_done);
/* Same with HOST buffers (pinned or pageable): copies the Gaussians in, renders, copies the
 * image out and synchronises the stream. device_scratch must hold
 * bsplat_render_host_scratch_bytes() in addition to the render workspace (a shorter one is BSPLAT_E_ARG;
 * BSPLAT_E_WORKSPACE / *needed_bytes always refer to `workspace`). */
size_t bsplat_render_host_scratch_bytes(int64_t N, int32_t channels, int32_t width, int32_t height);
int bsplat_render_fwd_host(int64_t N, const float* means3d, const float* log_scales,
                           const float* quats, const float* opacities, const float* colors,
                           int32_t channels, const bsplat_camera* cam_host, const float* background,
                           int32_t tile_size, int32_t semantics, int32_t flags,
                           float* image_host, void* device_scratch, size_t scratch_bytes,
                           void* workspace, size_t workspace_bytes, size_t* needed_bytes,
                           bsplat_render_aux* aux, void* stream);

/* ---- measurement helper (not on the product path) ------------------------------------------ */
/* Issue-rate micro-benchmark: kind 0 = FP32 FFMA chains, kind 1 = MUFU.EX2 chains; blocks x 256
 * threads x 8 x iters operations. Timed by the caller with CUDA events; gives the measured
 * FP32 / SFU denominators of the rasterizer roofline. */
int bsplat_microbench(int32_t kind, int32_t blocks, int32_t iters, float* out, void* stream);

/* The backward pass in the fast kernel's layout (16x16 tiles, RGB; two pixels per lane, packed FP32, MUFU ex2 / rcp):
 * same contract as bsplat_rasterize_bwd; gradients agree with it within float rounding (both are checked against
 * fp64 autograd at 1e-3). workspace: bsplat_rasterize_workspace_bytes(N), 16-byte aligned (the records are rebuilt
 * here); tile_order optional (bsplat_tile_order). */
int bsplat_rasterize_bwd_fast(int64_t N, const float* means2d, const float* conics, const float* colors,
                              const float* opacities, const float* background, const int32_t* tile_ranges,
                              const int32_t* tile_order, const int32_t* sorted_ids, int64_t M, int32_t width,
                              int32_t height, const float* final_T, const int32_t* last_idx,
                              const float* grad_image, float* grad_means2d, float* grad_conics, float* grad_colors,
                              float* grad_opacities, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BSPLAT_H_ */
