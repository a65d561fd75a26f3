// Shared device/host helpers for libbsplat (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bsplat.h"

#define BSPLAT_CUDA_TRY(expr)                                  \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return static_cast<int>(_e);    \
    } while (0)

// Checked build (make checked -> libbsplat_checked.so, -DBSPLAT_CHECKED): device-side bounds / invariant checks at
// the indexing hot spots (staging rings, shared-memory scatters, scatter positions, list bounds).  A failed check
// prints its location and traps.  compute-sanitizer is not available on every pool; benchmarks/selfcheck.py runs the
// whole kernel inventory under this build, with canaries around every buffer, and repeats frames for run-to-run
// determinism.
#ifdef BSPLAT_CHECKED
#include <cstdio>
#define BSPLAT_DASSERT(cond)                                                                    \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("BSPLAT_DASSERT failed: %s  (%s:%d)\n", #cond, __FILE__, __LINE__);           \
            __trap();                                                                           \
        }                                                                                       \
    } while (0)
#else
#define BSPLAT_DASSERT(cond) ((void)0)
#endif

// Programmatic dependent launch (sm_90+): a kernel launched with BSPLAT_LAUNCH_PDL may become resident while its
// predecessor in the stream is still draining; it must call pdl_wait() before it touches anything the predecessor
// wrote (first statement of every kernel launched this way) and pdl_trigger() right after it (so that ITS successor
// can be staged in turn; after the wait, so that a pre-staged kernel never overtakes two predecessors).  The frame is
// a chain of ~12 short dependent kernels: this takes the launch / ramp-up latency out of every boundary.
// BSPLAT_DEBUG=nopdl launches them the plain way (A/B).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#include <cstdlib>
#include <cstring>
#include <utility>
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* d = getenv("BSPLAT_DEBUG");
        return !(d && strstr(d, "nopdl"));
    }();
    return on;
}
// Per-call switch (thread-local): frames whose binning and rasterization go to two streams (the overlapped pipeline)
// launch the plain way -- a pre-staged kernel holds registers and shared memory that the OTHER stream's kernels would
// have used (measured: single frame 0.469 -> 0.453 ms with PDL, overlapped pipeline 2 772 -> 2 740 frames/s).
inline bool& pdl_call_switch() {
    static thread_local bool on = true;
    return on;
}
struct PdlScope {
    bool old;
    explicit PdlScope(bool on) : old(pdl_call_switch()) { pdl_call_switch() = on; }
    ~PdlScope() { pdl_call_switch() = old; }
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && pdl_call_switch()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#define BSPLAT_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                       \
    do {                                                                                                \
        cudaError_t e_ = launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__);       \
        if (e_ != cudaSuccess) return (int)e_;                                                          \
    } while (0)
#endif

#define BSPLAT_LAUNCH_CHECK()                                  \
    do {                                                       \
        cudaError_t _e = cudaGetLastError();                   \
        if (_e != cudaSuccess) return static_cast<int>(_e);    \
    } while (0)

namespace bsplat {

constexpr int kWarp = 32;

// Image buffers of the OTHER ranks (peer-mapped device pointers, same layout as the local image): the
// rasterizer stores every finished tile there as well, so the row-band exchange needs no separate collective.
constexpr int kMaxPeers = 7;
struct PeerImages {
    float* p[kMaxPeers];
    int n;
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Optional epilogue products of the projection kernel inside a fused frame (any pointer may be null):
// what the binning stage and the rasterizer would otherwise compute in passes of their own.
struct ProjExtra {
    uint2* rects;         // [N] full-frame tile rectangle: x0 | y0 << 16, w | h << 16 (rules of `semantics`)
    uint32_t* dkeys;      // [N] monotone depth keys
    uint32_t* hist;       // [4][256] digit histograms of the depth keys (caller-zeroed)
    void* rec;            // [N][3] float4 raster records (needs colors and opac)
    const float* colors;  // [N,3]
    const float* opac;    // [N] raw opacities
    int tile_size;
    int rec_row_begin, rec_row_end;  // (with rects) records only for Gaussians with a tile in these rows; 0, 0 = all
    // optional (fused-only launches): project list[0 .. *list_n) instead of all N Gaussians (device pointers)
    const int32_t* list = nullptr;
    const unsigned long long* list_n = nullptr;
};

// Monotone float -> uint32 map: ascending float order, -0.0 == +0.0, NaN last.
// Canonical depth order of the binning stage (reference: torch.argsort of float depths,
// binning.py:223).
__host__ __device__ inline uint32_t depth_key(float d) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(d);
#else
    union { float f; uint32_t u; } cv; cv.f = d; uint32_t b = cv.u;
#endif
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct TileRect {
    int x0, y0, x1, y1;  // exclusive ends, in tiles
};

// torch.clamp(v, lo, hi) for finite bounds: NaN propagates.
__device__ inline float clamp_torch(float v, float lo, float hi) {
    if (v != v) return v;
    v = v < lo ? lo : v;
    v = v > hi ? hi : v;
    return v;
}

__device__ inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Tile rectangle of one Gaussian.
//   BSPLAT_SEM_TORCH : binning.py:139-155,181-184 -- pixel-space clamp to [0, W-1] x [0, H-1] in
//     fp32, IEEE division by tile_size, truncation to int32, both ends inclusive; every Gaussian
//     (radii == 0 included) yields >= 1 tile.
//   BSPLAT_SEM_GSPLAT: floor / ceil in tile space, exclusive max, nothing for radii <= 0.
// row_begin/row_end clip the rectangle to a band of tile rows (multi-GPU row-band split).
__device__ inline TileRect tile_rect(float mx, float my, float rx, float ry, int W, int H,
                                     float tile_size_f, int tiles_w, int tiles_h, int semantics,
                                     int row_begin, int row_end) {
    TileRect r;
    if (semantics == BSPLAT_SEM_TORCH) {
        const float ax = clamp_torch(mx - rx, 0.0f, (float)(W - 1));
        const float bx = clamp_torch(mx + rx, 0.0f, (float)(W - 1));
        const float ay = clamp_torch(my - ry, 0.0f, (float)(H - 1));
        const float by = clamp_torch(my + ry, 0.0f, (float)(H - 1));
        // NaN -> int conversion: torch (x86 cvttss2si) gives INT_MIN, which the tile-space clamp
        // maps to 0; __float2int_rz(NaN) is 0 on the GPU -- same tile.
        r.x0 = clampi(__float2int_rz(__fdiv_rn(ax, tile_size_f)), 0, tiles_w - 1);
        r.x1 = clampi(__float2int_rz(__fdiv_rn(bx, tile_size_f)), 0, tiles_w - 1) + 1;
        r.y0 = clampi(__float2int_rz(__fdiv_rn(ay, tile_size_f)), 0, tiles_h - 1);
        r.y1 = clampi(__float2int_rz(__fdiv_rn(by, tile_size_f)), 0, tiles_h - 1) + 1;
    } else {
        if (!(rx > 0.0f) || !(ry > 0.0f)) {
            r.x0 = r.x1 = r.y0 = r.y1 = 0;
            return r;
        }
        const float fx0 = floorf(__fdiv_rn(mx - rx, tile_size_f));
        const float fx1 = ceilf(__fdiv_rn(mx + rx, tile_size_f));
        const float fy0 = floorf(__fdiv_rn(my - ry, tile_size_f));
        const float fy1 = ceilf(__fdiv_rn(my + ry, tile_size_f));
        r.x0 = (int)fminf(fmaxf(fx0, 0.0f), (float)tiles_w);
        r.x1 = (int)fminf(fmaxf(fx1, 0.0f), (float)tiles_w);
        r.y0 = (int)fminf(fmaxf(fy0, 0.0f), (float)tiles_h);
        r.y1 = (int)fminf(fmaxf(fy1, 0.0f), (float)tiles_h);
    }
    // band clip
    r.y0 = r.y0 < row_begin ? row_begin : r.y0;
    r.y1 = r.y1 > row_end ? row_end : r.y1;
    if (r.y1 < r.y0) r.y1 = r.y0;
    if (r.x1 < r.x0) r.x1 = r.x0;
    return r;
}

// Full-frame tile rectangle; inv_tile_size != 0 (power-of-two tile sizes only): x * (1 / tile_size) is then
// exactly x / tile_size, and the four IEEE divisions become multiplications.
__device__ inline TileRect tile_rect_inv(float mx, float my, float rx, float ry, int W, int H, float tile_size_f,
                                         float inv_tile_size, int tiles_w, int tiles_h, int semantics) {
    if (inv_tile_size == 0.0f)
        return tile_rect(mx, my, rx, ry, W, H, tile_size_f, tiles_w, tiles_h, semantics, 0, tiles_h);
    TileRect r;
    if (semantics == BSPLAT_SEM_TORCH) {
        const float ax = clamp_torch(mx - rx, 0.0f, (float)(W - 1));
        const float bx = clamp_torch(mx + rx, 0.0f, (float)(W - 1));
        const float ay = clamp_torch(my - ry, 0.0f, (float)(H - 1));
        const float by = clamp_torch(my + ry, 0.0f, (float)(H - 1));
        r.x0 = clampi(__float2int_rz(__fmul_rn(ax, inv_tile_size)), 0, tiles_w - 1);
        r.x1 = clampi(__float2int_rz(__fmul_rn(bx, inv_tile_size)), 0, tiles_w - 1) + 1;
        r.y0 = clampi(__float2int_rz(__fmul_rn(ay, inv_tile_size)), 0, tiles_h - 1);
        r.y1 = clampi(__float2int_rz(__fmul_rn(by, inv_tile_size)), 0, tiles_h - 1) + 1;
    } else {
        if (!(rx > 0.0f) || !(ry > 0.0f)) {
            r.x0 = r.x1 = r.y0 = r.y1 = 0;
            return r;
        }
        const float fx0 = floorf(__fmul_rn(mx - rx, inv_tile_size));
        const float fx1 = ceilf(__fmul_rn(mx + rx, inv_tile_size));
        const float fy0 = floorf(__fmul_rn(my - ry, inv_tile_size));
        const float fy1 = ceilf(__fmul_rn(my + ry, inv_tile_size));
        r.x0 = (int)fminf(fmaxf(fx0, 0.0f), (float)tiles_w);
        r.x1 = (int)fminf(fmaxf(fx1, 0.0f), (float)tiles_w);
        r.y0 = (int)fminf(fmaxf(fy0, 0.0f), (float)tiles_h);
        r.y1 = (int)fminf(fmaxf(fy1, 0.0f), (float)tiles_h);
    }
    if (r.y1 < r.y0) r.y1 = r.y0;
    if (r.x1 < r.x0) r.x1 = r.x0;
    return r;
}

// Packed tile rectangle (x0 | y0 << 16, w | h << 16) clipped to the tile rows [row_begin, row_end).
__device__ inline uint2 clip_rect_rows(const uint2 rc, const int row_begin, const int row_end) {
    int y0 = (int)(rc.x >> 16), h = (int)(rc.y >> 16);
    const int w = (int)(rc.y & 0xffffu);
    int y1 = y0 + h;
    y0 = y0 < row_begin ? row_begin : y0;
    y1 = y1 > row_end ? row_end : y1;
    h = y1 > y0 ? y1 - y0 : 0;
    return make_uint2((rc.x & 0xffffu) | ((uint32_t)y0 << 16), (uint32_t)(h > 0 ? w : 0) | ((uint32_t)h << 16));
}

__device__ inline uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ inline uint32_t lanemask_lt() {
    uint32_t m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// volatile / relaxed accesses used by the decoupled look-back chains
__device__ inline uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ inline void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ inline unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ inline void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace bsplat
