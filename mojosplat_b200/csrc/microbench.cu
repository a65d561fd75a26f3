// Roofline denominators for the rasterizer: measured FP32 FFMA and MUFU.EX2 issue peaks of the
// device (SURVEY.md 8d asks for measured, not nominal, FP32/SFU peaks).  Not on the product path.
#include "common.cuh"

namespace bsplat {

template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(const int iters, float* __restrict__ out, const float seed) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + (float)(threadIdx.x + k) * 1e-3f;
    const float m = 0.999f, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0) {
                a[k] = fmaf(a[k], m, c);
            } else if (KIND == 2) {
                // packed FP32 pair FMA (sm_100 FFMA2): 4 pair-chains over the 8 registers
                if (k < 4) {
                    unsigned long long x, mm, cc;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[2 * k]), "f"(a[2 * k + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(mm) : "f"(m), "f"(m));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(mm), "l"(cc));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * k]), "=f"(a[2 * k + 1]) : "l"(x));
                }
            } else {
                float y;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[k]));
                a[k] = y * 0.5f;  // one FMUL per EX2 keeps the chain bounded; MUFU stays the limiter
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; defeats DCE
}

}  // namespace bsplat

// kind 0: FFMA chain, kind 1: EX2 chain, kind 2: FFMA2 (packed pair) chain -- 4 instructions = 8 FMAs per 8 ops. Launches `blocks` x 256 threads, each doing 8*iters ops.
extern "C" int bsplat_microbench(int32_t kind, int32_t blocks, int32_t iters, float* out, void* stream) {
    if (blocks <= 0 || iters <= 0 || !out) return BSPLAT_E_ARG;
    if (kind == 0)
        bsplat::microbench_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out, 0.5f);
    else if (kind == 2)
        bsplat::microbench_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out, 0.5f);
    else
        bsplat::microbench_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out, 0.5f);
    BSPLAT_LAUNCH_CHECK();
    return BSPLAT_OK;
}
