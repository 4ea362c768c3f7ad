// ubench.cu — the FP32 FMA roofline denominator, measured on the device the chain runs on ("FP32 FMA peak is
// not measured by the driver ... the builder must measure it").  Same instruction the kernels spend their FLOPs in:
// fma.rn.f32x2 with a warp-uniform scalar tap (SASS FFMA2 ..., UR.F32), 8 independent accumulators per thread, 2 CTAs of
// 512 threads per SM.  Not on the data path; bench.py calls it once per run.
#include <cuda_runtime.h>

#include "../../include/iqgpu.h"

namespace {
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct Taps { float h[8]; };

__global__ void __launch_bounds__(512) ffma2_peak_kernel(float* out, int iters, const __grid_constant__ Taps T)
{
    const float tv = threadIdx.x * 1e-9f;
    u64 a[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = pk(threadIdx.x + i, threadIdx.x - i); x[i] = pk(1.0f + tv * i, 1.0f - tv * i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const u64 hh = pk(T.h[j], T.h[j]);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma2(x[(i + j) & 7], hh, a[i]);
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}
}  // namespace

extern "C" int iqgpu_ubench_fp32_peak(int device, double* tflops, double* ms_out)
{
    if (!tflops) return IQGPU_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return IQGPU_ENODEVICE; }
    if (device < 0 || device >= ndev) return IQGPU_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return IQGPU_ECUDA;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return IQGPU_ECUDA;
    const int grid = prop.multiProcessorCount * 2, threads = 512, iters = 20000;
    float* out = nullptr;
    if (cudaMalloc(&out, (size_t)grid * threads * sizeof(float)) != cudaSuccess) return IQGPU_ENOMEM;
    Taps T;
    for (int i = 0; i < 8; i++) T.h[i] = 0.001f * (i + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    ffma2_peak_kernel<<<grid, threads>>>(out, 2000, T);              // warm-up: clocks up, code resident
    double best = 0.0, best_ms = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        ffma2_peak_kernel<<<grid, threads>>>(out, iters, T);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return IQGPU_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)grid * threads * (double)iters * 64.0 * 4.0;   // 64 FFMA2 per iteration, 4 FLOP each
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) { best = tf; best_ms = ms; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    if (ms_out) *ms_out = best_ms;
    return cudaGetLastError() == cudaSuccess ? IQGPU_OK : IQGPU_ECUDA;
}
