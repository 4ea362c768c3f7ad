// ubench: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, per SM and whole chip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/_build/ffma2 tools/ubench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float h0)
{
    float h = h0 + threadIdx.x * 1e-9f;
    if (MODE == 0) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], h, 1.0f);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        u64 a[8];
        const u64 hh = pk(h, h), one = pk(1.0f, 1.0f);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = fma2(a[i], hh, one);
        }
        u64 s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
    }
}
template <int MODE>
static void run(const char* name, int flops_per_inst)
{
    float* out;
    cudaMalloc(&out, 148 * 4 * 512 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 512>>>(out, 100, 0.999f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, 512>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = 148.0 * 2 * 512 * (double)iters * 16;   // thread-level instructions
    printf("%-6s %8.3f ms  %7.2f Tinst/s (thread)  %7.2f TFLOP/s  err=%s\n", name, ms, inst / ms / 1e9,
           inst * flops_per_inst / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main()
{
    run<0>("FFMA", 2);
    run<1>("FFMA2", 4);
    run<0>("FFMA", 2);
    run<1>("FFMA2", 4);
    return 0;
}
