// ubench: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a in the operand forms the kernels use:
//   acc = x * h + acc with h either a per-thread register or a warp-uniform scalar (UR.F32 / constant-bank operand).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/_build/ffma2 tools/ubench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct Taps { float h[16]; };
// MODE 0: FFMA, register tap; 1: FFMA2, register tap pair; 2: FFMA, uniform tap; 3: FFMA2, uniform scalar tap
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, const __grid_constant__ Taps T)
{
    const float tv = threadIdx.x * 1e-9f;
    if (MODE == 0 || MODE == 2) {
        float a[16], x[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { a[i] = threadIdx.x + i; x[i] = 1.0f + tv * i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float h = (MODE == 2) ? T.h[j] : T.h[j] + tv;
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = fmaf(x[(i + j) & 15], h, a[i]);
            }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        u64 a[8], x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = pk(threadIdx.x + i, threadIdx.x - i); x[i] = pk(1.0f + tv * i, 1.0f - tv * i); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float h = (MODE == 3) ? T.h[j] : T.h[j] + tv;
                const u64 hh = pk(h, h);
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = fma2(x[(i + j) & 7], hh, a[i]);
            }
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
    Taps T;
    for (int i = 0; i < 16; i++) T.h[i] = 0.001f * (i + 1);
    const int iters = 10000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 512>>>(out, 100, T);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, 512>>>(out, iters, T);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = 148.0 * 2 * 512 * (double)iters * 64;   // thread-level FMA instructions
    printf("%-22s %8.3f ms  %7.2f Tinst/s (thread)  %7.2f TFLOP/s  err=%s\n", name, ms, inst / ms / 1e9,
           inst * flops_per_inst / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main()
{
    for (int rep = 0; rep < 2; rep++) {
        run<0>("FFMA  reg tap", 2);
        run<2>("FFMA  uniform tap", 2);
        run<1>("FFMA2 reg tap pair", 4);
        run<3>("FFMA2 uniform scalar", 4);
    }
    return 0;
}
