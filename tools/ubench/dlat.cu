// ubench: dependent-issue latency of the operations on the RMS-AGC critical path (one warp, one chain), sm_100a:
//   DFMA, DMUL, DADD, F2F.F64.F32, F2F.F32.F64, FFMA, FMUL, MUFU.EX2, LDS — cycles per dependent operation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/_build/dlat tools/ubench/dlat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, long long* cyc, int iters, double seed, float fseed)
{
    __shared__ float sh[64];
    sh[threadIdx.x & 63] = fseed;
    __syncthreads();
    double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
    float fa = fseed + threadIdx.x, fb = 1.0000001f, fc = 1e-9f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == 0) a = fma(a, b, c);
            if (OP == 1) a = a * b;
            if (OP == 2) a = a + c;
            if (OP == 3) { fa = (float)a; a = (double)fa; }          // F2F.F32.F64 + F2F.F64.F32
            if (OP == 4) fa = fmaf(fa, fb, fc);
            if (OP == 5) fa = fa * fb;
            if (OP == 6) fa = exp2f(fa) * 1e-3f;                      // MUFU.EX2 + FMUL
            if (OP == 7) fa = sh[(__float_as_uint(fa) >> 20) & 63];   // LDS (address depends on the previous load) + SHF + LOP
            if (OP == 8) { fa = (float)a; a = fma((double)fa, b, c); }   // F2F + F2F + DFMA
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + fa;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, 8 * sizeof(long long));
    const char* names[] = {"DFMA", "DMUL", "DADD", "F2F.F32.F64 + F2F.F64.F32", "FFMA", "FMUL", "MUFU.EX2 + FMUL", "LDS + SHF + LOP3", "F2F + F2F + DFMA"};
    const int iters = 2000;
    for (int warps = 1; warps <= 8; warps *= 8) {
        printf("-- %d warp(s) per SM sub-partition group (block of %d threads)\n", warps, 32 * warps);
        for (int op = 0; op < 9; op++) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: k<0><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 1: k<1><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 2: k<2><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 3: k<3><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 4: k<4><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 5: k<5><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 6: k<6><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 7: k<7><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                    case 8: k<8><<<1, 32 * warps>>>(out, cyc, iters, 1.0, 1.0f); break;
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            }
            printf("%-28s %7.2f cycles per dependent step   (%s)\n", names[op], (double)h / (iters * 16.0), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
