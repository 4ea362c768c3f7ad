// iq_metric.cu — K6: one optimisation pass of the I/Q-imbalance corrector on a 1024-frame block.
//
// Replaces iq_correct_run_optimization + _estimate_power + _calculate_imbalance_metric +
// _calculate_power_spectrum (reference src/iq_correct.c:154-235, 315-389): 27 windowed
// 1024-point spectra (1 power estimate, 1 baseline metric, 25 hill-climb candidates) and a 5 %
// smoothing step.  The work is tiny and strictly sequential between candidates, so ONE CTA keeps
// the block, the Hamming window, the twiddles and the spectrum in shared memory and runs all 27
// evaluations back to back; the FFT is the radix-4 network of fft_core.cuh (1024 = 4^5), whose
// base-4 digit-reversed output order is undone when the dB spectrum is written.
// The reference draws the +-1 step directions from rand() (:391) seeded with time() (:92); here
// the caller supplies them (deterministic).
#include <cuda_runtime.h>

#include "../../include/iqgpu.h"
#include "fft_core.cuh"
#include "kernels.hpp"

namespace iqgpu {

using namespace fftcore;

constexpr int IQ_NFFT = 1024;
constexpr int IQ_THREADS = 256;
constexpr int IQ_PASSES = 25;               // IQ_MAX_PASSES, include/constants.h:160

struct IqOptResult { float mag, phase, avg_power, power_range; int optimized; };

__device__ __forceinline__ unsigned digitrev4_1024(unsigned p)
{
    // reverse the five base-4 digits of p
    unsigned r = 0;
#pragma unroll
    for (int d = 0; d < 5; d++) { r = (r << 2) | (p & 3u); p >>= 2; }
    return r;
}

// spectrum_db[(k + 512) % 1024] = 20 log10(|X[k]| / 1024 + 1e-12), X = FFT(window * correct(x))
__device__ void iq_power_spectrum(const float2* __restrict__ x, const float* __restrict__ win, const float2* __restrict__ tw,
                                  float2* __restrict__ buf, float* __restrict__ spec, float gain_adj, float phase_adj)
{
    const int t = threadIdx.x;
    const float magp1 = __fadd_rn(1.0f, gain_adj);
    for (int i = t; i < IQ_NFFT; i += IQ_THREADS) {
        const float2 v = x[i];
        const float re = __fmul_rn(v.x, magp1);
        const float im = __fadd_rn(v.y, __fmul_rn(phase_adj, v.x));
        buf[i] = make_float2(__fmul_rn(re, win[i]), __fmul_rn(im, win[i]));
    }
    __syncthreads();
    for (unsigned L = IQ_NFFT; L >= 4; L >>= 2) {
        dif4(buf, L, (unsigned)t, tw, IQ_NFFT / L);       // 256 butterflies per stage, one per thread
        __syncthreads();
    }
    for (int p = t; p < IQ_NFFT; p += IQ_THREADS) {
        const unsigned k = digitrev4_1024((unsigned)p);
        const float2 X = buf[p];
        float mag = (float)sqrt((double)X.x * (double)X.x + (double)X.y * (double)X.y);   // cabsf
        mag = __fdiv_rn(mag, (float)IQ_NFFT);
        spec[(k + IQ_NFFT / 2) & (IQ_NFFT - 1)] = __fmul_rn(20.0f, log10f(__fadd_rn(mag, 1e-12f)));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(IQ_THREADS) iq_optimize_kernel(const float2* __restrict__ block, const float* __restrict__ dirs,
                                                                 float mag_in, float phase_in, IqOptResult* __restrict__ out)
{
    __shared__ float2 x[IQ_NFFT], buf[IQ_NFFT], tw[IQ_NFFT];
    __shared__ float win[IQ_NFFT], spec[IQ_NFFT];
    __shared__ float s_metric;
    __shared__ float s_avg, s_range;
    const int t = threadIdx.x;
    for (int i = t; i < IQ_NFFT; i += IQ_THREADS) {
        x[i] = block[i];
        float s, c;
        sincospif(-2.0f * (float)i / (float)IQ_NFFT, &s, &c);
        tw[i] = make_float2(c, s);
        // iq_correct.c:121-123 Hamming window, evaluated in float like the reference
        win[i] = __fsub_rn(0.54f, __fmul_rn(0.46f, cosf(__fdiv_rn(__fmul_rn(__fmul_rn(2.0f, 3.14159265358979323846f), (float)i), (float)(IQ_NFFT - 1)))));
    }
    __syncthreads();
    const int half = IQ_NFFT / 2;
    const int lo = (int)(0.05f * half), hi = (int)(0.95f * half);

    // ---- _estimate_power (:361-389)
    iq_power_spectrum(x, win, tw, buf, spec, 0.0f, 0.0f);
    if (t == 0) {
        float mx = -1000.0f;
        double sum = 0.0;
        int count = 0;
        for (int i = lo; i < hi; i++) {
            const float pn = spec[i], pp = spec[IQ_NFFT - 1 - i];
            if (pp > mx) mx = pp;
            if (pn > mx) mx = pn;
            sum += (double)__fadd_rn(pp, pn);
            count += 2;
        }
        s_avg = count ? (float)(sum / count) : 0.0f;
        s_range = count ? __fsub_rn(mx, s_avg) : 0.0f;
    }
    __syncthreads();
    if (s_range < 20.0f) {                                  // IQ_CORRECTION_POWER_THRESHOLD_DB (:168)
        if (t == 0) *out = IqOptResult{mag_in, phase_in, s_avg, s_range, 0};
        return;
    }
    // ---- hill climb (:177-201); the metric sum runs sequentially in float like the reference
    float cur_g = mag_in, cur_p = phase_in, best = 0.f;
    for (int pass = -1; pass < IQ_PASSES; pass++) {
        float cg = cur_g, cp = cur_p;
        if (pass >= 0) {
            cg = __fadd_rn(cur_g, __fmul_rn(0.0001f, dirs[2 * pass]));
            cp = __fadd_rn(cur_p, __fmul_rn(0.0001f, dirs[2 * pass + 1]));
        }
        iq_power_spectrum(x, win, tw, buf, spec, cg, cp);
        if (t == 0) {
            float total = 0.0f;
            for (int i = lo; i < hi; i++) {
                const float pn = spec[i], pp = spec[IQ_NFFT - 1 - i];
                if (pp > -80.0f || pn > -80.0f) {
                    const float d = __fsub_rn(pp, pn);
                    total = __fadd_rn(total, __fmul_rn(d, d));
                }
            }
            s_metric = total;
        }
        __syncthreads();
        const float m = s_metric;
        if (pass < 0) best = m;
        else if (m > best) { best = m; cur_g = cg; cur_p = cp; }   // keeps a candidate when the metric INCREASES (:196, quirk B8)
        __syncthreads();
    }
    if (t == 0) {
        // 5 % smoothing into the inactive slot (:206-216)
        const float sg = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, 0.05f), mag_in), __fmul_rn(0.05f, cur_g));
        const float sp = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, 0.05f), phase_in), __fmul_rn(0.05f, cur_p));
        *out = IqOptResult{sg, sp, s_avg, s_range, 1};
    }
}

cudaError_t iq_optimize_device(const float* host_block1024, const float* host_dirs50, float* mag, float* phase,
                               float* avg_power, float* power_range, int* optimized)
{
    float2* d_block = nullptr;
    float* d_dirs = nullptr;
    IqOptResult* d_out = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d_block, IQ_NFFT * sizeof(float2))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_dirs, 2 * IQ_PASSES * sizeof(float))) != cudaSuccess) { cudaFree(d_block); return e; }
    if ((e = cudaMalloc(&d_out, sizeof(IqOptResult))) != cudaSuccess) { cudaFree(d_block); cudaFree(d_dirs); return e; }
    IqOptResult r{};
    e = cudaMemcpy(d_block, host_block1024, IQ_NFFT * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_dirs, host_dirs50, 2 * IQ_PASSES * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        iq_optimize_kernel<<<1, IQ_THREADS>>>(d_block, d_dirs, *mag, *phase, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost);
    cudaFree(d_block); cudaFree(d_dirs); cudaFree(d_out);
    if (e != cudaSuccess) return e;
    *mag = r.mag; *phase = r.phase;
    if (avg_power) *avg_power = r.avg_power;
    if (power_range) *power_range = r.power_range;
    if (optimized) *optimized = r.optimized;
    return cudaSuccess;
}

}  // namespace iqgpu
