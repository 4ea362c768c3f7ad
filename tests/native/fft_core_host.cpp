// Host build of iq_tool_b200/csrc/fft_core.cuh for the CPU unit test (tests/test_fft_core.py):
// the same butterfly code the CUDA kernels run, driven serially.
#include <cstring>
#include <vector>
#include <cmath>
#include "../../iq_tool_b200/csrc/fft_core.cuh"
using namespace iqgpu::fftcore;
extern "C" {
void fftcore_twiddles(float* tw, unsigned NT)
{
    for (unsigned k = 0; k < NT; k++) {
        const double a = -2.0 * M_PI * (double)k / (double)NT;
        tw[2 * k] = (float)cos(a); tw[2 * k + 1] = (float)sin(a);
    }
}
void fftcore_forward(float* buf, unsigned M, const float* tw, unsigned NT) { forward_serial((float2*)buf, M, (const float2*)tw, NT); }
void fftcore_inverse(float* buf, unsigned M, const float* tw, unsigned NT) { inverse_serial((float2*)buf, M, (const float2*)tw, NT); }
// large transform the way the device does it: radix-2 DIF stages down to sub-blocks of Msub,
// then the sub-block network; inverse mirrors it
void fftcore_forward_split(float* buf, unsigned N, unsigned Msub, const float* tw)
{
    float2* b = (float2*)buf; const float2* w = (const float2*)tw;
    for (unsigned L = N; L > Msub; L >>= 1)
        for (unsigned t = 0; t < N / 2; t++) dif2(b, L, t, w, N / L);
    for (unsigned s = 0; s < N / Msub; s++) forward_serial(b + (size_t)s * Msub, Msub, w, N);
}
void fftcore_inverse_split(float* buf, unsigned N, unsigned Msub, const float* tw)
{
    float2* b = (float2*)buf; const float2* w = (const float2*)tw;
    for (unsigned s = 0; s < N / Msub; s++) inverse_serial(b + (size_t)s * Msub, Msub, w, N);
    for (unsigned L = 2 * Msub; L <= N; L <<= 1)
        for (unsigned t = 0; t < N / 2; t++) dit2(b, L, t, w, N / L);
}
}
