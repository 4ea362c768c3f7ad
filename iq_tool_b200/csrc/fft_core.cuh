// fft_core.cuh — butterflies of the hand-written power-of-two FFT used by the FFT block filter
// (K4) and by the I/Q optimiser's 1024-point spectrum (K6).  No cuFFT.
//
// Forward = decimation in frequency (natural order in, digit-reversed order out); inverse =
// the conjugate transpose of the same network run backwards (digit-reversed in, natural out).
// The filter multiplies by H in the digit-reversed domain, so no reordering pass ever runs:
// H is produced by pushing h||0 through the very same forward network.
//
// Every function here handles ONE butterfly `t` of ONE stage; the caller distributes t over
// threads (device) or loops over it (the host unit test in tests/ compiles this header with
// g++ and checks it against numpy.fft).
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define IQ_HD __host__ __device__ __forceinline__
#else
#define IQ_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif

namespace iqgpu {
namespace fftcore {

IQ_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
IQ_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
IQ_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
IQ_HD float2 cmulc(float2 a, float2 w) { return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }  // a * conj(w)
IQ_HD float2 mulmj(float2 a) { return make_float2(a.y, -a.x); }   // -j a
IQ_HD float2 mulpj(float2 a) { return make_float2(-a.y, a.x); }   // +j a

// tw[k] = exp(-j 2 pi k / NT); a stage of sub-length L uses stride tws = NT / L.

// radix-4 DIF butterfly t in [0, M/4) of the stage with sub-length L (L | M, L >= 4)
IQ_HD void dif4(float2* buf, unsigned L, unsigned t, const float2* tw, unsigned tws)
{
    const unsigned q = L >> 2, j = t & (q - 1);
    const unsigned i0 = ((t - j) << 2) + j, i1 = i0 + q, i2 = i1 + q, i3 = i2 + q;
    const float2 x0 = buf[i0], x1 = buf[i1], x2 = buf[i2], x3 = buf[i3];
    const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = mulmj(csub(x1, x3));
    float2 y0 = cadd(a, c), y1 = cadd(b, d), y2 = csub(a, c), y3 = csub(b, d);
    if (q > 1) {
        y1 = cmul(y1, tw[j * tws]);
        y2 = cmul(y2, tw[2 * j * tws]);
        y3 = cmul(y3, tw[3 * j * tws]);
    }
    buf[i0] = y0; buf[i1] = y1; buf[i2] = y2; buf[i3] = y3;
}

// conjugate transpose of dif4 (inverse DIT butterfly, un-normalised)
IQ_HD void dit4(float2* buf, unsigned L, unsigned t, const float2* tw, unsigned tws)
{
    const unsigned q = L >> 2, j = t & (q - 1);
    const unsigned i0 = ((t - j) << 2) + j, i1 = i0 + q, i2 = i1 + q, i3 = i2 + q;
    float2 u0 = buf[i0], u1 = buf[i1], u2 = buf[i2], u3 = buf[i3];
    if (q > 1) {
        u1 = cmulc(u1, tw[j * tws]);
        u2 = cmulc(u2, tw[2 * j * tws]);
        u3 = cmulc(u3, tw[3 * j * tws]);
    }
    const float2 a = cadd(u0, u2), b = csub(u0, u2), c = cadd(u1, u3), d = mulpj(csub(u1, u3));
    buf[i0] = cadd(a, c); buf[i1] = cadd(b, d); buf[i2] = csub(a, c); buf[i3] = csub(b, d);
}

// radix-2 DIF butterfly t in [0, M/2) of the stage with sub-length L
IQ_HD void dif2(float2* buf, unsigned L, unsigned t, const float2* tw, unsigned tws)
{
    const unsigned h = L >> 1, j = t & (h - 1);
    const unsigned i0 = ((t - j) << 1) + j, i1 = i0 + h;
    const float2 x0 = buf[i0], x1 = buf[i1];
    buf[i0] = cadd(x0, x1);
    buf[i1] = cmul(csub(x0, x1), tw[j * tws]);
}

IQ_HD void dit2(float2* buf, unsigned L, unsigned t, const float2* tw, unsigned tws)
{
    const unsigned h = L >> 1, j = t & (h - 1);
    const unsigned i0 = ((t - j) << 1) + j, i1 = i0 + h;
    const float2 u0 = buf[i0], u1 = cmulc(buf[i1], tw[j * tws]);
    buf[i0] = cadd(u0, u1);
    buf[i1] = csub(u0, u1);
}

// Stage schedule of an M-point transform (M = 2^p): if p is odd one radix-2 stage at L = M,
// then radix-4 stages L = M' , M'/4, ..., 4 with M' = M (p even) or M/2 (p odd).
// The helpers below run a whole transform serially (host reference / tiny device uses).
IQ_HD void forward_serial(float2* buf, unsigned M, const float2* tw, unsigned NT)
{
    unsigned L = M, p = 0;
    for (unsigned m = M; m > 1; m >>= 1) p++;
    if (p & 1) {
        for (unsigned t = 0; t < M / 2; t++) dif2(buf, L, t, tw, NT / L);
        L >>= 1;
    }
    for (; L >= 4; L >>= 2)
        for (unsigned t = 0; t < M / 4; t++) dif4(buf, L, t, tw, NT / L);
}

IQ_HD void inverse_serial(float2* buf, unsigned M, const float2* tw, unsigned NT)
{
    unsigned p = 0;
    for (unsigned m = M; m > 1; m >>= 1) p++;
    const unsigned Mq = (p & 1) ? M >> 1 : M;
    for (unsigned L = 4; L <= Mq; L <<= 2)
        for (unsigned t = 0; t < M / 4; t++) dit4(buf, L, t, tw, NT / L);
    if (p & 1)
        for (unsigned t = 0; t < M / 2; t++) dit2(buf, M, t, tw, NT / M);
}

}  // namespace fftcore
}  // namespace iqgpu
