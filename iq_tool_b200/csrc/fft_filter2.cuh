// fft_filter2.cuh — K4 v2: radix-16 register-butterfly FFT block filter (no cuFFT).
//
// Same job as fft_smem_kernel (fft_filter.cu): one CTA evaluates one block of liquid's fftfilt
// (reference src/filter.c:464-526 -> fftfilt_*_execute) in overlap-save form,
//     y[b n .. (b+1) n) = last n samples of IFFT( FFT(x[(b-1) n .. (b+1) n)) .* H ) / (2n),
// with the 2n-point transform resident in shared memory.  What changed, and why: the radix-4
// kernel makes log4(2n) = 7 shared-memory round trips per direction for 2n = 16384, re-reads its
// twiddles from global memory in every stage and its late stages run into 8-way bank conflicts;
// it reached ~5 % of the FP32 peak (cfg3: 16.4 ms per 1.2 G input frames).  Here
//   * every pass is a radix-16 (then radix-4 / radix-2 for the remaining bits) decimation-in-
//     frequency butterfly held in registers: 3 passes instead of 6 for 2n = 16384;
//   * the first forward pass reads the window straight from global memory and the last inverse
//     pass writes the kept half straight to the output (no staging copies);
//   * the forward network's last radix-4 pass, the multiplication by H and the inverse network's
//     first pass act on the same four adjacent points, so they are fused in registers;
//   * pass twiddles W^j, W^2j, W^4j, W^8j come from the table and the other eleven are products;
//   * the shared buffer is padded 4 points per 64 so that the stride-4 pass is conflict free.
// Forward = DIF passes (natural order in, digit-reversed out); inverse = the conjugate transpose of
// the same passes in reverse order.  H is produced by pushing h||0 through the forward network
// (fft2_forward_kernel), so its order always matches and no reordering pass exists.
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"
#include "device_common.cuh"

namespace iqgpu {
namespace fft2 {

// Complex arithmetic on packed {re, im} pairs (add/sub/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2): a complex add is one
// instruction instead of two, a complex multiply two instead of four — the swap of the halves and the sign of one half
// that the cross terms need are operand modifiers of the packed instructions (ptxas folds the mov.b64 re-packs into them).
// The radix-16 network is two thirds additions; the scalar form spent 66 % of its issue slots on FP instructions at
// 35 % FMA-pipe utilisation (profiles/r01f_fftfilt2_full_cfg3.md).  Rounding differs from the scalar form in the last
// bit of some products (fma contraction); H is produced by the same network, the parity bars are those of the oracle.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return unpk2(add2(pk2(a), pk2(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return unpk2(add2(pk2(a), pk2(-b.x, -b.y))); }
// a * w = {a.x w.x - a.y w.y, a.y w.x + a.x w.y} = {a.x, a.y} * w.x + {a.y, a.x} * {-w.y, w.y}
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return unpk2(fma2(pk2(a), pk2(w.x, w.x), mul2(pk2(a.y, a.x), pk2(-w.y, w.y))));
}
// a * conj(w) = {a.x w.x + a.y w.y, a.y w.x - a.x w.y}
__device__ __forceinline__ float2 cmulc(float2 a, float2 w)
{
    return unpk2(fma2(pk2(a), pk2(w.x, w.x), mul2(pk2(a.y, a.x), pk2(w.y, -w.y))));
}
using fftcore::mulmj;
using fftcore::mulpj;

constexpr int THREADS = 512;
constexpr unsigned MAX_POINTS = 16384;
__host__ __device__ constexpr unsigned pad(unsigned i) { return i + 4u * (i >> 6); }
constexpr unsigned SMEM_F2 = MAX_POINTS + 4 * (MAX_POINTS / 64) + 8;

// pass plan of an N = 2^p point transform: (p - 2) bits are consumed by radix-16, then radix-4,
// then radix-2 passes; the last two bits belong to the fused middle radix-4 pass.
struct Plan {
    unsigned N, log2n;
    unsigned n16, n4, n2;      // number of radix-16 / radix-4 / radix-2 passes
};
__host__ __device__ inline Plan make_plan(unsigned N)
{
    Plan p{N, 0, 0, 0, 0};
    for (unsigned m = N; m > 1; m >>= 1) p.log2n++;
    unsigned bits = p.log2n - 2;
    p.n16 = bits / 4; bits -= 4 * p.n16;
    p.n4 = bits / 2; bits -= 2 * p.n4;
    p.n2 = bits;
    return p;
}

template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d)
{
    // forward: X[k] = sum x[n] exp(-2 pi i n k / 4); INV: conjugate kernel
    const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d);
    const float2 s3 = INV ? mulpj(csub(b, d)) : mulmj(csub(b, d));
    a = cadd(s0, s2); b = cadd(s1, s3); c = csub(s0, s2); d = csub(s1, s3);
}

// 16-point DFT in registers, inputs u[q], outputs X[p] in place (natural order both sides).
//   q = b + 4a, p = k1 + 4 k2:  X[k1+4k2] = sum_b w4^(b k2) [ w16^(b k1) sum_a u[b+4a] w4^(a k1) ]
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&u)[16])
{
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; b++) dft4<INV>(u[b], u[b + 4], u[b + 8], u[b + 12]);   // over a: result index k1 at u[b + 4 k1]
    // internal twiddles w16^(b k1), b,k1 in 1..3 (conjugated for the inverse)
    const float2 w1 = make_float2(C1, INV ? S1 : -S1), w2 = make_float2(R2, INV ? R2 : -R2), w3 = make_float2(S1, INV ? C1 : -C1);
    const float2 w6 = make_float2(-R2, INV ? R2 : -R2), w9 = make_float2(-C1, INV ? -S1 : S1);
    u[1 + 4] = cmul(u[1 + 4], w1);  u[1 + 8] = cmul(u[1 + 8], w2);   u[1 + 12] = cmul(u[1 + 12], w3);
    u[2 + 4] = cmul(u[2 + 4], w2);  u[2 + 8] = INV ? mulpj(u[2 + 8]) : mulmj(u[2 + 8]);  u[2 + 12] = cmul(u[2 + 12], w6);
    u[3 + 4] = cmul(u[3 + 4], w3);  u[3 + 8] = cmul(u[3 + 8], w6);   u[3 + 12] = cmul(u[3 + 12], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4<INV>(u[4 * k1], u[4 * k1 + 1], u[4 * k1 + 2], u[4 * k1 + 3]);   // over b: k2 at u[4 k1 + k2]
    // u[4 k1 + k2] holds X[k1 + 4 k2]: transpose to natural order
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; k2++) { const float2 t = u[4 * k1 + k2]; u[4 * k1 + k2] = u[4 * k2 + k1]; u[4 * k2 + k1] = t; }
}

// twiddles W_L^(j p), p = 1..R-1, from the table tw[k] = exp(-2 pi i k / NT): the powers of two are read,
// the others are products
// Radix-16 passes read their four table twiddles from COMPACT records behind the natural table (tw + NT): the record of
// butterfly offset j of the pass with sub-length L is {W_L^j, W_L^2j, W_L^4j, W_L^8j} (32 bytes), records of one pass are
// contiguous in j, passes follow each other from L = NT downwards.  In the natural table the four entries of one lane are
// j NT/L, 2j NT/L, ... entries apart from the next lane's: at L = 1024 of a 16384-point transform every lane of every load
// touches its own 128-byte line — 128 L1 wavefronts per warp and load instruction, more LSU traffic than the pass's own
// shared-memory round trip (profiles/r02p_fftfilt2_full_cfg3.md: LSU data pipe 71 %, issue 34 %).
__host__ __device__ inline unsigned compact_offset(unsigned NT, unsigned L)      // in records; L = NT >> (4 k)
{
    unsigned off = 0;
    for (unsigned l = NT; l > L; l >>= 4) off += l >> 4;
    return off;
}
// records of the INNER radix-16 passes (L < NT: at most 64 + 4 of them) are staged in shared memory once per CTA: a pass
// between two barriers has nothing to hide a global-memory round trip behind (one CTA per SM at 16384 points)
constexpr unsigned INNER_RECORDS = 72;
__device__ __forceinline__ void stage_inner_twiddles(const float2* __restrict__ tw, unsigned NT, float4* __restrict__ twi, unsigned tid)
{
    const unsigned first = NT >> 4;                                  // records of the L = NT pass come first
    const Plan P = make_plan(NT);
    unsigned n = 0, l = NT >> 4;
    for (unsigned k = 1; k < P.n16; k++, l >>= 4) n += l >> 4;       // records of the other radix-16 passes: L = NT/16, NT/256, ...
    const float4* src = reinterpret_cast<const float4*>(tw + NT) + 2 * (size_t)first;
    for (unsigned i = tid; i < 2 * n && i < 2 * INNER_RECORDS; i += THREADS) twi[i] = __ldg(src + i);
}
template <int R>
__device__ __forceinline__ void pass_twiddles(const float2* __restrict__ tw, unsigned j, unsigned tws, float2 (&w)[R],
                                              unsigned NT = 0, unsigned L = 0, const float4* __restrict__ twi = nullptr)
{
    w[0] = make_float2(1.f, 0.f);
    if constexpr (R == 16) {
        if (NT) {
            float4 a, b;
            if (twi && L < NT) {
                const float4* rec = twi + 2 * (compact_offset(NT, L) - (NT >> 4) + j);
                a = rec[0]; b = rec[1];
            } else {
                const float4* rec = reinterpret_cast<const float4*>(tw + NT) + 2 * (size_t)(compact_offset(NT, L) + j);
                a = __ldg(rec); b = __ldg(rec + 1);
            }
            w[1] = make_float2(a.x, a.y); w[2] = make_float2(a.z, a.w); w[4] = make_float2(b.x, b.y); w[8] = make_float2(b.z, b.w);
            w[3] = cmul(w[1], w[2]);
            w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
#pragma unroll
            for (int p = 1; p < 8; p++) w[8 + p] = cmul(w[p], w[8]);
            return;
        }
    }
    if (R >= 2) w[1] = __ldg(tw + j * tws);
    if (R >= 4) { w[2] = __ldg(tw + 2 * j * tws); w[3] = cmul(w[1], w[2]); }
    if (R >= 16) {
        w[4] = __ldg(tw + 4 * j * tws); w[8] = __ldg(tw + 8 * j * tws);
        w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
#pragma unroll
        for (int p = 1; p < 8; p++) w[8 + p] = cmul(w[p], w[8]);
    }
}

template <int R, bool INV>
__device__ __forceinline__ void dft_r(float2 (&u)[R])
{
    if constexpr (R == 16) dft16<INV>(u);
    else if constexpr (R == 4) dft4<INV>(u[0], u[1], u[2], u[3]);
    else { const float2 a = u[0], b = u[1]; u[0] = cadd(a, b); u[1] = csub(a, b); }
}

// Padded address of element base + q*s of a butterfly (base = blk*L + j, j < s, L = R*s): when s >= 64 the pad
// term is linear in q, and when the whole butterfly lies inside one 64-point group (L <= 64) it is constant, so
// the R addresses are p0 + q*stride with one pad() per butterfly; otherwise (s < 64 < L) every address is padded.
template <bool LINEAR>
__device__ __forceinline__ unsigned pad_at(unsigned p0, unsigned base, unsigned q, unsigned s, unsigned stride)
{
    return LINEAR ? p0 + q * stride : pad(base + q * s);
}
__device__ __forceinline__ bool pad_linear(unsigned L, unsigned s) { return s >= 64 || L <= 64; }
__device__ __forceinline__ unsigned pad_stride(unsigned s) { return s >= 64 ? s + 4u * (s >> 6) : s; }

// one in-place pass over the padded shared buffer: sub-length L, radix R, stride s = L / R.
// forward (DIF): u <- DFT_R(u), u[p] *= W_L^(j p).   inverse (DIT): u[p] *= conj(W_L^(j p)), u <- IDFT_R(u).
template <int R, bool INV, bool LINEAR>
__device__ __forceinline__ void smem_pass_impl(float2* __restrict__ buf, unsigned N, unsigned L, const float2* __restrict__ tw,
                                               unsigned NT, unsigned tid, const float4* __restrict__ twi)
{
    const unsigned s = L / R, tws = NT / L, stride = pad_stride(s);
    for (unsigned t = tid; t < N / R; t += THREADS) {
        const unsigned j = t & (s - 1), base = (t - j) * R + j, p0 = pad(base);
        float2 u[R], w[R];
#pragma unroll
        for (int q = 0; q < R; q++) u[q] = buf[pad_at<LINEAR>(p0, base, q, s, stride)];
        pass_twiddles<R>(tw, j, tws, w, NT, L, twi);
        if (INV) {
#pragma unroll
            for (int p = 1; p < R; p++) u[p] = cmulc(u[p], w[p]);
        }
        dft_r<R, INV>(u);
        if (!INV) {
#pragma unroll
            for (int p = 1; p < R; p++) u[p] = cmul(u[p], w[p]);
        }
#pragma unroll
        for (int q = 0; q < R; q++) buf[pad_at<LINEAR>(p0, base, q, s, stride)] = u[q];
    }
}
template <int R, bool INV>
__device__ __forceinline__ void smem_pass(float2* __restrict__ buf, unsigned N, unsigned L, const float2* __restrict__ tw,
                                          unsigned NT, unsigned tid, const float4* __restrict__ twi)
{
    if (pad_linear(L, L / R)) smem_pass_impl<R, INV, true>(buf, N, L, tw, NT, tid, twi);
    else smem_pass_impl<R, INV, false>(buf, N, L, tw, NT, tid, twi);
}

// first forward pass (L = N) with its inputs read from global memory
template <int R>
__device__ __forceinline__ void first_pass_from_global(const float2* __restrict__ x, float2* __restrict__ buf, unsigned N,
                                                       const float2* __restrict__ tw, unsigned NT, unsigned tid)
{
    const unsigned s = N / R, tws = NT / N, stride = pad_stride(s);
    const bool lin = pad_linear(N, s);
    for (unsigned t = tid; t < s; t += THREADS) {
        float2 u[R], w[R];
#pragma unroll
        for (int q = 0; q < R; q++) u[q] = x[t + q * s];
        pass_twiddles<R>(tw, t, tws, w, NT, N);
        dft_r<R, false>(u);
#pragma unroll
        for (int p = 1; p < R; p++) u[p] = cmul(u[p], w[p]);
        const unsigned p0 = pad(t);
        if (lin) {
#pragma unroll
            for (int q = 0; q < R; q++) buf[p0 + q * stride] = u[q];
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) buf[pad(t + q * s)] = u[q];
        }
    }
}

// last inverse pass (L = N) writing the second half (the kept n samples), scaled, to global memory
template <int R>
__device__ __forceinline__ void last_pass_to_global(float2* __restrict__ buf, float2* __restrict__ y, unsigned N,
                                                    const float2* __restrict__ tw, unsigned NT, float scale, unsigned tid)
{
    const unsigned s = N / R, tws = NT / N, stride = pad_stride(s);
    const bool lin = pad_linear(N, s);
    for (unsigned t = tid; t < s; t += THREADS) {
        float2 u[R], w[R];
        const unsigned p0 = pad(t);
        if (lin) {
#pragma unroll
            for (int q = 0; q < R; q++) u[q] = buf[p0 + q * stride];
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) u[q] = buf[pad(t + q * s)];
        }
        pass_twiddles<R>(tw, t, tws, w, NT, N);
#pragma unroll
        for (int p = 1; p < R; p++) u[p] = cmulc(u[p], w[p]);
        dft_r<R, true>(u);
#pragma unroll
        for (int q = R / 2; q < R; q++) {     // natural index t + q s >= N/2
            y[t + q * s - N / 2] = make_float2(u[q].x * scale, u[q].y * scale);
        }
    }
}

// middle: forward radix-4 (L = 4), multiply by H, inverse radix-4 — four adjacent points per thread
__device__ __forceinline__ void middle_pass(float2* __restrict__ buf, unsigned N, const float2* __restrict__ H, unsigned tid)
{
    for (unsigned t = tid; t < N / 4; t += THREADS) {
        float4* p = reinterpret_cast<float4*>(buf + pad(4 * t));
        const float4 v0 = p[0], v1 = p[1];
        float2 a = make_float2(v0.x, v0.y), b = make_float2(v0.z, v0.w), c = make_float2(v1.x, v1.y), d = make_float2(v1.z, v1.w);
        dft4<false>(a, b, c, d);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(H + 4 * t)), h1 = __ldg(reinterpret_cast<const float4*>(H + 4 * t) + 1);
        a = cmul(a, make_float2(h0.x, h0.y)); b = cmul(b, make_float2(h0.z, h0.w));
        c = cmul(c, make_float2(h1.x, h1.y)); d = cmul(d, make_float2(h1.z, h1.w));
        dft4<true>(a, b, c, d);
        p[0] = make_float4(a.x, a.y, b.x, b.y);
        p[1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

// passes between the first/last (L = N) pass and the middle: sub-lengths go N/R0, ... down to 4 (exclusive)
template <bool INV>
__device__ __forceinline__ void inner_passes(float2* __restrict__ buf, const Plan& P, unsigned L_after_first, const float2* __restrict__ tw,
                                             unsigned tid, const float4* __restrict__ twi)
{
    // radices in forward order after the first pass
    unsigned n16 = P.n16, n4 = P.n4, n2 = P.n2;
    // the first pass consumed the widest available radix
    if (n16) n16--; else if (n4) n4--; else if (n2) n2--;
    if (!INV) {
        unsigned L = L_after_first;
        for (unsigned i = 0; i < n16; i++, L >>= 4) { smem_pass<16, false>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
        for (unsigned i = 0; i < n4; i++, L >>= 2) { smem_pass<4, false>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
        for (unsigned i = 0; i < n2; i++, L >>= 1) { smem_pass<2, false>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
    } else {
        // reverse order, starting just above the middle radix-4: radix-2 passes first, then radix-4, then radix-16
        unsigned L = 4;
        for (unsigned i = 0; i < n2; i++) { L <<= 1; smem_pass<2, true>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
        for (unsigned i = 0; i < n4; i++) { L <<= 2; smem_pass<4, true>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
        for (unsigned i = 0; i < n16; i++) { L <<= 4; smem_pass<16, true>(buf, P.N, L, tw, P.N, tid, twi); __syncthreads(); }
    }
}

// One CTA = one block of the filter.  x points at block 0's first sample (history below).
__global__ void __launch_bounds__(THREADS) fftfilt2_kernel(const float2* __restrict__ x, float2* __restrict__ y, unsigned N,
                                                           const float2* __restrict__ tw, const float2* __restrict__ H, float scale,
                                                           unsigned wave)
{
    extern __shared__ __align__(16) float2 fbuf[];
    __shared__ float4 twi[2 * INNER_RECORDS];
    const unsigned tid = threadIdx.x, B = N >> 1;
    const Plan P = make_plan(N);
    stage_inner_twiddles(tw, N, twi, tid);          // (read after the first pass's barrier)
    {
        // One CTA per SM (16384 points fill the shared memory) leaves nothing to overlap the first pass's DRAM round trip
        // with: pull the new half of the window of the block this SM will most likely run next — one wave of CTAs ahead —
        // into L2 now (the other half is this wave's input and already there)
        const unsigned long long nb = (unsigned long long)blockIdx.x + wave;
        if (wave && nb < gridDim.x) {
            const char* p = reinterpret_cast<const char*>(x + (nb - 0) * (unsigned long long)B);
            for (unsigned o = tid * 128u; o < B * (unsigned)sizeof(float2); o += THREADS * 128u)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p + o));
        }
    }
    const float2* w = x + ((long long)blockIdx.x - 1) * (long long)B;
    float2* out = y + (size_t)blockIdx.x * B;
    unsigned L;
    if (P.n16) { first_pass_from_global<16>(w, fbuf, N, tw, N, tid); L = N >> 4; }
    else if (P.n4) { first_pass_from_global<4>(w, fbuf, N, tw, N, tid); L = N >> 2; }
    else { first_pass_from_global<2>(w, fbuf, N, tw, N, tid); L = N >> 1; }
    __syncthreads();
    inner_passes<false>(fbuf, P, L, tw, tid, twi);
    middle_pass(fbuf, N, H, tid);
    __syncthreads();
    inner_passes<true>(fbuf, P, L, tw, tid, twi);
    if (P.n16) last_pass_to_global<16>(fbuf, out, N, tw, N, scale, tid);
    else if (P.n4) last_pass_to_global<4>(fbuf, out, N, tw, N, scale, tid);
    else last_pass_to_global<2>(fbuf, out, N, tw, N, scale, tid);
}

// forward network only (H generation): in -> out in the network's digit-reversed order
__global__ void __launch_bounds__(THREADS) fft2_forward_kernel(const float2* __restrict__ in, float2* __restrict__ out, unsigned N,
                                                               const float2* __restrict__ tw)
{
    extern __shared__ __align__(16) float2 fbuf[];
    __shared__ float4 twi[2 * INNER_RECORDS];
    const unsigned tid = threadIdx.x;
    const Plan P = make_plan(N);
    stage_inner_twiddles(tw, N, twi, tid);
    unsigned L;
    if (P.n16) { first_pass_from_global<16>(in, fbuf, N, tw, N, tid); L = N >> 4; }
    else if (P.n4) { first_pass_from_global<4>(in, fbuf, N, tw, N, tid); L = N >> 2; }
    else { first_pass_from_global<2>(in, fbuf, N, tw, N, tid); L = N >> 1; }
    __syncthreads();
    inner_passes<false>(fbuf, P, L, tw, tid, twi);
    for (unsigned t = tid; t < N / 4; t += THREADS) {
        float2 a = fbuf[pad(4 * t)], b = fbuf[pad(4 * t + 1)], c = fbuf[pad(4 * t + 2)], d = fbuf[pad(4 * t + 3)];
        dft4<false>(a, b, c, d);
        out[4 * t] = a; out[4 * t + 1] = b; out[4 * t + 2] = c; out[4 * t + 3] = d;
    }
}

}  // namespace fft2
}  // namespace iqgpu
