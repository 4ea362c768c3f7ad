// design.cpp — K0 host-side design (see design.hpp).  Compile with -ffp-contract=off: the
// float evaluation order below is deliberate (it mirrors liquid-dsp's constructors so the GPU
// taps equal the taps the reference's liquid objects would hold).
#include "design.hpp"

#include <cmath>
#include <cstring>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace iqgpu {

// -------------------------------------------------------------------------------------------
// liquid-dsp math helpers (src/math/src/math.gamma.c, math.bessel.c, windows.c; firdes.c)
// All mixed float/double promotions are written out explicitly.
// -------------------------------------------------------------------------------------------
float lq_lngammaf(float z)
{
    if (z < 0.0f) return 0.0f;
    if (z < 10.0f) return lq_lngammaf(z + 1.0f) - ::logf(z);
    // g = 0.5*( logf(2*pi) - log(z) )  evaluated in double, stored as float
    float g = (float)(0.5 * ((double)::logf((float)(2.0 * M_PI)) - ::log((double)z)));
    float inner = z + (1.0f / (12.0f * z - 0.1f / z));
    g += z * (::logf(inner) - 1.0f);
    return g;
}

float lq_besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    for (unsigned k = 0; k < 32; k++) {
        float t = (float)k * ::logf(0.5f * z) - lq_lngammaf((float)k + 1.0f);
        y += ::expf(2.0f * t);
    }
    return y;
}

float lq_kaiser_beta(float as)
{
    as = ::fabsf(as);
    if (as > 50.0f) return 0.1102f * (as - 8.7f);
    if (as > 21.0f) return (float)(0.5842 * (double)::powf(as - 21.0f, 0.4f) + (double)(0.07886f * (as - 21.0f)));
    return 0.0f;
}

float lq_kaiser(unsigned i, unsigned wlen, float beta)
{
    float t = (float)i - (float)(wlen - 1) / 2.0f;
    float r = 2.0f * t / (float)(wlen - 1);
    float a = lq_besseli0f(beta * ::sqrtf(1.0f - r * r));
    float b = lq_besseli0f(beta);
    return a / b;
}

float lq_sincf(float x)
{
    if (::fabsf(x) < 0.01f) {
        float c2 = ::cosf((float)(M_PI * (double)x / 2.0));
        float c4 = ::cosf((float)(M_PI * (double)x / 4.0));
        float c8 = ::cosf((float)(M_PI * (double)x / 8.0));
        return c2 * c4 * c8;
    }
    double px = M_PI * (double)x;
    return (float)((double)::sinf((float)px) / px);
}

unsigned lq_estimate_req_filter_len(float df, float as)
{
    if (df > 0.5f || df <= 0.0f || as <= 0.0f) return 0;
    float len = (as - 7.95f) / (14.26f * df);
    return (unsigned)len;
}

bool lq_firdes_kaiser(unsigned n, float fc, float as, float mu, float* h)
{
    if (mu < -0.5f || mu > 0.5f || fc <= 0.0f || fc > 0.5f || n == 0) return false;
    float beta = lq_kaiser_beta(as);
    for (unsigned i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2.0f + mu;
        float h1 = lq_sincf(2.0f * fc * t);
        float h2 = lq_kaiser(i, n, beta);
        h[i] = h1 * h2;
    }
    return true;
}

// -------------------------------------------------------------------------------------------
// NCO (src/nco/src/nco.proto.c, LIQUID_NCO)
// -------------------------------------------------------------------------------------------
uint32_t nco_constrain(float theta)
{
    float p = (float)((double)theta * 0.159154943091895);
    float fpart = p - (float)((long)p);
    if (fpart < 0.0f) fpart += 1.0f;
    float scaled = fpart * 4294967296.0f;  // (float)0xffffffff
    return (uint32_t)(int64_t)scaled;
}

void nco_sine_table(float* tab)
{
    for (unsigned i = 0; i < 1024; i++)
        tab[i] = ::sinf((float)((double)(2.0f) * M_PI * (double)(float)i / (double)1024.0f));
}

uint32_t nco_dtheta_for_shift(double shift_hz, double rate)
{
    float w = (float)(2.0 * M_PI * std::fabs(shift_hz) / rate);
    return nco_constrain(w);
}

// -------------------------------------------------------------------------------------------
// DC blocker
// -------------------------------------------------------------------------------------------
DcPlan design_dc(bool enable, int samplerate)
{
    DcPlan p;
    p.enable = enable;
    if (!enable) return p;
    p.alpha = (float)(2.0 * M_PI * (double)10.0f / (double)samplerate);
    float a1 = -1.0f + p.alpha;   // liquid: a = {1, -1 + alpha}
    p.c = -a1;
    p.one_minus_c = 1.0f - p.c;   // exact (Sterbenz)
    return p;
}

// -------------------------------------------------------------------------------------------
// resampler
// -------------------------------------------------------------------------------------------
static void design_halfband(unsigned m, float f0, float as, HalfbandStage& st)
{
    st.m = m;
    st.as = as;
    const unsigned h_len = 4 * m + 1;
    st.h.resize(h_len);
    float beta = lq_kaiser_beta(as);
    for (unsigned i = 0; i < h_len; i++) {
        float t = (float)i - (float)(h_len - 1) / 2.0f;
        float h1 = lq_sincf(t / 2.0f);
        float h2 = lq_kaiser(i, h_len, beta);
        float h3 = ::cosf((float)((double)2.0f * M_PI * (double)t * (double)f0));
        st.h[i] = h1 * h2 * h3;
    }
    st.h1.resize(2 * m);
    unsigned j = 0;
    for (unsigned i = 1; i < h_len; i += 2) st.h1[j++] = st.h[h_len - i - 1];
}

bool design_resampler(float r, float as, bool passthrough, ResamplerPlan& p, std::string& err)
{
    p = ResamplerPlan();
    p.ratio = r;
    if (passthrough) {
        p.passthrough = true;
        return true;
    }
    if (!(r > 0.0f)) {
        err = "resample ratio must be positive";
        return false;
    }
    p.is_interp = r > 1.0f;
    float rate_arb = r;
    unsigned S = 0;
    if (p.is_interp) {
        while (rate_arb > 2.0f) { S++; rate_arb *= 0.5f; }
    } else {
        while (rate_arb < 0.5f) { S++; rate_arb *= 2.0f; }
    }
    if (S > 16) {
        err = "too many halfband stages";
        return false;
    }
    p.num_halfband = S;
    p.rate_arbitrary = rate_arb;
    p.zeta = 1.0f / (float)(1u << S);

    // msresamp2 design loop (fc = 0.4, f0 = 0, As + 5 dB margin)
    float fc = 0.4f, f0 = 0.0f;
    const float as_m = as + 5.0f;
    p.stages.resize(S);
    for (unsigned i = 0; i < S; i++) {
        fc = (i == 1) ? (float)((0.5 - (double)fc) / (double)2.0f) : 0.5f * fc;
        f0 = 0.5f * f0;
        float ft = 2.0f * (0.25f - fc);
        unsigned h_len = lq_estimate_req_filter_len(ft, as_m);
        unsigned m = (unsigned)::ceilf((float)(h_len - 1) / 4.0f);
        if (m < 3) m = 3;
        design_halfband(m, f0, as_m, p.stages[i]);
    }

    // arbitrary stage: resamp_crcf_create(rate_arb, 7, min(0.515 rate, 0.49), As, 256)
    float afc = 0.515f * rate_arb;
    if (afc > 0.49f) afc = 0.49f;
    p.arb_fc = afc;
    p.arb_m = 7;
    p.npfb = 256;
    p.arb_sub_len = 2 * p.arb_m;
    float stepf = (float)(1 << 24) / rate_arb;
    p.step = (uint32_t)std::round((double)stepf);
    const unsigned n = 2 * p.arb_m * p.npfb + 1;
    std::vector<float> hf(n);
    if (!lq_firdes_kaiser(n, afc / (float)p.npfb, as, 0.0f, hf.data())) {
        err = "arbitrary resampler prototype design failed";
        return false;
    }
    float gain = 0.0f;
    for (unsigned i = 0; i < n; i++) gain += hf[i];
    gain = (float)p.npfb / gain;
    p.arb_h.resize(n - 1);
    for (unsigned i = 0; i + 1 < n; i++) p.arb_h[i] = hf[i] * gain;
    p.bank.resize((size_t)p.npfb * p.arb_sub_len);
    for (unsigned i = 0; i < p.npfb; i++)
        for (unsigned k = 0; k < p.arb_sub_len; k++)
            p.bank[(size_t)i * p.arb_sub_len + (p.arb_sub_len - k - 1)] = p.arb_h[i + k * p.npfb];

    // history a stateless recomputation needs, in input frames (decimation):
    //   stage executed at depth d (input rate Fs/2^d) looks back 4m samples of its own input,
    //   the arbitrary stage looks back 13 decimated samples.
    if (!p.is_interp) {
        uint64_t halo = 0;
        for (unsigned d = 0; d < S; d++) {
            unsigned g = S - 1 - d;  // design index executed at depth d
            halo += (uint64_t)(4 * p.stages[g].m) << d;
        }
        halo += (uint64_t)(p.arb_sub_len) << S;
        p.halo_input_frames = halo;
    } else {
        // interpolation: the arbitrary stage runs first, at the input rate, over a window of 2*7 input frames; interpolator s
        // (design index s, input rate rate_arb * 2^s > 2^s input rates) looks back 2 m_s samples of its own input, i.e. fewer
        // than 2 m_s / 2^s input frames; two frames of slack per stage for the rounding of the mapped-back positions
        uint64_t halo = p.arb_sub_len;
        for (unsigned s = 0; s < S; s++) halo += ((uint64_t)(2 * p.stages[s].m) + ((1ull << s) - 1)) / (1ull << s) + 2;
        p.halo_input_frames = halo;
    }
    return true;
}

uint64_t arb_outputs_after(uint64_t pushes, uint32_t step)
{
    // ceil(pushes * 2^24 / step); 64-bit arithmetic while it fits (the per-chunk bookkeeping calls this once per chunk),
    // 128-bit intermediate from 2^39 pushes on (pushes * 2^24 + step - 1 must stay below 2^64, and step < 2^32)
    if (pushes < (1ull << 39)) return ((pushes << 24) + step - 1) / step;
    unsigned __int128 num = (unsigned __int128)pushes << 24;
    return (uint64_t)((num + step - 1) / step);
}

uint64_t resampler_outputs_after(const ResamplerPlan& p, uint64_t n_in)
{
    if (p.passthrough) return n_in;
    if (!p.is_interp) return arb_outputs_after(n_in >> p.num_halfband, p.step);
    return arb_outputs_after(n_in, p.step) << p.num_halfband;
}

// -------------------------------------------------------------------------------------------
// user filter (reference src/filter.c:43-393)
// -------------------------------------------------------------------------------------------
static inline cfloat cmul(cfloat a, cfloat b)
{
    return cfloat(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}

bool design_filter(const iqgpu_chain_config& g, int in_rate, double target_rate, FilterPlan& out, std::string& err)
{
    out = FilterPlan();
    if (g.num_filter_requests <= 0) return true;
    if (g.num_filter_requests > IQGPU_MAX_FILTER_CHAIN) {
        err = "too many filter requests";
        return false;
    }
    // placement (filter.c:43-92)
    out.post_resample = false;
    if (!g.no_resample && target_rate < (double)in_rate) {
        float fmax = 0.0f;
        for (int i = 0; i < g.num_filter_requests; i++) {
            const iqgpu_filter_request& r = g.filter_requests[i];
            float cur = 0.0f;
            if (r.type == IQGPU_FILTER_LOWPASS || r.type == IQGPU_FILTER_HIGHPASS) cur = ::fabsf(r.freq1_hz);
            else if (r.type == IQGPU_FILTER_PASSBAND || r.type == IQGPU_FILTER_STOPBAND) cur = ::fabsf(r.freq1_hz) + (r.freq2_hz / 2.0f);
            if (cur > fmax) fmax = cur;
        }
        if ((double)fmax > target_rate / 2.0) {
            err = "filter configuration is incompatible with the output sample rate";
            return false;
        }
        out.post_resample = true;
    }
    const double fs = out.post_resample ? target_rate : (double)in_rate;
    const float fsf = (float)fs;

    std::vector<cfloat> master(1, cfloat(1.0f, 0.0f));
    bool is_complex = false, by_peak = false;
    for (int i = 0; i < g.num_filter_requests; i++) {
        const iqgpu_filter_request& r = g.filter_requests[i];
        if (r.type != IQGPU_FILTER_LOWPASS) by_peak = true;
        const float as = (g.attenuation_db > 0.0f) ? g.attenuation_db : 60.0f;
        unsigned len;
        if (g.filter_taps > 0) {
            len = (unsigned)g.filter_taps;
        } else {
            float tw;
            if (g.transition_width_hz > 0.0f) tw = g.transition_width_hz;
            else {
                float ref = (r.type == IQGPU_FILTER_LOWPASS || r.type == IQGPU_FILTER_HIGHPASS) ? r.freq1_hz : r.freq2_hz;
                tw = ::fabsf(ref) * 0.25f;
            }
            if (tw < 1.0f) tw = 1.0f;
            float ntw = tw / fsf;
            len = lq_estimate_req_filter_len(ntw, as);
            if (len % 2 == 0) len++;
            if (len < 21) len = 21;
        }
        std::vector<float> rt(len);
        std::vector<cfloat> cur(len);
        const bool stage_complex = (r.type == IQGPU_FILTER_PASSBAND && ::fabsf(r.freq1_hz) > 1e-9f);
        bool ok = true;
        if (stage_complex) {
            is_complex = true;
            float hbw = (r.freq2_hz / 2.0f) / fsf;
            ok = lq_firdes_kaiser(len, hbw, as, 0.0f, rt.data());
            float fcn = r.freq1_hz / fsf;
            // modulation by a LIQUID_NCO (phase-quantised, filter.c:211-217)
            float sintab[1024];
            nco_sine_table(sintab);
            uint32_t theta = 0, dtheta = nco_constrain((float)((double)2.0f * M_PI * (double)fcn));
            for (unsigned k = 0; k < len; k++) {
                unsigned idx = ((theta + (1u << 21)) >> 22) & 0x3ff;
                float s = sintab[idx], c = sintab[(idx + 256) & 0x3ff];
                cur[k] = cfloat(c * rt[k], s * rt[k]);
                theta += dtheta;
            }
        } else {
            float fc, bw;
            switch (r.type) {
                case IQGPU_FILTER_LOWPASS:
                    fc = r.freq1_hz / fsf; ok = lq_firdes_kaiser(len, fc, as, 0.0f, rt.data()); break;
                case IQGPU_FILTER_HIGHPASS:
                    fc = r.freq1_hz / fsf; ok = lq_firdes_kaiser(len, fc, as, 0.0f, rt.data());
                    for (unsigned k = 0; k < len; k++) rt[k] = -rt[k];
                    rt[(len - 1) / 2] += 1.0f;
                    break;
                case IQGPU_FILTER_PASSBAND:
                    bw = r.freq2_hz / fsf; ok = lq_firdes_kaiser(len, bw / 2.0f, as, 0.0f, rt.data()); break;
                case IQGPU_FILTER_STOPBAND:  // centre ignored: notch always at DC (filter.c:237-241)
                    bw = r.freq2_hz / fsf; ok = lq_firdes_kaiser(len, bw / 2.0f, as, 0.0f, rt.data());
                    for (unsigned k = 0; k < len; k++) rt[k] = -rt[k];
                    rt[(len - 1) / 2] += 1.0f;
                    break;
                default:
                    std::fill(rt.begin(), rt.end(), 0.0f); break;
            }
            for (unsigned k = 0; k < len; k++) cur[k] = cfloat(rt[k], 0.0f);
        }
        if (!ok) {
            err = "filter stage design failed (cutoff outside (0, 0.5] of the design rate?)";
            return false;
        }
        const int mlen = (int)master.size(), clen = (int)len;
        std::vector<cfloat> nm((size_t)(mlen + clen - 1), cfloat(0.0f, 0.0f));
        for (int a = 0; a < mlen + clen - 1; a++) {
            int j0 = (a >= mlen) ? (a - mlen + 1) : 0;
            int j1 = (a < clen - 1) ? a : (clen - 1);
            cfloat acc(0.0f, 0.0f);
            for (int j = j0; j <= j1; j++) acc += cmul(master[a - j], cur[j]);
            nm[a] = acc;
        }
        master.swap(nm);
    }

    // gain normalisation (filter.c:272-299)
    const int mlen = (int)master.size();
    if (by_peak || is_complex) {
        float peak = 0.0f;
        for (int i = 0; i < 2048; i++) {
            float f = ((float)i / 2048.0f) - 0.5f;
            // liquid firfilt_cccf_freqresponse works on its reversed tap copy
            cfloat H(0.0f, 0.0f);
            for (int k = 0; k < mlen; k++) {
                float ang = (float)(2.0 * M_PI * (double)f * (double)k);
                H += cmul(master[mlen - 1 - k], cfloat(::cosf(ang), ::sinf(ang)));
            }
            float mag = ::hypotf(H.real(), H.imag());
            if (mag > peak) peak = mag;
        }
        if (peak > 1e-9f)
            for (auto& t : master) t = cfloat(t.real() / peak, t.imag() / peak);
    } else {
        double dc = 0.0;
        for (auto& t : master) dc += (double)t.real();
        if (std::fabs(dc) > (double)1e-9f) {
            float d = (float)dc;
            for (auto& t : master) t = cfloat(t.real() / d, t.imag() / d);
        }
    }

    bool want_fft;
    if (g.filter_type_request != IQGPU_FILTER_REQ_AUTO) want_fft = (g.filter_type_request == IQGPU_FILTER_REQ_FFT);
    else want_fft = is_complex;
    out.is_complex = is_complex;
    if (want_fft) {
        unsigned block;
        if (g.filter_fft_size > 0) {
            block = (unsigned)g.filter_fft_size / 2;
            if (block < (unsigned)mlen - 1) {
                err = "--filter-fft-size too small for the number of taps";
                return false;
            }
        } else {
            block = 1;
            while (block < (unsigned)mlen - 1) block *= 2;
            if (block < (unsigned)mlen * 2) block *= 2;
        }
        if (block > 1024u * 1024u) {
            err = "FFT block exceeds MAX_ALLOWED_FFT_BLOCK_SIZE";
            return false;
        }
        out.block = block;
        out.impl = is_complex ? IQGPU_FILTER_IMPL_FFT_ASYM : IQGPU_FILTER_IMPL_FFT_SYM;
    } else {
        out.impl = is_complex ? IQGPU_FILTER_IMPL_FIR_ASYM : IQGPU_FILTER_IMPL_FIR_SYM;
    }
    if (!is_complex)
        for (auto& t : master) t = cfloat(t.real(), 0.0f);
    out.taps.swap(master);
    return true;
}

}  // namespace iqgpu
