/*
 * iq_oracle.c — TEST INFRASTRUCTURE ONLY (the parity oracle). Not part of the product;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * A plain-C restatement of pclov3r/iq_tool's per-block sample-processing chain
 * (convert -> DC block -> I/Q apply -> NCO shift -> [filter] -> resample -> [filter] ->
 * NCO shift -> AGC -> convert), written as one stream object instead of the reference's
 * AppResources/SampleChunk plumbing.  Arithmetic owned by the reference is restated here
 * (each function cites the reference file:line it follows); arithmetic the reference
 * delegates to liquid-dsp is taken from oracle/liquid_compat (the restated liquid layer).
 *
 * PINNING: this oracle is pinned against the reference's OWN code by tests/test_oracle_*.py:
 *   - bit-exact against oracle/_ref/libiqref.so (reference sources compiled in place, same
 *     liquid_compat underneath) on every BASELINE config and on the conversion KATs;
 *   - against golden vectors generated from that build (tests/golden/, tools/make_golden.py).
 * The liquid layer underneath is itself PARITY UNPINNED (no real libliquid available here);
 * see oracle/README.md.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include "iq_oracle.h"
#include "liquid/liquid.h"

#include <complex.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef float complex cf;

/* constants restated from reference include/constants.h */
#define CHUNK            16384      /* :123 PIPELINE_CHUNK_BASE_SAMPLES */
#define OUT_MARGIN       128        /* :129 RESAMPLER_OUTPUT_SAFETY_MARGIN */
#define RESAMP_AS_DB     60.0f      /* :137 */
#define TRANSITION_FACTOR 0.25f     /* :142 */
#define DC_CUTOFF_HZ     10.0f      /* :149 */
#define MIN_TAPS         21         /* :152 */
#define GAIN_ZERO_THRESH 1e-9f      /* :153 */
#define RESPONSE_POINTS  2048       /* :154 */
#define IQ_NFFT          1024       /* :157 */
#define IQ_INCREMENT     0.0001f    /* :159 */
#define IQ_PASSES        25         /* :160 */
#define IQ_POWER_THRESH  20.0f      /* :161 */
#define IQ_SMOOTHING     0.05f      /* :162 */
#define AGC_DX_BW        1e-4f      /* :169 */
#define AGC_LOCAL_BW     1e-2f      /* :174 */
#define AGC_DIG_TARGET   0.9f       /* :184 */
#define AGC_DIG_LOCK_S   2.0f       /* :185 */
#define AGC_DIG_HANG_S   4.0f       /* :188 */
#define AGC_DIG_RECOVER  1.0005f    /* :191 */
#define AGC_DIG_LOWER    0.75f      /* :192 */
#define MIN_RATIO        0.001f     /* :245 */
#define MAX_RATIO        1000.0f    /* :246 */
#define SHIFT_LIMIT      5.0        /* :248 */

/* ---------------------------------------------------------------------------------------
 * clock (the reference reads a monotonic wall clock inside the digital AGC, agc.c:176)
 * ------------------------------------------------------------------------------------- */
static int    g_fake_clock_enabled = 0;
static double g_fake_clock = 0.0;
void iqo_set_fake_clock(int enable, double t) { g_fake_clock_enabled = enable; g_fake_clock = t; }
static double now_sec(void)
{
    if (g_fake_clock_enabled) return g_fake_clock;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec / 1e9;
}

/* =======================================================================================
 * C3 / C1 / C2 — sample conversion           reference src/sample_convert.c:102,127,213
 * ===================================================================================== */
size_t iqo_get_bytes_per_sample(int fmt)
{
    switch (fmt) { /* sample_convert.c:103-122 */
        case IQF_S8: case IQF_U8: return 1;
        case IQF_S16: case IQF_U16: return 2;
        case IQF_S32: case IQF_U32: case IQF_F32: return 4;
        case IQF_CS8: case IQF_CU8: return 2;
        case IQF_CS16: case IQF_CU16: case IQF_SC16Q11: return 4;
        case IQF_CS24: return 6;
        case IQF_CS32: case IQF_CU32: case IQF_CF32: return 8;
        default: return 0;
    }
}

/* signed ints: x * (1/2^k) * gain; unsigned: (x - mid) * (1/2^k) * gain   (:75-98,136-205) */
int iqo_convert_block_to_cf32(const void *in, float *out, size_t n, int fmt, float gain)
{
    size_t i;
    switch (fmt) {
        case IQF_CS8: {
            const int8_t *p = (const int8_t *)in; const float k = 1.0f / 128.0f;
            for (i = 0; i < 2 * n; i++) { float v = (float)p[i] * k; out[i] = v * gain; }
            return 0;
        }
        case IQF_CU8: {
            const uint8_t *p = (const uint8_t *)in; const float k = 1.0f / 128.0f;
            for (i = 0; i < 2 * n; i++) { float v = ((float)p[i] - 127.5f) * k; out[i] = v * gain; }
            return 0;
        }
        case IQF_CS16: {
            const int16_t *p = (const int16_t *)in; const float k = 1.0f / 32768.0f;
            for (i = 0; i < 2 * n; i++) { float v = (float)p[i] * k; out[i] = v * gain; }
            return 0;
        }
        case IQF_SC16Q11: {
            const int16_t *p = (const int16_t *)in; const float k = 1.0f / 2048.0f;
            for (i = 0; i < 2 * n; i++) { float v = (float)p[i] * k; out[i] = v * gain; }
            return 0;
        }
        case IQF_CU16: {
            const uint16_t *p = (const uint16_t *)in; const float k = 1.0f / 32768.0f;
            for (i = 0; i < 2 * n; i++) { float v = ((float)p[i] - 32767.5f) * k; out[i] = v * gain; }
            return 0;
        }
        case IQF_CS24: { /* 3-byte little-endian, sign-extended (:154-169) */
            const unsigned char *p = (const unsigned char *)in; const float k = 1.0f / 8388608.0f;
            for (i = 0; i < 2 * n; i++, p += 3) {
                int32_t v = (int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24));
                v >>= 8;
                out[i] = (float)v * k * gain;
            }
            return 0;
        }
        case IQF_CS32: { /* double intermediate (:174-183) */
            const int32_t *p = (const int32_t *)in; const double k = 1.0 / 2147483648.0;
            for (i = 0; i < 2 * n; i++) { double v = (double)p[i] * k; out[i] = (float)(v * gain); }
            return 0;
        }
        case IQF_CU32: { /* (:185-196) */
            const uint32_t *p = (const uint32_t *)in; const double k = 1.0 / 2147483648.0;
            for (i = 0; i < 2 * n; i++) { double v = ((double)p[i] - 2147483647.5) * k; out[i] = (float)(v * gain); }
            return 0;
        }
        case IQF_CF32: { /* complex * real gain (:198-203) */
            const float *p = (const float *)in;
            for (i = 0; i < 2 * n; i++) out[i] = p[i] * gain;
            return 0;
        }
        default: return -1;
    }
}

/* signed: v*S, +-0.5 away from zero, clamp AFTER the offset, truncate (:40-57);
 * unsigned: v*S+OFF, clamp to [0,MAX], (T)(v+0.5) (:59-73) */
#define OUT_SIGNED(T, VMAX, VMIN, S)                                                   \
    do {                                                                                \
        T *o = (T *)out; const float hi = (float)(VMAX), lo = (float)(VMIN);            \
        for (i = 0; i < 2 * n; i++) {                                                   \
            float v = in[i] * (S);                                                      \
            v = (v > 0.0f) ? v + 0.5f : v - 0.5f;                                       \
            if (v > hi) v = hi;                                                         \
            if (v < lo) v = lo;                                                         \
            o[i] = (T)v;                                                                \
        }                                                                               \
    } while (0)
#define OUT_UNSIGNED(T, VMAX, S, OFF)                                                  \
    do {                                                                                \
        T *o = (T *)out; const float hi = (float)(VMAX);                                \
        for (i = 0; i < 2 * n; i++) {                                                   \
            float v = (in[i] * (S)) + (OFF);                                            \
            if (v > hi) v = hi;                                                         \
            if (v < 0.0f) v = 0.0f;                                                     \
            o[i] = (T)(v + 0.5f);                                                       \
        }                                                                               \
    } while (0)

int iqo_convert_cf32_to_block(const float *in, void *out, size_t n, int fmt)
{
    size_t i;
    switch (fmt) {
        case IQF_CS8:     OUT_SIGNED(int8_t, SCHAR_MAX, SCHAR_MIN, (float)SCHAR_MAX); return 0;   /* :219 */
        case IQF_CU8:     OUT_UNSIGNED(uint8_t, UCHAR_MAX, 127.0f, 127.5f); return 0;             /* :222 */
        case IQF_CS16:    OUT_SIGNED(int16_t, SHRT_MAX, SHRT_MIN, (float)SHRT_MAX); return 0;     /* :225 */
        case IQF_SC16Q11: OUT_SIGNED(int16_t, SHRT_MAX, SHRT_MIN, 2048.0f); return 0;             /* :228 */
        case IQF_CU16:    OUT_UNSIGNED(uint16_t, USHRT_MAX, 32767.0f, 32767.5f); return 0;        /* :231 */
        case IQF_CS24: { /* int32 clamp after the rounding cast (:233-261) */
            unsigned char *o = (unsigned char *)out;
            for (i = 0; i < 2 * n; i++, o += 3) {
                float f = in[i] * 8388607.0f;
                int32_t v = (int32_t)((f > 0.0f) ? f + 0.5f : f - 0.5f);
                if (v > 8388607) v = 8388607;
                if (v < -8388608) v = -8388608;
                o[0] = (unsigned char)(v & 0xFF);
                o[1] = (unsigned char)((v >> 8) & 0xFF);
                o[2] = (unsigned char)((v >> 16) & 0xFF);
            }
            return 0;
        }
        case IQF_CS32: { /* double (:263-281) */
            int32_t *o = (int32_t *)out; const double hi = (double)INT_MAX, lo = (double)INT_MIN;
            for (i = 0; i < 2 * n; i++) {
                double v = (double)in[i] * hi;
                v = (v > 0.0) ? v + 0.5 : v - 0.5;
                if (v > hi) v = hi;
                if (v < lo) v = lo;
                o[i] = (int32_t)v;
            }
            return 0;
        }
        case IQF_CU32: { /* (:283-298) */
            uint32_t *o = (uint32_t *)out; const double hi = (double)UINT_MAX;
            for (i = 0; i < 2 * n; i++) {
                double v = ((double)in[i] * 2147483647.0) + 2147483647.5;
                if (v > hi) v = hi;
                if (v < 0.0) v = 0.0;
                o[i] = (uint32_t)(v + 0.5);
            }
            return 0;
        }
        case IQF_CF32: memcpy(out, in, n * 2 * sizeof(float)); return 0;                          /* :301 */
        default: return -1;
    }
}

/* =======================================================================================
 * Chain object
 * ===================================================================================== */
typedef struct {
    iq_chain_cfg cfg;
    int    in_rate;               /* source_info.samplerate */
    double target_rate;
    float  ratio;
    size_t in_bps, out_bps, cap;
    int    passthrough;

    /* D1 */
    iirfilt_crcf dc;
    /* Q1 */
    float iq_mag, iq_phase;
    float iq_window[IQ_NFFT];
    cf    iq_fft_buf[IQ_NFFT];
    float iq_spectrum[IQ_NFFT];
    fftplan iq_plan;
    float iq_avg_power, iq_power_range;
    /* N1 */
    nco_crcf nco_pre, nco_post;
    double   shift_hz;
    /* R1 */
    msresamp_crcf rs;
    /* F1..F3 */
    int      filt_impl, filt_post;
    unsigned filt_block, filt_len;
    void    *filt_obj;
    cf      *filt_taps;
    cf      *rem; unsigned rem_len;
    /* G1/G2 */
    agc_crcf agc_rms;
    int      agc_locked;
    float    agc_gain, agc_peak_mem;
    uint64_t agc_seen;
    double   agc_last_strong;

    /* work buffers (one chunk) */
    cf *buf_a, *buf_b;
    unsigned char *out_tmp;

    cf *cap_buf[3]; int64_t cap_cap[3], cap_len[3];
    uint32_t *trace; int64_t trace_cap, trace_len;
} chain_t;

/* ---------------------------------------------------------------------------------------
 * F1 — filter design                                   reference src/filter.c:43-393
 * ------------------------------------------------------------------------------------- */
static cf cmul(cf a, cf b)
{
    float ar = crealf(a), ai = cimagf(a), br = crealf(b), bi = cimagf(b);
    return (ar * br - ai * bi) + (ar * bi + ai * br) * _Complex_I;
}

/* placement rule: a down-sampling chain always filters AFTER the resampler (:43-92) */
static int filter_place(chain_t *c)
{
    const iq_chain_cfg *g = &c->cfg;
    c->filt_post = 0;
    if (g->num_filter_requests == 0 || g->no_resample) return 0;
    double in_rate = (double)c->in_rate, out_rate = c->target_rate;
    if (out_rate < in_rate) {
        float fmax = 0.0f;
        for (int i = 0; i < g->num_filter_requests; i++) {
            const iq_filter_request *r = &g->filter_requests[i];
            float cur = 0.0f;
            if (r->type == IQ_FILTER_LOWPASS || r->type == IQ_FILTER_HIGHPASS) cur = fabsf(r->freq1_hz);
            else if (r->type == IQ_FILTER_PASSBAND || r->type == IQ_FILTER_STOPBAND) cur = fabsf(r->freq1_hz) + (r->freq2_hz / 2.0f);
            if (cur > fmax) fmax = cur;
        }
        if (fmax > out_rate / 2.0) return -1; /* fatal in the reference (:80-84) */
        c->filt_post = 1;
    }
    return 0;
}

static void spectral_invert(float *t, unsigned len) /* :94-99 */
{
    for (unsigned k = 0; k < len; k++) t[k] = -t[k];
    t[(len - 1) / 2] += 1.0f;
}

static int filter_design(chain_t *c)
{
    const iq_chain_cfg *g = &c->cfg;
    c->filt_impl = IQ_FILTER_IMPL_NONE; c->filt_obj = NULL; c->filt_block = 0; c->filt_len = 0;
    if (g->num_filter_requests == 0) return 0;
    if (filter_place(c) != 0) return -1;

    int mlen = 1;
    cf *master = (cf *)malloc(sizeof(cf));
    master[0] = 1.0f;
    double fs = c->filt_post ? c->target_rate : (double)c->in_rate;           /* :162-164 */
    int is_complex = 0, by_peak = 0;

    for (int i = 0; i < g->num_filter_requests; i++) {                         /* :169-256 */
        const iq_filter_request *r = &g->filter_requests[i];
        if (r->type != IQ_FILTER_LOWPASS) by_peak = 1;
        float as = (g->attenuation_db > 0.0f) ? g->attenuation_db : RESAMP_AS_DB;
        unsigned len;
        if (g->filter_taps > 0) {
            len = (unsigned)g->filter_taps;
        } else {                                                               /* :182-195 */
            float tw;
            if (g->transition_width_hz > 0.0f) tw = g->transition_width_hz;
            else {
                float ref = (r->type == IQ_FILTER_LOWPASS || r->type == IQ_FILTER_HIGHPASS) ? r->freq1_hz : r->freq2_hz;
                tw = fabsf(ref) * TRANSITION_FACTOR;
            }
            if (tw < 1.0f) tw = 1.0f;
            float ntw = tw / (float)fs;
            len = estimate_req_filter_len(ntw, as);
            if (len % 2 == 0) len++;
            if (len < MIN_TAPS) len = MIN_TAPS;
        }
        cf *cur = (cf *)malloc(len * sizeof(cf));
        float *rt = (float *)malloc(len * sizeof(float));
        int stage_complex = (r->type == IQ_FILTER_PASSBAND && fabsf(r->freq1_hz) > 1e-9f);
        if (stage_complex) {                                                   /* :205-218 */
            is_complex = 1;
            float hbw = (r->freq2_hz / 2.0f) / (float)fs;
            liquid_firdes_kaiser(len, hbw, as, 0.0f, rt);
            float fcn = r->freq1_hz / (float)fs;
            nco_crcf sh = nco_crcf_create(LIQUID_NCO);
            nco_crcf_set_frequency(sh, 2.0f * M_PI * fcn);
            for (unsigned k = 0; k < len; k++) {
                cf e; nco_crcf_cexpf(sh, &e);
                cur[k] = (crealf(e) * rt[k]) + (cimagf(e) * rt[k]) * _Complex_I;
                nco_crcf_step(sh);
            }
            nco_crcf_destroy(sh);
        } else {                                                               /* :219-247 */
            float fc, bw;
            switch (r->type) {
                case IQ_FILTER_LOWPASS:
                    fc = r->freq1_hz / (float)fs; liquid_firdes_kaiser(len, fc, as, 0.0f, rt); break;
                case IQ_FILTER_HIGHPASS:
                    fc = r->freq1_hz / (float)fs; liquid_firdes_kaiser(len, fc, as, 0.0f, rt); spectral_invert(rt, len); break;
                case IQ_FILTER_PASSBAND:
                    bw = r->freq2_hz / (float)fs; liquid_firdes_kaiser(len, bw / 2.0f, as, 0.0f, rt); break;
                case IQ_FILTER_STOPBAND: /* centre ignored: notch always at DC (:237-241) */
                    bw = r->freq2_hz / (float)fs; liquid_firdes_kaiser(len, bw / 2.0f, as, 0.0f, rt); spectral_invert(rt, len); break;
                default: memset(rt, 0, len * sizeof(float)); break;
            }
            for (unsigned k = 0; k < len; k++) cur[k] = rt[k];
        }
        free(rt);
        /* master = master (*) cur   (:114-136) */
        int nlen = mlen + (int)len - 1;
        cf *nm = (cf *)calloc((size_t)nlen, sizeof(cf));
        for (int a = 0; a < nlen; a++) {
            int j0 = (a >= mlen) ? (a - mlen + 1) : 0;
            int j1 = (a < (int)len - 1) ? a : ((int)len - 1);
            for (int j = j0; j <= j1; j++) nm[a] += cmul(master[a - j], cur[j]);
        }
        free(master); free(cur);
        master = nm; mlen = nlen;
    }

    /* normalise (:272-299) */
    if (by_peak || is_complex) {
        float peak = 0.0f;
        firfilt_cccf tmp = firfilt_cccf_create(master, (unsigned)mlen);
        if (tmp) {
            for (int i = 0; i < RESPONSE_POINTS; i++) {
                cf H;
                float f = ((float)i / (float)RESPONSE_POINTS) - 0.5f;
                firfilt_cccf_freqresponse(tmp, f, &H);
                float mag = cabsf(H);
                if (mag > peak) peak = mag;
            }
            firfilt_cccf_destroy(tmp);
        }
        if (peak > GAIN_ZERO_THRESH)
            for (int i = 0; i < mlen; i++) master[i] = (crealf(master[i]) / peak) + (cimagf(master[i]) / peak) * _Complex_I;
    } else {
        double dc = 0.0;
        for (int i = 0; i < mlen; i++) dc += crealf(master[i]);
        if (fabs(dc) > GAIN_ZERO_THRESH) {
            float d = (float)dc;
            for (int i = 0; i < mlen; i++) master[i] = (crealf(master[i]) / d) + (cimagf(master[i]) / d) * _Complex_I;
        }
    }

    /* implementation choice (:301-312) and FFT block size (:314-336) */
    int want_fft;
    if (g->filter_type_request != IQ_FILTER_REQ_AUTO) want_fft = (g->filter_type_request == IQ_FILTER_REQ_FFT);
    else want_fft = is_complex;
    c->filt_len = (unsigned)mlen;
    c->filt_taps = master;
    if (want_fft) {
        unsigned block;
        if (g->filter_fft_size > 0) {
            block = (unsigned)g->filter_fft_size / 2;
            if (block < (unsigned)mlen - 1) return -1;
        } else {
            block = 1;
            while (block < (unsigned)mlen - 1) block *= 2;
            if (block < (unsigned)mlen * 2) block *= 2;
        }
        c->filt_block = block;
        if (is_complex) {
            c->filt_obj = fftfilt_cccf_create(master, (unsigned)mlen, block);
            c->filt_impl = IQ_FILTER_IMPL_FFT_ASYM;
        } else {
            float *rt = (float *)malloc((size_t)mlen * sizeof(float));
            for (int i = 0; i < mlen; i++) rt[i] = crealf(master[i]);
            c->filt_obj = fftfilt_crcf_create(rt, (unsigned)mlen, block);
            free(rt);
            c->filt_impl = IQ_FILTER_IMPL_FFT_SYM;
        }
        c->rem = (cf *)calloc(block, sizeof(cf));
        c->rem_len = 0;
    } else {
        if (is_complex) {
            c->filt_obj = firfilt_cccf_create(master, (unsigned)mlen);
            c->filt_impl = IQ_FILTER_IMPL_FIR_ASYM;
        } else {
            float *rt = (float *)malloc((size_t)mlen * sizeof(float));
            for (int i = 0; i < mlen; i++) rt[i] = crealf(master[i]);
            c->filt_obj = firfilt_crcf_create(rt, (unsigned)mlen);
            free(rt);
            c->filt_impl = IQ_FILTER_IMPL_FIR_SYM;
        }
    }
    return c->filt_obj ? 0 : -1;
}

static void filter_free(chain_t *c)
{
    if (c->filt_obj) {
        switch (c->filt_impl) {
            case IQ_FILTER_IMPL_FIR_SYM:  firfilt_crcf_destroy((firfilt_crcf)c->filt_obj); break;
            case IQ_FILTER_IMPL_FIR_ASYM: firfilt_cccf_destroy((firfilt_cccf)c->filt_obj); break;
            case IQ_FILTER_IMPL_FFT_SYM:  fftfilt_crcf_destroy((fftfilt_crcf)c->filt_obj); break;
            case IQ_FILTER_IMPL_FFT_ASYM: fftfilt_cccf_destroy((fftfilt_cccf)c->filt_obj); break;
            default: break;
        }
    }
    free(c->filt_taps); free(c->rem);
}

/* history reset; NOTE the FFT remainder length is NOT cleared by the reference (:417-436) */
static void filter_reset(chain_t *c)
{
    if (!c->filt_obj) return;
    switch (c->filt_impl) {
        case IQ_FILTER_IMPL_FIR_SYM:  firfilt_crcf_reset((firfilt_crcf)c->filt_obj); break;
        case IQ_FILTER_IMPL_FIR_ASYM: firfilt_cccf_reset((firfilt_cccf)c->filt_obj); break;
        case IQ_FILTER_IMPL_FFT_SYM:  fftfilt_crcf_reset((fftfilt_crcf)c->filt_obj); break;
        case IQ_FILTER_IMPL_FFT_ASYM: fftfilt_cccf_reset((fftfilt_cccf)c->filt_obj); break;
        default: break;
    }
}

/* F2/F3 — apply on one chunk.  FIR: in place in `in`.  FFT: remainder||in -> whole blocks of
 * `filt_block` into `out`, leftover kept (:438-526).  Returns frames produced; *where tells
 * the caller which buffer now holds the data. */
static unsigned filter_apply(chain_t *c, cf *in, unsigned n, cf *out, cf **where)
{
    *where = in;
    if (!c->filt_obj) return n;
    if (n == 0) return 0;
    switch (c->filt_impl) {
        case IQ_FILTER_IMPL_FIR_SYM:  firfilt_crcf_execute_block((firfilt_crcf)c->filt_obj, in, n, in); return n;
        case IQ_FILTER_IMPL_FIR_ASYM: firfilt_cccf_execute_block((firfilt_cccf)c->filt_obj, in, n, in); return n;
        default: break;
    }
    unsigned total = c->rem_len + n, done = 0, produced = 0;
    memcpy(out, c->rem, c->rem_len * sizeof(cf));
    memcpy(out + c->rem_len, in, n * sizeof(cf));
    while (total - done >= c->filt_block) {
        if (c->filt_impl == IQ_FILTER_IMPL_FFT_SYM) fftfilt_crcf_execute((fftfilt_crcf)c->filt_obj, out + done, out + produced);
        else fftfilt_cccf_execute((fftfilt_cccf)c->filt_obj, out + done, out + produced);
        done += c->filt_block; produced += c->filt_block;
    }
    c->rem_len = total - done;
    memmove(c->rem, out + done, c->rem_len * sizeof(cf));
    *where = out;
    return produced;
}

/* ---------------------------------------------------------------------------------------
 * Q1 — I/Q correction apply                        reference src/iq_correct.c:141,307-313
 * ------------------------------------------------------------------------------------- */
static void iq_apply(cf *x, int n, float mag, float phase)
{
    const float magp1 = 1.0f + mag;
    for (int i = 0; i < n; i++) {
        float re = crealf(x[i]), im = cimagf(x[i]);
        x[i] = (re * magp1) + (im + phase * re) * _Complex_I;
    }
}

/* Q2 — spectrum / metric / power estimate / hill climb   src/iq_correct.c:154-235,315-393 */
static void iq_spectrum(chain_t *c, const cf *blk, float mag, float phase)
{
    memcpy(c->iq_fft_buf, blk, IQ_NFFT * sizeof(cf));
    iq_apply(c->iq_fft_buf, IQ_NFFT, mag, phase);
    for (int i = 0; i < IQ_NFFT; i++)
        c->iq_fft_buf[i] = (crealf(c->iq_fft_buf[i]) * c->iq_window[i]) + (cimagf(c->iq_fft_buf[i]) * c->iq_window[i]) * _Complex_I;
    fft_execute(c->iq_plan);
    for (int i = 0; i < IQ_NFFT; i++) { /* fftshift + dB (:328-335) */
        cf v = c->iq_fft_buf[(i + IQ_NFFT / 2) % IQ_NFFT];
        float m = cabsf(v);
        m /= (float)IQ_NFFT;
        c->iq_spectrum[i] = 20.0f * log10f(m + 1e-12f);
    }
}
static float iq_metric(chain_t *c, const cf *blk, float mag, float phase)
{
    const int half = IQ_NFFT / 2;
    iq_spectrum(c, blk, mag, phase);
    float util = 0.0f;
    const int lo = (int)(0.05f * half), hi = (int)(0.95f * half);
    for (int i = lo; i < hi; i++) {
        float pn = c->iq_spectrum[i], pp = c->iq_spectrum[IQ_NFFT - 1 - i];
        if (pp > -80.0f || pn > -80.0f) { float d = pp - pn; util += d * d; }
    }
    return util;
}
static void iq_estimate_power(chain_t *c, const cf *blk)
{
    const int half = IQ_NFFT / 2;
    iq_spectrum(c, blk, 0.0f, 0.0f);
    float maxp = -1000.0f; double sum = 0.0; int cnt = 0;
    const int lo = (int)(0.05f * half), hi = (int)(0.95f * half);
    for (int i = lo; i < hi; i++) {
        float pn = c->iq_spectrum[i], pp = c->iq_spectrum[IQ_NFFT - 1 - i];
        if (pp > maxp) maxp = pp;
        if (pn > maxp) maxp = pn;
        sum += pp + pn; cnt += 2;
    }
    if (cnt > 0) { c->iq_avg_power = (float)(sum / cnt); c->iq_power_range = maxp - c->iq_avg_power; }
    else { c->iq_avg_power = 0.0f; c->iq_power_range = 0.0f; }
}
static float rand_dir(void) { return (rand() > (RAND_MAX / 2)) ? 1.0f : -1.0f; } /* :391 */

int iqo_iq_optimize(void *hv, const float *block1024, unsigned int seed, float *mag, float *phase,
                    float *avg_power, float *power_range)
{
    chain_t *c = (chain_t *)hv;
    if (!c->cfg.iq_correction_enable) return -1;
    const cf *blk = (const cf *)block1024;
    srand(seed);
    iq_estimate_power(c, blk);
    *avg_power = c->iq_avg_power; *power_range = c->iq_power_range;
    if (c->iq_power_range >= IQ_POWER_THRESH) {
        float g = c->iq_mag, p = c->iq_phase;
        float best = iq_metric(c, blk, g, p);
        for (int i = 0; i < IQ_PASSES; i++) {
            float cg = g + IQ_INCREMENT * rand_dir();
            float cp = p + IQ_INCREMENT * rand_dir();
            float m = iq_metric(c, blk, cg, cp);
            if (m > best) { best = m; g = cg; p = cp; } /* keeps a candidate when the metric INCREASES (:196) */
        }
        c->iq_mag = ((1.0f - IQ_SMOOTHING) * c->iq_mag) + (IQ_SMOOTHING * g);
        c->iq_phase = ((1.0f - IQ_SMOOTHING) * c->iq_phase) + (IQ_SMOOTHING * p);
    }
    *mag = c->iq_mag; *phase = c->iq_phase;
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * G1/G2 — output AGC                                         reference src/agc.c:21-245
 * ------------------------------------------------------------------------------------- */
static int agc_setup(chain_t *c)
{
    const iq_chain_cfg *g = &c->cfg;
    c->agc_rms = NULL;
    if (!g->agc_enable) return 0;
    c->agc_locked = 0; c->agc_gain = 1.0f; c->agc_seen = 0;
    c->agc_last_strong = now_sec();
    if (g->agc_profile != IQ_AGC_DIGITAL) {                                    /* :38-68 */
        agc_crcf q = agc_crcf_create();
        float bw = (g->agc_profile == IQ_AGC_DX) ? AGC_DX_BW : AGC_LOCAL_BW;
        float target = (g->agc_target_level_arg > 0) ? g->agc_target_level_arg : 0.5f;
        agc_crcf_set_bandwidth(q, bw);
        agc_crcf_set_signal_level(q, target);
        agc_crcf_set_gain(q, 1.0f);
        c->agc_rms = q;
        c->agc_peak_mem = 0.001f;
    } else {
        c->agc_peak_mem = 0.05f;                                               /* :78 */
    }
    return 0;
}
static void agc_reset_state(chain_t *c)                                        /* :225-238 */
{
    if (c->agc_rms) { agc_crcf_reset(c->agc_rms); agc_crcf_set_gain(c->agc_rms, 1.0f); }
    c->agc_locked = 0; c->agc_seen = 0; c->agc_peak_mem = 0.05f; c->agc_gain = 1.0f;
    c->agc_last_strong = now_sec();
}
static float block_peak(const cf *x, unsigned n)
{
    float pk = 0.0f;
    for (unsigned i = 0; i < n; i++) { float m = cabsf(x[i]); if (m > pk) pk = m; }
    return pk;
}
static void scale_block(cf *x, unsigned n, float g)
{
    for (unsigned i = 0; i < n; i++) x[i] = (crealf(x[i]) * g) + (cimagf(x[i]) * g) * _Complex_I;
}
static void agc_apply(chain_t *c, cf *x, unsigned n)
{
    const iq_chain_cfg *g = &c->cfg;
    if (!g->agc_enable || n == 0) return;
    if (c->agc_rms) { agc_crcf_execute_block(c->agc_rms, x, n, x); return; }    /* :92-100 */
    if (g->agc_profile != IQ_AGC_DIGITAL) return;
    float target = (g->agc_target_level_arg > 0) ? g->agc_target_level_arg : AGC_DIG_TARGET;
    if (!c->agc_locked) {                                                      /* :117-160 */
        float pk = block_peak(x, n);
        if (pk > c->agc_peak_mem) c->agc_peak_mem = pk;
        float safe = (c->agc_peak_mem < 1e-4f) ? 1e-4f : c->agc_peak_mem;
        float gain = target / safe;
        scale_block(x, n, gain);
        double elapsed = (double)c->agc_seen / c->target_rate;
        if (elapsed > AGC_DIG_LOCK_S) {
            c->agc_locked = 1; c->agc_gain = gain; c->agc_last_strong = now_sec();
        }
    } else {                                                                   /* :165-218 */
        float gain = c->agc_gain;
        float pk = block_peak(x, n);
        float opk = pk * gain;
        double t = now_sec();
        if (opk > 1.0f) {
            gain = 0.99f / pk;
            c->agc_last_strong = t;
        } else if (opk > (target * AGC_DIG_LOWER)) {
            c->agc_last_strong = t;
        } else if (t - c->agc_last_strong > AGC_DIG_HANG_S) {
            gain *= AGC_DIG_RECOVER;
        }
        c->agc_gain = gain;
        scale_block(x, n, gain);
    }
    c->agc_seen += n;                                                          /* :220 */
}

/* ---------------------------------------------------------------------------------------
 * create / destroy
 * ------------------------------------------------------------------------------------- */
void iqo_destroy(void *hv)
{
    chain_t *c = (chain_t *)hv;
    if (!c) return;
    if (c->agc_rms) agc_crcf_destroy(c->agc_rms);
    filter_free(c);
    if (c->rs) msresamp_crcf_destroy(c->rs);
    if (c->nco_pre) nco_crcf_destroy(c->nco_pre);
    if (c->nco_post) nco_crcf_destroy(c->nco_post);
    if (c->iq_plan) fft_destroy_plan(c->iq_plan);
    if (c->dc) iirfilt_crcf_destroy(c->dc);
    free(c->buf_a); free(c->buf_b); free(c->out_tmp);
    free(c);
}

void *iqo_create(const iq_chain_cfg *cfg)
{
    chain_t *c = (chain_t *)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->cfg = *cfg;
    c->in_rate = (int)cfg->input_rate_hz;
    c->target_rate = cfg->target_rate_hz;
    c->in_bps = iqo_get_bytes_per_sample(cfg->input_format);
    c->out_bps = iqo_get_bytes_per_sample(cfg->output_format);
    if (!c->in_bps || !c->out_bps) goto fail;

    /* ratio: float r = (float)(target / input)            reference src/setup.c:94-113 */
    if (cfg->no_resample) { c->target_rate = (double)c->in_rate; c->passthrough = 1; }
    c->ratio = (float)(c->target_rate / (double)c->in_rate);
    if (!isfinite(c->ratio) || c->ratio < MIN_RATIO || c->ratio > MAX_RATIO) goto fail;

    /* D1: alpha = (float)(2 pi 10 / Fs)                    reference src/dc_block.c:32,54 */
    if (cfg->dc_block_enable) {
        float alpha = (float)(2.0 * M_PI * DC_CUTOFF_HZ / c->in_rate);
        if (alpha <= 0.0f) goto fail;
        c->dc = iirfilt_crcf_create_dc_blocker(alpha);
        if (!c->dc) goto fail;
    }
    /* Q1/Q2 init: Hamming window, FFT plan                 reference src/iq_correct.c:86-139 */
    if (cfg->iq_correction_enable) {
        c->iq_plan = fft_create_plan(IQ_NFFT, c->iq_fft_buf, c->iq_fft_buf, LIQUID_FFT_FORWARD, 0);
        for (unsigned i = 0; i < IQ_NFFT; i++)
            c->iq_window[i] = 0.54f - 0.46f * cosf(2.0f * (float)M_PI * (float)i / (float)(IQ_NFFT - 1));
        c->iq_mag = cfg->iq_mag; c->iq_phase = cfg->iq_phase;
    }
    /* N1: one NCO, pre at Fs_in or post at target rate     reference src/frequency_shift.c:24-84 */
    /* AppResources.nco_shift_hz is a double: a float CLI argument arrives widened (frequency_shift.c:32-33),
     * a WAV centre-target shift does not fit a float at all (input_wav.c:614-628) */
    c->shift_hz = cfg->freq_shift_hz;
    if (cfg->shift_after_resample && fabs(c->shift_hz) < 1e-9) goto fail;
    if (fabs(c->shift_hz) >= 1e-9) {
        double rate = cfg->shift_after_resample ? c->target_rate : (double)c->in_rate;
        if (fabs(c->shift_hz) > SHIFT_LIMIT * rate) goto fail;
        nco_crcf q = nco_crcf_create(LIQUID_NCO);
        float w = (float)(2.0 * M_PI * fabs(c->shift_hz) / rate);
        nco_crcf_set_frequency(q, w);
        if (cfg->shift_after_resample) c->nco_post = q; else c->nco_pre = q;
    }
    /* R1                                                    reference src/resampler.c:20-34 */
    if (!c->passthrough) {
        c->rs = msresamp_crcf_create(c->ratio, RESAMP_AS_DB);
        if (!c->rs) goto fail;
    }
    if (filter_design(c) != 0) goto fail;
    if (agc_setup(c) != 0) goto fail;

    /* cf32 capacity per chunk                               reference src/pipeline.c:232-265 */
    size_t max_pre = CHUNK;
    int is_fft = (c->filt_impl == IQ_FILTER_IMPL_FFT_SYM || c->filt_impl == IQ_FILTER_IMPL_FFT_ASYM);
    if (c->filt_obj && !c->filt_post && is_fft && c->filt_block > max_pre) max_pre = c->filt_block;
    size_t rs_cap = (size_t)ceil((double)max_pre * fmax(1.0, (double)c->ratio)) + OUT_MARGIN;
    size_t cap = max_pre > rs_cap ? max_pre : rs_cap;
    if (c->filt_obj && c->filt_post && is_fft && c->filt_block > cap) cap = c->filt_block;
    if (is_fft) cap += c->filt_block; /* head room the reference lacks */
    c->cap = cap;
    c->buf_a = (cf *)calloc(cap, sizeof(cf));
    c->buf_b = (cf *)calloc(cap, sizeof(cf));
    c->out_tmp = (unsigned char *)calloc(cap, c->out_bps);
    if (!c->buf_a || !c->buf_b || !c->out_tmp) goto fail;
    return c;
fail:
    iqo_destroy(c);
    return NULL;
}

/* stream discontinuity: pre_processor_reset + resampler_reset + post_processor_reset
 * (reference src/pre_processor.c:57-61, src/resampler.c:43-47, src/post_processor.c:72-76) */
void iqo_reset(void *hv)
{
    chain_t *c = (chain_t *)hv;
    if (c->dc) iirfilt_crcf_reset(c->dc);
    if (c->nco_pre) nco_crcf_set_phase(c->nco_pre, 0.0f);
    filter_reset(c);
    if (c->rs) msresamp_crcf_reset(c->rs);
    if (c->nco_post) nco_crcf_set_phase(c->nco_post, 0.0f);
    filter_reset(c);
    if (c->cfg.agc_enable || 1) agc_reset_state(c);
}

void iqo_set_capture(void *hv, int stage, float *buf, int64_t cap_samples)
{
    chain_t *c = (chain_t *)hv;
    if (stage < 0 || stage > 2) return;
    c->cap_buf[stage] = (cf *)buf; c->cap_cap[stage] = cap_samples; c->cap_len[stage] = 0;
}
int64_t iqo_get_capture_len(void *hv, int stage) { return ((chain_t *)hv)->cap_len[stage]; }
void iqo_set_trace(void *hv, uint32_t *buf, int64_t cap) { chain_t *c = hv; c->trace = buf; c->trace_cap = cap; c->trace_len = 0; }
int64_t iqo_get_trace_len(void *hv) { return ((chain_t *)hv)->trace_len; }
static void capture(chain_t *c, int stage, const cf *p, size_t n)
{
    if (!c->cap_buf[stage]) return;
    int64_t room = c->cap_cap[stage] - c->cap_len[stage];
    if ((int64_t)n > room) n = room > 0 ? (size_t)room : 0;
    memcpy(c->cap_buf[stage] + c->cap_len[stage], p, n * sizeof(cf));
    c->cap_len[stage] += (int64_t)n;
}

static void mix(nco_crcf q, double shift_hz, cf *in, cf *out, unsigned n)      /* frequency_shift.c:86-96 */
{
    if (!q || n == 0) return;
    if (shift_hz >= 0) nco_crcf_mix_block_up(q, in, out, n);
    else nco_crcf_mix_block_down(q, in, out, n);
}

/* ---------------------------------------------------------------------------------------
 * P1 + resampler stage + P2, one 16384-frame chunk at a time
 *   reference src/pre_processor.c:10-55, src/pipeline.c:512-528, src/post_processor.c:9-70
 * ------------------------------------------------------------------------------------- */
int iqo_process(void *hv, const void *raw_in, int64_t n_frames, void *out, int64_t out_cap_bytes,
                int64_t *out_frames)
{
    chain_t *c = (chain_t *)hv;
    const iq_chain_cfg *g = &c->cfg;
    const char *src = (const char *)raw_in;
    char *dst = (char *)out;
    int64_t done = 0, written = 0;
    while (done < n_frames) {
        int64_t n = n_frames - done;
        if (n > CHUNK) n = CHUNK;
        cf *a = c->buf_a, *b = c->buf_b;
        unsigned frames = (unsigned)n, nout = 0;

        /* ---- pre-processor: everything in place in buffer A ---- */
        if (iqo_convert_block_to_cf32(src + done * (int64_t)c->in_bps, (float *)a, (size_t)n, g->input_format, g->gain) != 0) return -1;
        if (c->dc) iirfilt_crcf_execute_block(c->dc, a, frames, a);
        if (g->iq_correction_enable) iq_apply(a, (int)frames, c->iq_mag, c->iq_phase);
        mix(c->nco_pre, c->shift_hz, a, a, frames);
        if (c->filt_obj && !c->filt_post) {
            cf *where;
            frames = filter_apply(c, a, frames, a, &where); /* reference aliases in/out here */
        }
        if (frames > 0) {
            capture(c, 0, a, frames);
            /* ---- resampler: A -> B ---- */
            if (c->passthrough) { nout = frames; memcpy(b, a, nout * sizeof(cf)); }
            else msresamp_crcf_execute(c->rs, a, frames, b, &nout);
            capture(c, 1, b, nout);
            /* ---- post-processor ---- */
            if (nout > 0) {
                cf *cur = b, *other = a;
                if (c->filt_obj && c->filt_post) {
                    cf *where;
                    nout = filter_apply(c, cur, nout, other, &where);
                    if (where != cur) { other = cur; cur = where; }
                }
                if (c->nco_post) {
                    mix(c->nco_post, c->shift_hz, cur, other, nout);
                    cf *t = cur; cur = other; other = t;
                }
                agc_apply(c, cur, nout);
                capture(c, 2, cur, nout);
                if (iqo_convert_cf32_to_block((const float *)cur, c->out_tmp, nout, g->output_format) != 0) return -1;
            }
        }
        if (c->trace && c->trace_len < c->trace_cap) c->trace[c->trace_len++] = nout;
        if (nout > 0) {
            size_t nb = (size_t)nout * c->out_bps;
            if (written * (int64_t)c->out_bps + (int64_t)nb > out_cap_bytes) return -2;
            memcpy(dst + written * (int64_t)c->out_bps, c->out_tmp, nb);
            written += nout;
        }
        done += n;
    }
    *out_frames = written;
    return 0;
}

void iqo_get_info(void *hv, iqo_info *o)
{
    chain_t *c = (chain_t *)hv;
    memset(o, 0, sizeof(*o));
    o->ratio = c->ratio;
    o->filter_impl = c->filt_impl;
    o->filter_post_resample = c->filt_post;
    o->filter_block_size = c->filt_block;
    o->filter_num_taps = c->filt_len;
    nco_crcf q = c->nco_pre ? c->nco_pre : c->nco_post;
    if (q) o->nco_dtheta = liquid_compat_nco_get_dtheta(q);
    o->nco_is_post = c->nco_post != NULL;
    o->cap_samples = (uint32_t)c->cap;
    o->agc_locked = (uint32_t)c->agc_locked;
    o->agc_gain = c->agc_gain;
    o->agc_peak_memory = c->agc_peak_mem;
    o->agc_samples_seen = c->agc_seen;
}
uint32_t iqo_get_filter_taps(void *hv, float *out, uint32_t cap)
{
    chain_t *c = (chain_t *)hv;
    cf *o = (cf *)out;
    for (uint32_t i = 0; i < c->filt_len && i < cap; i++) o[i] = c->filt_taps[i];
    return c->filt_len;
}
void *iqo_get_msresamp(void *hv) { return ((chain_t *)hv)->rs; }
