/*
 * liquid_compat.c — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product.
 *
 * Restatement of the liquid-dsp primitives iq_tool's chain calls (see liquid/liquid.h in
 * this directory for the reference call sites).  liquid-dsp is NOT vendored in the
 * reference and NOT installed here; every object below restates the upstream algorithm
 * (liquid-dsp 1.3.2..1.6: src/nco/src/nco.proto.c, src/filter/src/{iirfilt,firfilt,
 * fftfilt,firdes,msresamp,msresamp2,resamp2,resamp.fixed,firpfb}.proto.c,
 * src/agc/src/agc.proto.c, src/math/src/{math.bessel,math.gamma,windows}.c) in plain
 * scalar C with strictly sequential float accumulation.  PARITY UNPINNED against a real
 * libliquid: no liquid binary/source exists in this environment to diff against.
 *
 * Build with -ffp-contract=off so results do not depend on FMA availability.
 */
#include "liquid/liquid.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef liquid_float_complex cf;

static inline cf cf_make(float re, float im) { return re + im * _Complex_I; }
/* complex * complex with separately rounded products (what __mulsc3 computes for finite inputs) */
static inline cf cf_mul(cf a, cf b)
{
    float ar = crealf(a), ai = cimagf(a), br = crealf(b), bi = cimagf(b);
    return cf_make(ar * br - ai * bi, ar * bi + ai * br);
}
static inline cf cf_scale(cf a, float s) { return cf_make(crealf(a) * s, cimagf(a) * s); }

/* ======================================================================================
 * math / design helpers            (liquid: src/math/src/math.gamma.c, math.bessel.c,
 *                                   src/math/src/windows.c, src/filter/src/firdes.c)
 * ==================================================================================== */

/* log(Gamma(z)); recursion below 10, Stirling-like high-value approximation above */
float liquid_lngammaf(float z)
{
    float g;
    if (z < 0) {
        fprintf(stderr, "liquid_compat: lngammaf undefined for z < 0\n");
        return 0.0f;
    } else if (z < 10.0f) {
        /* gamma(z+1) = z*gamma(z)  =>  lngamma(z) = lngamma(z+1) - ln(z) */
        return liquid_lngammaf(z + 1.0f) - logf(z);
    } else {
        g = 0.5 * (logf(2 * M_PI) - log(z));
        g += z * (logf(z + (1 / (12.0f * z - 0.1f / z))) - 1);
    }
    return g;
}

/* I0(z): 32-term series evaluated in the log domain */
float liquid_besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    unsigned int k;
    float t, y = 0.0f;
    for (k = 0; k < 32; k++) {
        t = k * logf(0.5f * z) - liquid_lngammaf((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}

float kaiser_beta_As(float as)
{
    as = fabsf(as);
    float beta;
    if (as > 50.0f)
        beta = 0.1102f * (as - 8.7f);
    else if (as > 21.0f)
        beta = 0.5842 * powf(as - 21, 0.4f) + 0.07886f * (as - 21);
    else
        beta = 0.0f;
    return beta;
}

float liquid_kaiser(unsigned int i, unsigned int wlen, float beta)
{
    float t = (float)i - (float)(wlen - 1) / 2;
    float r = 2.0f * t / (float)(wlen - 1);
    float a = liquid_besseli0f(beta * sqrtf(1 - r * r));
    float b = liquid_besseli0f(beta);
    return a / b;
}

float sincf(float x)
{
    /* product expansion near zero: sinc(z) = prod_k cos(pi z / 2^k) */
    if (fabsf(x) < 0.01f)
        return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
    return sinf(M_PI * x) / (M_PI * x);
}

/* Kaiser's length estimate (liquid: estimate_req_filter_len -> _Kaiser) */
unsigned int estimate_req_filter_len(float df, float as)
{
    if (df > 0.5f || df <= 0.0f) {
        fprintf(stderr, "liquid_compat: estimate_req_filter_len(), invalid bandwidth %g\n", df);
        return 0;
    }
    if (as <= 0.0f) {
        fprintf(stderr, "liquid_compat: estimate_req_filter_len(), invalid stopband %g\n", as);
        return 0;
    }
    float len = (as - 7.95f) / (14.26f * df);
    return (unsigned int)len;
}

/* Kaiser-windowed sinc, NOT gain-normalised (DC gain ~ 1/(2 fc)) */
int liquid_firdes_kaiser(unsigned int n, float fc, float as, float mu, float *h)
{
    if (mu < -0.5f || mu > 0.5f || fc <= 0.0f || fc > 0.5f || n == 0) {
        fprintf(stderr, "liquid_compat: liquid_firdes_kaiser(), invalid config (n=%u fc=%g mu=%g)\n", n, fc, mu);
        return -1;
    }
    float beta = kaiser_beta_As(as);
    unsigned int i;
    for (i = 0; i < n; i++) {
        float t  = (float)i - (float)(n - 1) / 2 + mu;
        float h1 = sincf(2.0f * fc * t);
        float h2 = liquid_kaiser(i, n, beta);
        h[i] = h1 * h2;
    }
    return LIQUID_OK;
}

/* ======================================================================================
 * nco_crcf, LIQUID_NCO                               (liquid: src/nco/src/nco.proto.c)
 *   32-bit phase accumulator, 1024-entry sine table, nearest-entry rounding
 * ==================================================================================== */
struct nco_crcf_s {
    liquid_ncotype type;
    uint32_t theta, d_theta;
    float sintab[1024];
};

uint32_t liquid_compat_nco_constrain(float theta)
{
    float p = theta * 0.159154943091895; /* 1/(2 pi); product in double, stored to float */
    float fpart = p - ((long)p);         /* in (-1,1) */
    if (fpart < 0.) fpart += 1.;
    /* fpart * 0xffffffff: the unsigned constant converts to float 4294967296.0f.
     * x86-64 converts float->uint32 through a 64-bit integer and keeps the low word. */
    float scaled = fpart * (float)0xffffffffu;
    return (uint32_t)(int64_t)scaled;
}

nco_crcf nco_crcf_create(liquid_ncotype type)
{
    nco_crcf q = (nco_crcf)malloc(sizeof(struct nco_crcf_s));
    if (!q) return NULL;
    q->type = type;
    unsigned int i;
    for (i = 0; i < 1024; i++)
        q->sintab[i] = sinf(2.0f * M_PI * (float)i / (float)1024);
    q->theta = 0;
    q->d_theta = 0;
    return q;
}
int nco_crcf_destroy(nco_crcf q) { free(q); return LIQUID_OK; }
int nco_crcf_reset(nco_crcf q) { q->theta = 0; q->d_theta = 0; return LIQUID_OK; }
int nco_crcf_set_frequency(nco_crcf q, float dtheta) { q->d_theta = liquid_compat_nco_constrain(dtheta); return LIQUID_OK; }
int nco_crcf_set_phase(nco_crcf q, float phi) { q->theta = liquid_compat_nco_constrain(phi); return LIQUID_OK; }
int nco_crcf_step(nco_crcf q) { q->theta += q->d_theta; return LIQUID_OK; }
uint32_t liquid_compat_nco_get_theta(nco_crcf q) { return q->theta; }
uint32_t liquid_compat_nco_get_dtheta(nco_crcf q) { return q->d_theta; }

static inline void nco_sincos(nco_crcf q, float *s, float *c)
{
    if (q->type == LIQUID_NCO) {
        unsigned int idx = ((q->theta + (1u << 21)) >> 22) & 0x3ff; /* round to nearest entry */
        *s = q->sintab[idx];
        *c = q->sintab[(idx + 256) & 0x3ff];
    } else {
        float th = (float)q->theta * (float)(2.0 * M_PI / 4294967296.0);
        *s = sinf(th);
        *c = cosf(th);
    }
}
int nco_crcf_cexpf(nco_crcf q, cf *y)
{
    float s, c;
    nco_sincos(q, &s, &c);
    *y = cf_make(c, s);
    return LIQUID_OK;
}
int nco_crcf_mix_block_up(nco_crcf q, cf *x, cf *y, unsigned int n)
{
    unsigned int i;
    for (i = 0; i < n; i++) {
        float s, c;
        nco_sincos(q, &s, &c);
        y[i] = cf_mul(x[i], cf_make(c, s));
        q->theta += q->d_theta;
    }
    return LIQUID_OK;
}
int nco_crcf_mix_block_down(nco_crcf q, cf *x, cf *y, unsigned int n)
{
    unsigned int i;
    for (i = 0; i < n; i++) {
        float s, c;
        nco_sincos(q, &s, &c);
        y[i] = cf_mul(x[i], cf_make(c, -s));
        q->theta += q->d_theta;
    }
    return LIQUID_OK;
}

/* ======================================================================================
 * iirfilt_crcf DC blocker                      (liquid: src/filter/src/iirfilt.proto.c)
 *   b = {1,-1}, a = {1, -(1-alpha)}; direct-form II:  v0 = x - a1*v1 ; y = v0 - v1
 * ==================================================================================== */
struct iirfilt_crcf_s {
    float b[2], a[2];
    cf v[2];
};
iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha)
{
    if (alpha <= 0.0f) return NULL;
    iirfilt_crcf q = (iirfilt_crcf)malloc(sizeof(struct iirfilt_crcf_s));
    if (!q) return NULL;
    float a1 = -1.0f + alpha;
    q->b[0] = 1.0f; q->b[1] = -1.0f;
    q->a[0] = 1.0f; q->a[1] = a1;
    q->v[0] = q->v[1] = 0;
    return q;
}
int iirfilt_crcf_destroy(iirfilt_crcf q) { free(q); return LIQUID_OK; }
int iirfilt_crcf_reset(iirfilt_crcf q) { q->v[0] = q->v[1] = 0; return LIQUID_OK; }
int iirfilt_crcf_execute_block(iirfilt_crcf q, cf *x, unsigned int n, cf *y)
{
    unsigned int i;
    for (i = 0; i < n; i++) {
        q->v[1] = q->v[0];
        cf v0 = x[i];
        v0 = cf_make(crealf(v0) - q->a[1] * crealf(q->v[1]), cimagf(v0) - q->a[1] * cimagf(q->v[1]));
        q->v[0] = v0;
        float yr = 0.0f, yi = 0.0f;
        yr += q->b[0] * crealf(q->v[0]); yi += q->b[0] * cimagf(q->v[0]);
        yr += q->b[1] * crealf(q->v[1]); yi += q->b[1] * cimagf(q->v[1]);
        y[i] = cf_make(yr, yi);
    }
    return LIQUID_OK;
}

/* ======================================================================================
 * firfilt_crcf / firfilt_cccf                  (liquid: src/filter/src/firfilt.proto.c)
 *   taps stored reversed; y[n] = sum_k h[k] x[n-k]; summation runs oldest sample first
 * ==================================================================================== */
struct firfilt_crcf_s { unsigned int n; float *hrev; cf *w; unsigned int wi; float scale; };
struct firfilt_cccf_s { unsigned int n; cf *hrev; cf *w; unsigned int wi; cf scale; };

/* delay line: linear buffer of 2n-? kept simple: circular with index of the oldest sample */
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n)
{
    if (n == 0) return NULL;
    firfilt_crcf q = (firfilt_crcf)malloc(sizeof(*q));
    q->n = n;
    q->hrev = (float *)malloc(n * sizeof(float));
    q->w = (cf *)calloc(n, sizeof(cf));
    unsigned int i;
    for (i = n; i > 0; i--) q->hrev[i - 1] = h[n - i];
    q->wi = 0;
    q->scale = 1.0f;
    return q;
}
int firfilt_crcf_destroy(firfilt_crcf q) { free(q->hrev); free(q->w); free(q); return LIQUID_OK; }
int firfilt_crcf_reset(firfilt_crcf q) { memset(q->w, 0, q->n * sizeof(cf)); q->wi = 0; return LIQUID_OK; }
int firfilt_crcf_execute_block(firfilt_crcf q, cf *x, unsigned int nx, cf *y)
{
    unsigned int t, i;
    const unsigned int n = q->n;
    for (t = 0; t < nx; t++) {
        q->w[q->wi] = x[t];              /* overwrite the oldest with the newest */
        q->wi = (q->wi + 1 == n) ? 0 : q->wi + 1; /* wi now indexes the oldest */
        float sr = 0.0f, si = 0.0f;
        unsigned int p = q->wi;
        for (i = 0; i < n; i++) {
            sr += q->hrev[i] * crealf(q->w[p]);
            si += q->hrev[i] * cimagf(q->w[p]);
            p = (p + 1 == n) ? 0 : p + 1;
        }
        y[t] = cf_make(sr * q->scale, si * q->scale);
    }
    return LIQUID_OK;
}

firfilt_cccf firfilt_cccf_create(cf *h, unsigned int n)
{
    if (n == 0) return NULL;
    firfilt_cccf q = (firfilt_cccf)malloc(sizeof(*q));
    q->n = n;
    q->hrev = (cf *)malloc(n * sizeof(cf));
    q->w = (cf *)calloc(n, sizeof(cf));
    unsigned int i;
    for (i = n; i > 0; i--) q->hrev[i - 1] = h[n - i];
    q->wi = 0;
    q->scale = 1.0f;
    return q;
}
int firfilt_cccf_destroy(firfilt_cccf q) { free(q->hrev); free(q->w); free(q); return LIQUID_OK; }
int firfilt_cccf_reset(firfilt_cccf q) { memset(q->w, 0, q->n * sizeof(cf)); q->wi = 0; return LIQUID_OK; }
int firfilt_cccf_execute_block(firfilt_cccf q, cf *x, unsigned int nx, cf *y)
{
    unsigned int t, i;
    const unsigned int n = q->n;
    for (t = 0; t < nx; t++) {
        q->w[q->wi] = x[t];
        q->wi = (q->wi + 1 == n) ? 0 : q->wi + 1;
        cf s = 0;
        unsigned int p = q->wi;
        for (i = 0; i < n; i++) {
            s += cf_mul(q->hrev[i], q->w[p]);
            p = (p + 1 == n) ? 0 : p + 1;
        }
        y[t] = cf_mul(s, q->scale);
    }
    return LIQUID_OK;
}
unsigned int liquid_compat_firfilt_crcf_get_taps(firfilt_crcf q, float *h, unsigned int cap)
{
    unsigned int i;
    for (i = 0; i < q->n && i < cap; i++) h[i] = q->hrev[q->n - 1 - i];
    return q->n;
}
unsigned int liquid_compat_firfilt_cccf_get_taps(firfilt_cccf q, cf *h, unsigned int cap)
{
    unsigned int i;
    for (i = 0; i < q->n && i < cap; i++) h[i] = q->hrev[q->n - 1 - i];
    return q->n;
}
int firfilt_cccf_freqresponse(firfilt_cccf q, float fc, cf *H)
{
    unsigned int i;
    cf acc = 0.0f;
    for (i = 0; i < q->n; i++) {
        /* liquid: cexpf(_Complex_I*2*M_PI*fc*i) -- the angle is formed in double but cexpf()
         * receives it as a FLOAT complex, so cos/sin see the float-rounded angle */
        float ang = (float)(2 * M_PI * fc * i);
        cf e = cf_make(cosf(ang), sinf(ang));
        acc += cf_mul(q->hrev[i], e);
    }
    *H = cf_mul(acc, q->scale);
    return LIQUID_OK;
}

/* ======================================================================================
 * fft                                               (liquid: src/fft/src/fft_*.proto.c)
 *   un-normalised DFT; radix-2 for powers of two, direct DFT otherwise
 * ==================================================================================== */
struct fftplan_s {
    unsigned int n;
    cf *x, *y;
    int dir;
    int pow2;
    cf *tw;            /* n/2 twiddles exp(-+ j 2 pi k / n) */
    unsigned int *rev; /* bit reversal */
};
fftplan fft_create_plan(unsigned int n, cf *x, cf *y, int dir, int flags)
{
    (void)flags;
    if (n == 0) return NULL;
    fftplan p = (fftplan)calloc(1, sizeof(*p));
    p->n = n; p->x = x; p->y = y; p->dir = dir;
    p->pow2 = (n & (n - 1)) == 0;
    double sgn = (dir == LIQUID_FFT_FORWARD) ? -1.0 : 1.0;
    unsigned int k;
    if (p->pow2) {
        p->tw = (cf *)malloc((n / 2 + 1) * sizeof(cf));
        for (k = 0; k < n / 2; k++) {
            double a = sgn * 2.0 * M_PI * (double)k / (double)n;
            p->tw[k] = cf_make((float)cos(a), (float)sin(a));
        }
        p->rev = (unsigned int *)malloc(n * sizeof(unsigned int));
        unsigned int bits = 0;
        while ((1u << bits) < n) bits++;
        for (k = 0; k < n; k++) {
            unsigned int r = 0, b;
            for (b = 0; b < bits; b++)
                if (k & (1u << b)) r |= 1u << (bits - 1 - b);
            p->rev[k] = r;
        }
    } else {
        p->tw = (cf *)malloc(n * sizeof(cf));
        for (k = 0; k < n; k++) {
            double a = sgn * 2.0 * M_PI * (double)k / (double)n;
            p->tw[k] = cf_make((float)cos(a), (float)sin(a));
        }
    }
    return p;
}
int fft_destroy_plan(fftplan p)
{
    if (!p) return LIQUID_OK;
    free(p->tw); free(p->rev); free(p);
    return LIQUID_OK;
}
int fft_execute(fftplan p)
{
    const unsigned int n = p->n;
    unsigned int i, k;
    if (p->pow2) {
        cf *tmp = p->y;
        cf *scratch = NULL;
        if (p->x == p->y) {
            scratch = (cf *)malloc(n * sizeof(cf));
            memcpy(scratch, p->x, n * sizeof(cf));
            for (i = 0; i < n; i++) tmp[p->rev[i]] = scratch[i];
            free(scratch);
        } else {
            for (i = 0; i < n; i++) tmp[p->rev[i]] = p->x[i];
        }
        unsigned int half;
        for (half = 1; half < n; half <<= 1) {
            unsigned int stride = n / (2 * half);
            for (i = 0; i < n; i += 2 * half) {
                for (k = 0; k < half; k++) {
                    cf w = p->tw[k * stride];
                    cf a = tmp[i + k];
                    cf b = cf_mul(tmp[i + k + half], w);
                    tmp[i + k] = a + b;
                    tmp[i + k + half] = a - b;
                }
            }
        }
    } else {
        cf *out = (cf *)malloc(n * sizeof(cf));
        for (k = 0; k < n; k++) {
            double sr = 0.0, si = 0.0;
            for (i = 0; i < n; i++) {
                cf w = p->tw[((unsigned long long)i * k) % n];
                sr += (double)crealf(p->x[i]) * crealf(w) - (double)cimagf(p->x[i]) * cimagf(w);
                si += (double)crealf(p->x[i]) * cimagf(w) + (double)cimagf(p->x[i]) * crealf(w);
            }
            out[k] = cf_make((float)sr, (float)si);
        }
        memcpy(p->y, out, n * sizeof(cf));
        free(out);
    }
    return LIQUID_OK;
}

/* ======================================================================================
 * fftfilt_crcf / fftfilt_cccf                  (liquid: src/filter/src/fftfilt.proto.c)
 *   overlap-add, block n, FFT size 2n, H = FFT(h || 0), scale = 1/(2n), tail w[n]
 * ==================================================================================== */
struct fftfilt_cccf_s {
    unsigned int h_len, n;
    cf *h;
    cf *time_buf, *freq_buf, *H, *w;
    fftplan fft, ifft;
    float scale;
};
struct fftfilt_crcf_s { struct fftfilt_cccf_s core; };

static int fftfilt_core_init(struct fftfilt_cccf_s *q, const cf *h, unsigned int h_len, unsigned int n)
{
    if (h_len == 0 || n < h_len - 1) {
        fprintf(stderr, "liquid_compat: fftfilt_create(), block length must be at least h_len-1\n");
        return -1;
    }
    q->h_len = h_len; q->n = n;
    q->h = (cf *)malloc(h_len * sizeof(cf));
    memcpy(q->h, h, h_len * sizeof(cf));
    q->time_buf = (cf *)malloc(2 * n * sizeof(cf));
    q->freq_buf = (cf *)malloc(2 * n * sizeof(cf));
    q->H = (cf *)malloc(2 * n * sizeof(cf));
    q->w = (cf *)calloc(n, sizeof(cf));
    q->fft = fft_create_plan(2 * n, q->time_buf, q->freq_buf, LIQUID_FFT_FORWARD, 0);
    q->ifft = fft_create_plan(2 * n, q->freq_buf, q->time_buf, LIQUID_FFT_BACKWARD, 0);
    unsigned int i;
    for (i = 0; i < 2 * n; i++) q->time_buf[i] = (i < h_len) ? h[i] : 0;
    fft_execute(q->fft);
    memmove(q->H, q->freq_buf, 2 * n * sizeof(cf));
    q->scale = 1.0f / (float)(2 * n);
    return 0;
}
static void fftfilt_core_free(struct fftfilt_cccf_s *q)
{
    free(q->h); free(q->time_buf); free(q->freq_buf); free(q->H); free(q->w);
    fft_destroy_plan(q->fft); fft_destroy_plan(q->ifft);
}
static void fftfilt_core_execute(struct fftfilt_cccf_s *q, const cf *x, cf *y)
{
    unsigned int i;
    const unsigned int n = q->n;
    for (i = 0; i < n; i++) q->time_buf[i] = x[i];
    for (; i < 2 * n; i++) q->time_buf[i] = 0;
    fft_execute(q->fft);
    for (i = 0; i < 2 * n; i++) q->freq_buf[i] = cf_mul(q->freq_buf[i], q->H[i]);
    fft_execute(q->ifft);
    for (i = 0; i < n; i++) y[i] = cf_scale(q->time_buf[i] + q->w[i], q->scale);
    memmove(q->w, &q->time_buf[n], n * sizeof(cf));
}

unsigned int liquid_compat_fftfilt_get_taps(void *obj, cf *h, unsigned int cap)
{
    struct fftfilt_cccf_s *q = (struct fftfilt_cccf_s *)obj; /* crcf wraps the same core at offset 0 */
    unsigned int i;
    for (i = 0; i < q->h_len && i < cap; i++) h[i] = q->h[i];
    return q->h_len;
}

fftfilt_cccf fftfilt_cccf_create(cf *h, unsigned int h_len, unsigned int n)
{
    fftfilt_cccf q = (fftfilt_cccf)calloc(1, sizeof(*q));
    if (fftfilt_core_init(q, h, h_len, n) != 0) { free(q); return NULL; }
    return q;
}
int fftfilt_cccf_destroy(fftfilt_cccf q) { fftfilt_core_free(q); free(q); return LIQUID_OK; }
int fftfilt_cccf_reset(fftfilt_cccf q) { memset(q->w, 0, q->n * sizeof(cf)); return LIQUID_OK; }
int fftfilt_cccf_execute(fftfilt_cccf q, cf *x, cf *y) { fftfilt_core_execute(q, x, y); return LIQUID_OK; }

fftfilt_crcf fftfilt_crcf_create(float *h, unsigned int h_len, unsigned int n)
{
    fftfilt_crcf q = (fftfilt_crcf)calloc(1, sizeof(*q));
    cf *hc = (cf *)malloc((h_len ? h_len : 1) * sizeof(cf));
    unsigned int i;
    for (i = 0; i < h_len; i++) hc[i] = h[i];
    int rc = fftfilt_core_init(&q->core, hc, h_len, n);
    free(hc);
    if (rc != 0) { free(q); return NULL; }
    return q;
}
int fftfilt_crcf_destroy(fftfilt_crcf q) { fftfilt_core_free(&q->core); free(q); return LIQUID_OK; }
int fftfilt_crcf_reset(fftfilt_crcf q) { memset(q->core.w, 0, q->core.n * sizeof(cf)); return LIQUID_OK; }
int fftfilt_crcf_execute(fftfilt_crcf q, cf *x, cf *y) { fftfilt_core_execute(&q->core, x, y); return LIQUID_OK; }

/* ======================================================================================
 * agc_crcf                                            (liquid: src/agc/src/agc.proto.c)
 * ==================================================================================== */
struct agc_crcf_s {
    float g, scale, bandwidth, alpha, y2_prime;
    int is_locked;
};
agc_crcf agc_crcf_create(void)
{
    agc_crcf q = (agc_crcf)malloc(sizeof(*q));
    q->bandwidth = 1e-2f; q->alpha = 1e-2f;
    q->g = 1.0f; q->y2_prime = 1.0f; q->is_locked = 0; q->scale = 1.0f;
    return q;
}
int agc_crcf_destroy(agc_crcf q) { free(q); return LIQUID_OK; }
int agc_crcf_reset(agc_crcf q) { q->g = 1.0f; q->y2_prime = 1.0f; q->is_locked = 0; return LIQUID_OK; }
int agc_crcf_set_bandwidth(agc_crcf q, float bt)
{
    if (bt < 0 || bt > 1.0f) return -1;
    q->bandwidth = bt; q->alpha = bt;
    return LIQUID_OK;
}
int agc_crcf_set_signal_level(agc_crcf q, float x2)
{
    if (x2 <= 0) return -1;
    q->g = 1.0f / x2;
    q->y2_prime = 1.0f;
    return LIQUID_OK;
}
int agc_crcf_set_gain(agc_crcf q, float gain)
{
    if (gain <= 0) return -1;
    q->g = gain;
    return LIQUID_OK;
}
float agc_crcf_get_gain(agc_crcf q) { return q->g; }
int agc_crcf_execute_block(agc_crcf q, cf *x, unsigned int n, cf *y)
{
    unsigned int i;
    for (i = 0; i < n; i++) {
        float yr = crealf(x[i]) * q->g, yi = cimagf(x[i]) * q->g;
        float y2 = yr * yr + yi * yi;
        q->y2_prime = (1.0 - q->alpha) * q->y2_prime + q->alpha * y2;
        if (!q->is_locked) {
            if (q->y2_prime > 1e-6f)
                q->g *= expf(-0.5f * q->alpha * logf(q->y2_prime));
            if (q->g > 1e6f) q->g = 1e6f;
        }
        y[i] = cf_make(yr * q->scale, yi * q->scale);
    }
    return LIQUID_OK;
}

/* ======================================================================================
 * resamp2_crcf (halfband)                      (liquid: src/filter/src/resamp2.proto.c)
 * ==================================================================================== */
typedef struct {
    unsigned int m, h_len, h1_len;
    float *h, *h1;
    cf *w0, *w1;         /* delay lines of length 2m, linear shift (index 0 = oldest) */
    float f0, as;
} resamp2;

static resamp2 *resamp2_create(unsigned int m, float f0, float as)
{
    resamp2 *q = (resamp2 *)calloc(1, sizeof(*q));
    q->m = m; q->f0 = f0; q->as = as;
    q->h_len = 4 * m + 1;
    q->h = (float *)malloc(q->h_len * sizeof(float));
    q->h1_len = 2 * m;
    q->h1 = (float *)malloc(q->h1_len * sizeof(float));
    unsigned int i;
    float beta = kaiser_beta_As(as);
    for (i = 0; i < q->h_len; i++) {
        float t = (float)i - (float)(q->h_len - 1) / 2.0f;
        float h1 = sincf(t / 2.0f);
        float h2 = liquid_kaiser(i, q->h_len, beta);
        float h3 = cosf(2.0f * M_PI * t * f0);
        q->h[i] = h1 * h2 * h3;
    }
    unsigned int j = 0;
    for (i = 1; i < q->h_len; i += 2) q->h1[j++] = q->h[q->h_len - i - 1];
    q->w0 = (cf *)calloc(2 * m, sizeof(cf));
    q->w1 = (cf *)calloc(2 * m, sizeof(cf));
    return q;
}
static void resamp2_destroy(resamp2 *q) { free(q->h); free(q->h1); free(q->w0); free(q->w1); free(q); }
static void resamp2_reset(resamp2 *q)
{
    memset(q->w0, 0, 2 * q->m * sizeof(cf));
    memset(q->w1, 0, 2 * q->m * sizeof(cf));
}
static inline void win_push(cf *w, unsigned int len, cf x)
{
    memmove(w, w + 1, (len - 1) * sizeof(cf));
    w[len - 1] = x;
}
static void resamp2_decim_execute(resamp2 *q, const cf *x, cf *y)
{
    unsigned int i, L = 2 * q->m;
    win_push(q->w1, L, x[0]);
    float sr = 0.0f, si = 0.0f;
    for (i = 0; i < L; i++) { sr += q->h1[i] * crealf(q->w1[i]); si += q->h1[i] * cimagf(q->w1[i]); }
    win_push(q->w0, L, x[1]);
    cf y0 = q->w0[q->m - 1];
    *y = cf_make(crealf(y0) + sr, cimagf(y0) + si);
}
static void resamp2_interp_execute(resamp2 *q, cf x, cf *y)
{
    unsigned int i, L = 2 * q->m;
    win_push(q->w0, L, x);
    y[0] = q->w0[q->m - 1];
    win_push(q->w1, L, x);
    float sr = 0.0f, si = 0.0f;
    for (i = 0; i < L; i++) { sr += q->h1[i] * crealf(q->w1[i]); si += q->h1[i] * cimagf(q->w1[i]); }
    y[1] = cf_make(sr, si);
}

/* ======================================================================================
 * msresamp2_crcf                              (liquid: src/filter/src/msresamp2.proto.c)
 *   design loop: as_stage = As + 5 dB margin; fc halves per stage (special-cased at i==1)
 * ==================================================================================== */
#ifndef LIQUID_COMPAT_MSRESAMP2_AS_MARGIN
#define LIQUID_COMPAT_MSRESAMP2_AS_MARGIN 5.0f
#endif
typedef struct {
    int is_interp;
    unsigned int num_stages, M;
    float zeta;
    unsigned int m_stage[16];
    float fc_stage[16], f0_stage[16], as_stage[16];
    resamp2 *stage[16];
    cf *buffer0, *buffer1;
} msresamp2;

static msresamp2 *msresamp2_create(int is_interp, unsigned int num_stages, float fc, float f0, float as)
{
    if (num_stages > 16 || fc <= 0.0f || fc >= 0.5f) return NULL;
    msresamp2 *q = (msresamp2 *)calloc(1, sizeof(*q));
    q->is_interp = is_interp;
    q->num_stages = num_stages;
    q->M = 1u << num_stages;
    q->zeta = 1.0f / (float)q->M;
    q->buffer0 = (cf *)calloc(q->M, sizeof(cf));
    q->buffer1 = (cf *)calloc(q->M, sizeof(cf));
    unsigned int i;
    float as_m = as + LIQUID_COMPAT_MSRESAMP2_AS_MARGIN;
    for (i = 0; i < num_stages; i++) {
        fc = (i == 1) ? (0.5 - fc) / 2.0f : 0.5f * fc;
        f0 = 0.5f * f0;
        float ft = 2 * (0.25f - fc);
        unsigned int h_len = estimate_req_filter_len(ft, as_m);
        unsigned int m = ceilf((float)(h_len - 1) / 4.0f);
        q->fc_stage[i] = fc; q->f0_stage[i] = f0; q->as_stage[i] = as_m;
        q->m_stage[i] = m < 3 ? 3 : m;
    }
    for (i = 0; i < num_stages; i++)
        q->stage[i] = resamp2_create(q->m_stage[i], q->f0_stage[i], q->as_stage[i]);
    return q;
}
static void msresamp2_destroy(msresamp2 *q)
{
    unsigned int i;
    for (i = 0; i < q->num_stages; i++) resamp2_destroy(q->stage[i]);
    free(q->buffer0); free(q->buffer1); free(q);
}
static void msresamp2_reset(msresamp2 *q)
{
    unsigned int i;
    for (i = 0; i < q->num_stages; i++) resamp2_reset(q->stage[i]);
}
/* M inputs -> 1 output; highest design index runs first (at the highest rate) */
static void msresamp2_decim_execute(msresamp2 *q, cf *x, cf *y)
{
    cf *b0 = x, *b1 = q->buffer1;
    unsigned int s, k, g, i;
    for (s = 0; s < q->num_stages; s++) {
        k = 1u << (q->num_stages - s - 1);
        g = q->num_stages - s - 1;
        for (i = 0; i < k; i++) resamp2_decim_execute(q->stage[g], &b0[2 * i], &b1[i]);
        b0 = (s % 2) == 0 ? q->buffer1 : q->buffer0;
        b1 = (s % 2) == 0 ? q->buffer0 : q->buffer1;
    }
    *y = cf_scale(b0[0], q->zeta);
}
/* 1 input -> M outputs; design index 0 runs first (at the lowest rate) */
static void msresamp2_interp_execute(msresamp2 *q, cf x, cf *y)
{
    cf *b0 = q->buffer0, *b1 = q->buffer1;
    b0[0] = x;
    if (q->num_stages == 0) { y[0] = x; return; }
    unsigned int s, k, i;
    for (s = 0; s < q->num_stages; s++) {
        k = 1u << s;
        cf *out = (s == q->num_stages - 1) ? y : b1;
        for (i = 0; i < k; i++) resamp2_interp_execute(q->stage[s], b0[i], &out[2 * i]);
        cf *t = b0; b0 = b1; b1 = t;
    }
}

/* ======================================================================================
 * firpfb + resamp_crcf, fixed-point phase  (liquid: firpfb.proto.c, resamp.fixed.proto.c)
 * ==================================================================================== */
typedef struct {
    unsigned int m, npfb, bits_index, h_sub_len;
    float rate, fc, as;
    uint32_t step, phase;
    float *h;     /* prototype, length 2*m*npfb (the +1th tap is dropped by firpfb) */
    float *bank;  /* [npfb][h_sub_len], reversed per sub-filter */
    cf *w;        /* window length h_sub_len, index 0 = oldest */
} resamp;

static resamp *resamp_create(float rate, unsigned int m, float fc, float as, unsigned int npfb)
{
    if (rate <= 0 || m == 0 || fc <= 0.0f || fc >= 0.5f || as <= 0.0f) return NULL;
    resamp *q = (resamp *)calloc(1, sizeof(*q));
    q->m = m; q->fc = fc; q->as = as;
    unsigned int bits = 0;
    while ((1u << bits) < npfb) bits++;
    q->bits_index = bits;
    q->npfb = 1u << bits;
    q->rate = rate;
    q->step = (uint32_t)round((1 << 24) / q->rate);
    unsigned int n = 2 * q->m * q->npfb + 1;
    float *hf = (float *)malloc(n * sizeof(float));
    liquid_firdes_kaiser(n, q->fc / ((float)(q->npfb)), q->as, 0.0f, hf);
    unsigned int i, k;
    float gain = 0.0f;
    for (i = 0; i < n; i++) gain += hf[i];
    gain = (q->npfb) / (gain);
    q->h = (float *)malloc(n * sizeof(float));
    for (i = 0; i < n; i++) q->h[i] = hf[i] * gain;
    free(hf);
    q->h_sub_len = (n - 1) / q->npfb; /* 2m */
    q->bank = (float *)malloc(q->npfb * q->h_sub_len * sizeof(float));
    for (i = 0; i < q->npfb; i++)
        for (k = 0; k < q->h_sub_len; k++)
            q->bank[i * q->h_sub_len + (q->h_sub_len - k - 1)] = q->h[i + k * q->npfb];
    q->w = (cf *)calloc(q->h_sub_len, sizeof(cf));
    q->phase = 0;
    return q;
}
static void resamp_destroy(resamp *q) { free(q->h); free(q->bank); free(q->w); free(q); }
static void resamp_reset(resamp *q) { q->phase = 0; memset(q->w, 0, q->h_sub_len * sizeof(cf)); }
static void resamp_execute(resamp *q, cf x, cf *y, unsigned int *nw)
{
    win_push(q->w, q->h_sub_len, x);
    unsigned int n = 0, i;
    while (q->phase < (1u << 24)) {
        unsigned int index = q->phase >> (24 - q->bits_index);
        const float *hs = &q->bank[index * q->h_sub_len];
        float sr = 0.0f, si = 0.0f;
        for (i = 0; i < q->h_sub_len; i++) { sr += hs[i] * crealf(q->w[i]); si += hs[i] * cimagf(q->w[i]); }
        y[n++] = cf_make(sr, si);
        q->phase += q->step;
    }
    q->phase -= (1u << 24);
    *nw = n;
}

/* ======================================================================================
 * msresamp_crcf                                (liquid: src/filter/src/msresamp.proto.c)
 * ==================================================================================== */
struct msresamp_crcf_s {
    float rate, as;
    int is_interp;
    unsigned int num_halfband_stages;
    msresamp2 *halfband;
    float rate_halfband, rate_arbitrary;
    resamp *arbitrary;
    unsigned int buffer_len, buffer_index;
    cf *buffer;
};

msresamp_crcf msresamp_crcf_create(float r, float as)
{
    if (r <= 0.0f) return NULL;
    msresamp_crcf q = (msresamp_crcf)calloc(1, sizeof(*q));
    q->rate = r; q->as = as;
    q->is_interp = (q->rate > 1.0f) ? 1 : 0;
    q->rate_arbitrary = q->rate;
    q->rate_halfband = 1.0f;
    q->num_halfband_stages = 0;
    if (q->is_interp) {
        while (q->rate_arbitrary > 2.0f) {
            q->num_halfband_stages++;
            q->rate_halfband *= 2.0f;
            q->rate_arbitrary *= 0.5f;
        }
    } else {
        while (q->rate_arbitrary < 0.5f) {
            q->num_halfband_stages++;
            q->rate_halfband *= 0.5f;
            q->rate_arbitrary *= 2.0f;
        }
    }
    q->buffer_len = 4 + (1u << q->num_halfband_stages);
    q->buffer = (cf *)calloc(q->buffer_len, sizeof(cf));
    q->buffer_index = 0;
    q->halfband = msresamp2_create(q->is_interp, q->num_halfband_stages, 0.4f, 0.0f, q->as);
    float fc = 0.515f * q->rate_arbitrary;
    if (fc > 0.49f) fc = 0.49f;
    q->arbitrary = resamp_create(q->rate_arbitrary, 7, fc, q->as, 256);
    if (!q->halfband || !q->arbitrary) { free(q->buffer); free(q); return NULL; }
    return q;
}
int msresamp_crcf_destroy(msresamp_crcf q)
{
    msresamp2_destroy(q->halfband);
    resamp_destroy(q->arbitrary);
    free(q->buffer); free(q);
    return LIQUID_OK;
}
int msresamp_crcf_reset(msresamp_crcf q)
{
    msresamp2_reset(q->halfband);
    resamp_reset(q->arbitrary);
    q->buffer_index = 0;
    return LIQUID_OK;
}
float msresamp_crcf_get_delay(msresamp_crcf q) { (void)q; return 0.0f; }

int msresamp_crcf_execute(msresamp_crcf q, cf *x, unsigned int nx, cf *y, unsigned int *ny_out)
{
    unsigned int i, ny = 0, nw;
    const unsigned int M = 1u << q->num_halfband_stages;
    if (!q->is_interp) {
        for (i = 0; i < nx; i++) {
            q->buffer[q->buffer_index++] = x[i];
            if (q->buffer_index == M) {
                cf hb;
                msresamp2_decim_execute(q->halfband, q->buffer, &hb);
                resamp_execute(q->arbitrary, hb, &y[ny], &nw);
                ny += nw;
                q->buffer_index = 0;
            }
        }
    } else {
        for (i = 0; i < nx; i++) {
            cf tmp[4];
            unsigned int k;
            resamp_execute(q->arbitrary, x[i], tmp, &nw);
            for (k = 0; k < nw; k++) {
                msresamp2_interp_execute(q->halfband, tmp[k], &y[ny]);
                ny += M;
            }
        }
    }
    *ny_out = ny;
    return LIQUID_OK;
}

void liquid_compat_msresamp_get_info(msresamp_crcf q, liquid_compat_msresamp_info *info)
{
    memset(info, 0, sizeof(*info));
    info->is_interp = q->is_interp;
    info->num_halfband = q->num_halfband_stages;
    unsigned int i;
    for (i = 0; i < q->num_halfband_stages && i < 16; i++) {
        info->m_stage[i] = q->halfband->m_stage[i];
        info->as_stage[i] = q->halfband->as_stage[i];
    }
    info->rate_arbitrary = q->rate_arbitrary;
    info->step = q->arbitrary->step;
    info->npfb = q->arbitrary->npfb;
    info->arb_m = q->arbitrary->m;
    info->arb_fc = q->arbitrary->fc;
}
unsigned int liquid_compat_msresamp_get_halfband(msresamp_crcf q, unsigned int i, float *h, unsigned int cap)
{
    if (i >= q->num_halfband_stages) return 0;
    resamp2 *s = q->halfband->stage[i];
    unsigned int n = s->h_len < cap ? s->h_len : cap;
    memcpy(h, s->h, n * sizeof(float));
    return s->h_len;
}
unsigned int liquid_compat_msresamp_get_arb_taps(msresamp_crcf q, float *h, unsigned int cap)
{
    unsigned int len = 2 * q->arbitrary->m * q->arbitrary->npfb;
    unsigned int n = len < cap ? len : cap;
    memcpy(h, q->arbitrary->h, n * sizeof(float));
    return len;
}
