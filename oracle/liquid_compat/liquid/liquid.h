/*
 * liquid_compat — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product.
 *
 * A from-scratch restatement of the subset of the liquid-dsp C API that
 * pclov3r/iq_tool calls on its per-block sample-processing chain.  liquid-dsp
 * (github.com/jgaeddert/liquid-dsp) is an UN-VENDORED, UN-PINNED system
 * dependency of the reference (reference CMakeLists.txt:184-240, README.md:74
 * "libliquid-dev"), so its source is not available in this build environment.
 * The algorithms below restate liquid-dsp's published behaviour (1.3.2 .. 1.6
 * era: fixed-point-phase resamp, uint32-phase LUT NCO) from documentation and
 * recollection; see oracle/README.md for the per-object statement and the
 * "PARITY UNPINNED" caveat at this layer.
 *
 * Reference call sites this header serves (file:line in /root/reference):
 *   resampler.c:27,39,45,51          msresamp_crcf_*
 *   frequency_shift.c:54-119         nco_crcf_*
 *   filter.c:192,209-239             estimate_req_filter_len, liquid_firdes_kaiser, nco_crcf_cexpf/step
 *   filter.c:275-284,348-460         firfilt_crcf_*, firfilt_cccf_*
 *   filter.c:339-344,405-430,513-515 fftfilt_crcf_*, fftfilt_cccf_*
 *   dc_block.c:54,73,82,90           iirfilt_crcf_*
 *   agc.c:39-62,93,228-242           agc_crcf_*
 *   iq_correct.c:116,226,326         fft_create_plan / fft_execute / fft_destroy_plan
 */
#ifndef LIQUID_COMPAT_LIQUID_H
#define LIQUID_COMPAT_LIQUID_H

#include <complex.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float complex liquid_float_complex;

#define LIQUID_OK 0

/* ---- design helpers (firdes / math) ---- */
unsigned int estimate_req_filter_len(float df, float as);
float        kaiser_beta_As(float as);
float        liquid_besseli0f(float z);
float        liquid_lngammaf(float z);
float        liquid_kaiser(unsigned int i, unsigned int wlen, float beta);
float        sincf(float x);
int          liquid_firdes_kaiser(unsigned int n, float fc, float as, float mu, float *h);

/* ---- nco ---- */
typedef enum { LIQUID_NCO = 0, LIQUID_VCO } liquid_ncotype;
typedef struct nco_crcf_s *nco_crcf;
nco_crcf nco_crcf_create(liquid_ncotype type);
int      nco_crcf_destroy(nco_crcf q);
int      nco_crcf_reset(nco_crcf q);
int      nco_crcf_set_frequency(nco_crcf q, float dtheta);
int      nco_crcf_set_phase(nco_crcf q, float phi);
int      nco_crcf_step(nco_crcf q);
int      nco_crcf_cexpf(nco_crcf q, liquid_float_complex *y);
int      nco_crcf_mix_block_up(nco_crcf q, liquid_float_complex *x, liquid_float_complex *y, unsigned int n);
int      nco_crcf_mix_block_down(nco_crcf q, liquid_float_complex *x, liquid_float_complex *y, unsigned int n);
/* compat-only introspection used by tests (not in liquid's API) */
uint32_t liquid_compat_nco_constrain(float theta);
uint32_t liquid_compat_nco_get_theta(nco_crcf q);
uint32_t liquid_compat_nco_get_dtheta(nco_crcf q);

/* ---- iirfilt (DC blocker only) ---- */
typedef struct iirfilt_crcf_s *iirfilt_crcf;
iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha);
int          iirfilt_crcf_destroy(iirfilt_crcf q);
int          iirfilt_crcf_reset(iirfilt_crcf q);
int          iirfilt_crcf_execute_block(iirfilt_crcf q, liquid_float_complex *x, unsigned int n, liquid_float_complex *y);

/* ---- firfilt ---- */
typedef struct firfilt_crcf_s *firfilt_crcf;
typedef struct firfilt_cccf_s *firfilt_cccf;
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n);
int          firfilt_crcf_destroy(firfilt_crcf q);
int          firfilt_crcf_reset(firfilt_crcf q);
int          firfilt_crcf_execute_block(firfilt_crcf q, liquid_float_complex *x, unsigned int n, liquid_float_complex *y);
firfilt_cccf firfilt_cccf_create(liquid_float_complex *h, unsigned int n);
int          firfilt_cccf_destroy(firfilt_cccf q);
int          firfilt_cccf_reset(firfilt_cccf q);
int          firfilt_cccf_execute_block(firfilt_cccf q, liquid_float_complex *x, unsigned int n, liquid_float_complex *y);
int          firfilt_cccf_freqresponse(firfilt_cccf q, float fc, liquid_float_complex *H);

/* compat-only: copy the (forward-order) taps of a filter object; returns tap count */
unsigned int liquid_compat_firfilt_crcf_get_taps(firfilt_crcf q, float *h, unsigned int cap);
unsigned int liquid_compat_firfilt_cccf_get_taps(firfilt_cccf q, liquid_float_complex *h, unsigned int cap);

/* ---- fft ---- */
#define LIQUID_FFT_FORWARD  (+1)
#define LIQUID_FFT_BACKWARD (-1)
typedef struct fftplan_s *fftplan;
fftplan fft_create_plan(unsigned int n, liquid_float_complex *x, liquid_float_complex *y, int dir, int flags);
int     fft_destroy_plan(fftplan p);
int     fft_execute(fftplan p);

/* ---- fftfilt ---- */
typedef struct fftfilt_crcf_s *fftfilt_crcf;
typedef struct fftfilt_cccf_s *fftfilt_cccf;
fftfilt_crcf fftfilt_crcf_create(float *h, unsigned int h_len, unsigned int n);
int          fftfilt_crcf_destroy(fftfilt_crcf q);
int          fftfilt_crcf_reset(fftfilt_crcf q);
int          fftfilt_crcf_execute(fftfilt_crcf q, liquid_float_complex *x, liquid_float_complex *y);
fftfilt_cccf fftfilt_cccf_create(liquid_float_complex *h, unsigned int h_len, unsigned int n);
int          fftfilt_cccf_destroy(fftfilt_cccf q);
int          fftfilt_cccf_reset(fftfilt_cccf q);
int          fftfilt_cccf_execute(fftfilt_cccf q, liquid_float_complex *x, liquid_float_complex *y);

unsigned int liquid_compat_fftfilt_get_taps(void *q /* crcf or cccf */, liquid_float_complex *h, unsigned int cap);

/* ---- agc ---- */
typedef struct agc_crcf_s *agc_crcf;
agc_crcf agc_crcf_create(void);
int      agc_crcf_destroy(agc_crcf q);
int      agc_crcf_reset(agc_crcf q);
int      agc_crcf_set_bandwidth(agc_crcf q, float bt);
int      agc_crcf_set_signal_level(agc_crcf q, float x2);
int      agc_crcf_set_gain(agc_crcf q, float gain);
float    agc_crcf_get_gain(agc_crcf q);
int      agc_crcf_execute_block(agc_crcf q, liquid_float_complex *x, unsigned int n, liquid_float_complex *y);

/* ---- multi-stage resampler ---- */
typedef struct msresamp_crcf_s *msresamp_crcf;
msresamp_crcf msresamp_crcf_create(float r, float as);
int           msresamp_crcf_destroy(msresamp_crcf q);
int           msresamp_crcf_reset(msresamp_crcf q);
int           msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex *x, unsigned int nx,
                                    liquid_float_complex *y, unsigned int *ny);
float         msresamp_crcf_get_delay(msresamp_crcf q);

/* compat-only: expose the resampler's design so tests can compare it against the
 * product's independent host-side design (K0).  Not part of liquid's API. */
typedef struct {
    int          is_interp;          /* 1: rate > 1 */
    unsigned int num_halfband;       /* S */
    unsigned int m_stage[16];        /* semi-length per design index i (i=0 runs at the LOWEST rate when decimating) */
    float        as_stage[16];
    float        rate_arbitrary;
    uint32_t     step;               /* 24-bit fixed-point phase step */
    unsigned int npfb;               /* 256 */
    unsigned int arb_m;              /* 7 */
    float        arb_fc;
} liquid_compat_msresamp_info;
void liquid_compat_msresamp_get_info(msresamp_crcf q, liquid_compat_msresamp_info *info);
/* copy halfband prototype h[0..4m] of design stage i; returns length (4m+1) */
unsigned int liquid_compat_msresamp_get_halfband(msresamp_crcf q, unsigned int i, float *h, unsigned int cap);
/* copy arbitrary prototype h[0..2*m*npfb) (gain-normalised); returns length */
unsigned int liquid_compat_msresamp_get_arb_taps(msresamp_crcf q, float *h, unsigned int cap);

#ifdef __cplusplus
}
#endif
#endif /* LIQUID_COMPAT_LIQUID_H */
