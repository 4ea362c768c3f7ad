// design.hpp — K0: host-side filter / resampler / NCO design for the GPU chain.
//
// The reference builds its DSP objects by calling liquid-dsp constructors
// (resampler.c:27 msresamp_crcf_create, filter.c:192-353, frequency_shift.c:54-77,
// dc_block.c:32-54).  The GPU kernels need the same coefficients as plain arrays, so this
// file restates those constructors' arithmetic in float, in the same evaluation order,
// and lays the results out the way the kernels consume them.  It is product code and shares
// nothing with oracle/ (the parity tests compare the two designs tap by tap).
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/iqgpu.h"

namespace iqgpu {

using cfloat = std::complex<float>;

// ---- liquid-dsp design primitives (float arithmetic, liquid's formulas) -----------------
float lq_sincf(float x);
float lq_lngammaf(float z);
float lq_besseli0f(float z);
float lq_kaiser_beta(float as);
float lq_kaiser(unsigned i, unsigned wlen, float beta);
unsigned lq_estimate_req_filter_len(float df, float as);
bool lq_firdes_kaiser(unsigned n, float fc, float as, float mu, float* h);

// ---- NCO (liquid LIQUID_NCO: uint32 phase, 1024-entry sine table) ------------------------
uint32_t nco_constrain(float theta);
void nco_sine_table(float* tab1024);
// d_theta for a shift of |shift_hz| at `rate` (frequency_shift.c:59,76)
uint32_t nco_dtheta_for_shift(double shift_hz, double rate);

// ---- DC blocker (dc_block.c:32, liquid iirfilt dc blocker) ------------------------------
struct DcPlan {
    bool enable = false;
    float alpha = 0.f;   // (float)(2 pi 10 / Fs)
    float c = 0.f;       // pole: -a1 = 1 - alpha rounded to float the way liquid forms it
    float one_minus_c = 0.f;
};
DcPlan design_dc(bool enable, int samplerate);

// ---- multi-stage resampler (liquid msresamp_crcf / msresamp2 / resamp2 / resamp fixed) ---
struct HalfbandStage {
    unsigned m = 0;              // semi-length; prototype length 4m+1
    float as = 0.f;
    std::vector<float> h;        // prototype h[0..4m]
    std::vector<float> h1;       // 2m dot-product taps, oldest sample first: h1[j] = h[4m-1-2j]
};
struct ResamplerPlan {
    bool passthrough = false;    // no_resample
    bool is_interp = false;
    float ratio = 1.f;
    unsigned num_halfband = 0;   // S
    std::vector<HalfbandStage> stages;   // by DESIGN index i (decim: executed S-1 .. 0)
    float zeta = 1.f;            // 2^-S (decimation only)
    float rate_arbitrary = 1.f;
    uint32_t step = 1u << 24;
    unsigned npfb = 256, arb_m = 7, arb_sub_len = 14;
    float arb_fc = 0.f;
    std::vector<float> arb_h;    // prototype, 2*m*npfb taps (gain-normalised)
    std::vector<float> bank;     // [npfb][14] dot-product order (oldest first): bank[i][k] = h[i + (13-k)*npfb]
    // raw-input history (in input frames) a stateless re-computation needs (decimation)
    uint64_t halo_input_frames = 0;
};
bool design_resampler(float ratio, float as, bool passthrough, ResamplerPlan& out, std::string& err);

// number of arbitrary-stage outputs emitted after K pushes: ceil(K * 2^24 / step)
uint64_t arb_outputs_after(uint64_t pushes, uint32_t step);
// total chain-resampler outputs after n_in input frames since reset
uint64_t resampler_outputs_after(const ResamplerPlan& p, uint64_t n_in);

// ---- user filter (filter.c:43-393) ---------------------------------------------------------
struct FilterPlan {
    int impl = IQGPU_FILTER_IMPL_NONE;
    bool post_resample = false;
    bool is_complex = false;
    unsigned block = 0;                 // FFT block n (FFT size 2n)
    std::vector<cfloat> taps;           // master taps (imag == 0 for symmetric designs)
};
bool design_filter(const iqgpu_chain_config& cfg, int in_rate, double target_rate, FilterPlan& out, std::string& err);

}  // namespace iqgpu
