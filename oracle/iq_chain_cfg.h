/*
 * iq_chain_cfg.h — TEST INFRASTRUCTURE ONLY (oracle).
 * Flat, pointer-free description of one iq_tool chain configuration, shared by the
 * reference harness (ref_harness.c, drives the reference's own stage code) and the
 * restated oracle (iq_oracle.c).  Field meanings follow the reference's AppConfig /
 * AppResources (include/app_context.h:66-138, 205-283).  Layout is mirrored by
 * include/iqgpu.h:iqgpu_chain_config so tests can hand one ctypes struct to all three.
 */
#ifndef ORACLE_IQ_CHAIN_CFG_H
#define ORACLE_IQ_CHAIN_CFG_H
#include <stdint.h>

/* numeric values of the reference's format_t (include/common_types.h:33-37) */
enum {
    IQF_UNKNOWN = 0, IQF_U8, IQF_S8, IQF_U16, IQF_S16, IQF_U32, IQF_S32, IQF_F32,
    IQF_CU8, IQF_CS8, IQF_CU16, IQF_CS16, IQF_CS24, IQF_CU32, IQF_CS32, IQF_CF32, IQF_SC16Q11
};
/* FilterType (common_types.h:45-51) */
enum { IQ_FILTER_NONE = 0, IQ_FILTER_LOWPASS, IQ_FILTER_HIGHPASS, IQ_FILTER_PASSBAND, IQ_FILTER_STOPBAND };
/* FilterTypeRequest (common_types.h:61-65) */
enum { IQ_FILTER_REQ_AUTO = 0, IQ_FILTER_REQ_FIR, IQ_FILTER_REQ_FFT };
/* FilterImplementationType (common_types.h:53-59) */
enum { IQ_FILTER_IMPL_NONE = 0, IQ_FILTER_IMPL_FIR_SYM, IQ_FILTER_IMPL_FIR_ASYM, IQ_FILTER_IMPL_FFT_SYM, IQ_FILTER_IMPL_FFT_ASYM };
/* AgcProfile (common_types.h:77-82) */
enum { IQ_AGC_OFF = 0, IQ_AGC_DX, IQ_AGC_LOCAL, IQ_AGC_DIGITAL };

#define IQ_MAX_FILTER_CHAIN 5
#define IQ_CHUNK_SAMPLES 16384 /* PIPELINE_CHUNK_BASE_SAMPLES, constants.h:123 */

typedef struct {
    int32_t type;
    float   freq1_hz;
    float   freq2_hz;
} iq_filter_request;

typedef struct {
    int32_t input_format;
    int32_t output_format;
    double  input_rate_hz;          /* source_info.samplerate (int Hz in the reference) */
    double  target_rate_hz;         /* AppConfig.target_rate */
    float   gain;                   /* AppConfig.gain */
    int32_t dc_block_enable;
    int32_t iq_correction_enable;
    float   iq_mag;                 /* pinned factors_buffer[active].mag */
    float   iq_phase;
    int32_t shift_after_resample;
    double  freq_shift_hz;          /* AppResources.nco_shift_hz */
    int32_t no_resample;
    int32_t num_filter_requests;
    iq_filter_request filter_requests[IQ_MAX_FILTER_CHAIN];
    float   transition_width_hz;    /* 0 = auto */
    int32_t filter_taps;            /* 0 = auto */
    float   attenuation_db;         /* 0 = default (60 dB) */
    int32_t filter_type_request;    /* IQ_FILTER_REQ_* ; AUTO also means "no --filter-type given" */
    int32_t filter_fft_size;        /* 0 = auto */
    int32_t agc_enable;
    int32_t agc_profile;
    float   agc_target_level_arg;   /* 0 = profile default */
    int32_t reserved;
} iq_chain_cfg;

#endif
