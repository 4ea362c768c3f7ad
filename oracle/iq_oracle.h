/*
 * iq_oracle.h — TEST INFRASTRUCTURE ONLY. CPU restatement of iq_tool's per-block chain.
 * See iq_oracle.c for the reference file:line each function follows.
 */
#ifndef IQ_ORACLE_H
#define IQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "iq_chain_cfg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float    ratio;
    int32_t  filter_impl;
    int32_t  filter_post_resample;
    uint32_t filter_block_size;
    uint32_t filter_num_taps;
    uint32_t nco_dtheta;
    int32_t  nco_is_post;
    uint32_t cap_samples;
    uint32_t agc_locked;
    float    agc_gain, agc_peak_memory;
    uint64_t agc_samples_seen;
} iqo_info;

void   *iqo_create(const iq_chain_cfg *cfg);
void    iqo_destroy(void *h);
void    iqo_reset(void *h);
int     iqo_process(void *h, const void *raw_in, int64_t n_frames, void *out, int64_t out_cap_bytes,
                    int64_t *out_frames);
void    iqo_set_capture(void *h, int stage, float *buf, int64_t cap_samples);
int64_t iqo_get_capture_len(void *h, int stage);
void    iqo_set_trace(void *h, uint32_t *buf, int64_t cap);
int64_t iqo_get_trace_len(void *h);
void    iqo_get_info(void *h, iqo_info *o);
uint32_t iqo_get_filter_taps(void *h, float *out, uint32_t cap);
void   *iqo_get_msresamp(void *h);
void    iqo_set_fake_clock(int enable, double t);

size_t  iqo_get_bytes_per_sample(int fmt);
int     iqo_convert_block_to_cf32(const void *in, float *out, size_t n, int fmt, float gain);
int     iqo_convert_cf32_to_block(const float *in, void *out, size_t n, int fmt);
int     iqo_iq_optimize(void *h, const float *block1024, unsigned int seed, float *mag, float *phase,
                        float *avg_power, float *power_range);

#ifdef __cplusplus
}
#endif
#endif
