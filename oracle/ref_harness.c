/*
 * ref_harness.c — TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product.
 *
 * Drives the REFERENCE'S OWN stage code — compiled in place from /root/reference/src
 * ({sample_convert,dc_block,iq_correct,frequency_shift,resampler,filter,agc,pre_processor,
 * post_processor,memory_arena,log}.c; see oracle/Makefile) against the liquid_compat shim —
 * through a flat C API that Python can call.  This file contains no DSP: it only builds the
 * reference's AppConfig/AppResources/SampleChunk (include/app_context.h,
 * include/pipeline_types.h), creates the DSP objects in the order of
 * _create_dsp_components (src/pipeline.c:138-147), sizes the chunk buffers like
 * _allocate_processing_buffers (src/pipeline.c:232-309) and then walks 16384-frame chunks
 * through the three stage calls exactly as the stage threads do
 * (src/pipeline.c:436-490, 492-537, 539-595).
 *
 * It is compiled ONLY where /root/reference exists; the resulting oracle/_ref/libiqref.so
 * is git-ignored and travels to the GPU box as a prebuilt checker.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "app_context.h"
#include "constants.h"
#include "dc_block.h"
#include "iq_correct.h"
#include "frequency_shift.h"
#include "resampler.h"
#include "filter.h"
#include "agc.h"
#include "pre_processor.h"
#include "post_processor.h"
#include "sample_convert.h"
#include "memory_arena.h"
#include "log.h"

/* IQ_HARNESS_NO_LIQUID: the same harness linked against the GPU drop-in host layer
 * (iq_tool_b200/host/ + libiqgpu.so) instead of the reference's stage sources; the DSP objects
 * are then not liquid objects, so liquid introspection is compiled out. */
#ifndef IQ_HARNESS_NO_LIQUID
#include "liquid/liquid.h"
#endif
#include "iq_chain_cfg.h"

/* ---------------------------------------------------------------------------------------
 * Symbols the reference's stage files expect from translation units we do not compile
 * (utils.c, signal_handler.c, libsndfile).
 * ------------------------------------------------------------------------------------- */
static int    g_fake_clock_enabled = 0;
static double g_fake_clock = 0.0;
static int    g_fatal_count = 0;

double get_monotonic_time_sec(void) /* utils.c:49 */
{
    if (g_fake_clock_enabled) return g_fake_clock;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec / 1e9;
}
void iqref_set_fake_clock(int enable, double t) { g_fake_clock_enabled = enable; g_fake_clock = t; }

void handle_fatal_thread_error(const char *context_msg, AppResources *resources) /* signal_handler.c:149 */
{
    fprintf(stderr, "iqref: fatal thread error: %s\n", context_msg);
    if (resources) resources->error_occurred = true;
    g_fatal_count++;
}
sf_count_t sf_read_raw(SNDFILE *s, void *p, sf_count_t b) { (void)s; (void)p; (void)b; abort(); }
sf_count_t sf_seek(SNDFILE *s, sf_count_t f, int w) { (void)s; (void)f; (void)w; abort(); }

/* ---------------------------------------------------------------------------------------
 * Handle
 * ------------------------------------------------------------------------------------- */
#define IQREF_POOL 8

typedef struct {
    AppConfig    config;
    AppResources res;
    MemoryArena  arena;
    float        ratio;
    size_t       in_bps, out_bps, cap;
    SampleChunk  chunk[IQREF_POOL];
    void        *slab;
    /* optional capture of intermediate cf32 streams */
    complex_float_t *cap_buf[3];
    int64_t          cap_cap[3], cap_len[3];
    /* per-chunk frames_to_write trace */
    uint32_t *trace; int64_t trace_cap, trace_len;
} iqref_t;

static void fill_config(iqref_t *h, const iq_chain_cfg *c)
{
    AppConfig *cfg = &h->config;
    memset(cfg, 0, sizeof(*cfg));
    cfg->gain = c->gain;
    cfg->gain_provided = (c->gain != 1.0f);
    cfg->freq_shift_hz_arg = (float)c->freq_shift_hz;
    cfg->shift_after_resample = c->shift_after_resample;
    cfg->no_resample = c->no_resample;
    cfg->iq_correction.enable = c->iq_correction_enable != 0;
    cfg->dc_block.enable = c->dc_block_enable != 0;
    cfg->output_agc.enable = c->agc_enable != 0;
    cfg->output_agc.profile = (AgcProfile)c->agc_profile;
    cfg->output_agc.target_level_arg = c->agc_target_level_arg;
    /* config.c:320-341: resolve target level */
    if (c->agc_target_level_arg != 0.0f) {
        cfg->output_agc.target_level = c->agc_target_level_arg;
    } else {
        switch (cfg->output_agc.profile) {
            case AGC_PROFILE_DIGITAL: cfg->output_agc.target_level = AGC_DIGITAL_PEAK_TARGET; break;
            case AGC_PROFILE_LOCAL:   cfg->output_agc.target_level = AGC_LOCAL_TARGET; break;
            default:                  cfg->output_agc.target_level = AGC_DX_TARGET; break;
        }
    }
    cfg->num_filter_requests = c->num_filter_requests;
    for (int i = 0; i < c->num_filter_requests && i < MAX_FILTER_CHAIN; i++) {
        cfg->filter_requests[i].type = (FilterType)c->filter_requests[i].type;
        cfg->filter_requests[i].freq1_hz = c->filter_requests[i].freq1_hz;
        cfg->filter_requests[i].freq2_hz = c->filter_requests[i].freq2_hz;
    }
    cfg->transition_width_hz_arg = c->transition_width_hz;
    cfg->filter_taps_arg = c->filter_taps;
    cfg->attenuation_db_arg = c->attenuation_db;
    cfg->filter_fft_size_arg = c->filter_fft_size;
    if (c->filter_type_request == IQ_FILTER_REQ_FIR) {
        cfg->filter_type_str_arg = "fir"; cfg->filter_type_request = FILTER_TYPE_FIR;
    } else if (c->filter_type_request == IQ_FILTER_REQ_FFT) {
        cfg->filter_type_str_arg = "fft"; cfg->filter_type_request = FILTER_TYPE_FFT;
    } else {
        cfg->filter_type_str_arg = NULL; cfg->filter_type_request = FILTER_TYPE_AUTO;
    }
    cfg->output_format = (format_t)c->output_format;
    cfg->target_rate = c->target_rate_hz;
}

void iqref_destroy(void *hv);

void *iqref_create(const iq_chain_cfg *c)
{
    log_set_level(LOG_ERROR);
    iqref_t *h = (iqref_t *)calloc(1, sizeof(*h));
    if (!h) return NULL;
    fill_config(h, c);
    AppResources *r = &h->res;
    r->config = &h->config;
    r->source_info.samplerate = (int)c->input_rate_hz;
    r->source_info.frames = -1;
    r->input_format = (format_t)c->input_format;
    r->input_bytes_per_sample_pair = get_bytes_per_sample(r->input_format);
    pthread_mutex_init(&r->progress_mutex, NULL);
    if (!mem_arena_init(&h->arena, 64u * 1024u * 1024u)) { free(h); return NULL; }

    /* setup.c:91-122 calculate_and_validate_resample_ratio */
    if (h->config.no_resample) {
        h->config.target_rate = (double)r->source_info.samplerate;
        r->is_passthrough = true;
    }
    float ratio = (float)(h->config.target_rate / (double)r->source_info.samplerate);
    if (!isfinite(ratio) || ratio < MIN_ACCEPTABLE_RATIO || ratio > MAX_ACCEPTABLE_RATIO) goto fail;
    h->ratio = ratio;
    r->resample_ratio = ratio;

    /* pipeline.c:138-147 _create_dsp_components */
    if (!dc_block_create(&h->config, r)) goto fail;
    if (!iq_correct_init(&h->config, r, &h->arena)) goto fail;
    if (h->config.iq_correction.enable) {
        r->iq_correction.factors_buffer[0].mag = c->iq_mag;
        r->iq_correction.factors_buffer[0].phase = c->iq_phase;
        r->iq_correction.factors_buffer[1] = r->iq_correction.factors_buffer[0];
    }
    /* a shift that is not a float (WAV centre-target metadata, input_wav.c:614-628) arrives in nco_shift_hz itself;
     * a float --freq-shift is widened by freq_shift_create (frequency_shift.c:32-33) */
    r->nco_shift_hz = ((double)(float)c->freq_shift_hz != c->freq_shift_hz) ? c->freq_shift_hz : 0.0;
    if (!freq_shift_create(&h->config, r)) goto fail;
    r->resampler = create_resampler(&h->config, r, ratio);
    if (!r->resampler && !r->is_passthrough) goto fail;
    if (!filter_create(&h->config, r, &h->arena)) goto fail;
    if (!agc_create(&h->config, r)) goto fail;

    /* pipeline.c:232-265 buffer capacity */
    size_t max_pre = PIPELINE_CHUNK_BASE_SAMPLES;
    bool is_fft = (r->user_filter_type_actual == FILTER_IMPL_FFT_SYMMETRIC ||
                   r->user_filter_type_actual == FILTER_IMPL_FFT_ASYMMETRIC);
    if (r->user_filter_object && !h->config.apply_user_filter_post_resample && is_fft &&
        r->user_filter_block_size > max_pre)
        max_pre = r->user_filter_block_size;
    size_t rs_cap = (size_t)ceil((double)max_pre * fmax(1.0, (double)ratio)) + RESAMPLER_OUTPUT_SAFETY_MARGIN;
    size_t cap = max_pre > rs_cap ? max_pre : rs_cap;
    if (r->user_filter_object && h->config.apply_user_filter_post_resample && is_fft &&
        r->user_filter_block_size > cap)
        cap = r->user_filter_block_size;
    /* The reference under-sizes the post-FFT scratch; the harness gives head room
     * of one FFT block so the reference code cannot overrun while being used as an oracle. */
    if (is_fft) cap += r->user_filter_block_size;
    h->cap = cap;
    r->max_out_samples = (unsigned int)cap;
    h->in_bps = r->input_bytes_per_sample_pair;
    h->out_bps = get_bytes_per_sample(h->config.output_format);
    r->output_bytes_per_sample_pair = h->out_bps;

    size_t raw_b = PIPELINE_CHUNK_BASE_SAMPLES * h->in_bps;
    size_t cpx_b = cap * sizeof(complex_float_t);
    size_t out_b = cap * h->out_bps;
    size_t per = raw_b + 2 * cpx_b + out_b;
    per = (per + 63) & ~(size_t)63;
    h->slab = aligned_alloc(64, per * IQREF_POOL);
    if (!h->slab) goto fail;
    memset(h->slab, 0, per * IQREF_POOL);
    for (int i = 0; i < IQREF_POOL; i++) {
        SampleChunk *it = &h->chunk[i];
        char *base = (char *)h->slab + (size_t)i * per;
        memset(it, 0, sizeof(*it));
        it->raw_input_data = base;
        it->complex_sample_buffer_a = (complex_float_t *)(base + raw_b);
        it->complex_sample_buffer_b = (complex_float_t *)(base + raw_b + cpx_b);
        it->final_output_data = (unsigned char *)(base + raw_b + 2 * cpx_b);
        it->raw_input_capacity_bytes = raw_b;
        it->complex_buffer_capacity_samples = cap;
        it->final_output_capacity_bytes = out_b;
        it->input_bytes_per_sample_pair = h->in_bps;
    }
    return h;
fail:
    iqref_destroy(h);
    return NULL;
}

void iqref_destroy(void *hv)
{
    iqref_t *h = (iqref_t *)hv;
    if (!h) return;
    AppResources *r = &h->res;
    /* pipeline.c:149-157 reverse order */
    agc_destroy(r);
    filter_destroy(r);
    destroy_resampler(r->resampler);
    freq_shift_destroy_ncos(r);
    if (r->config) iq_correct_destroy(r);
    dc_block_destroy(r);
    mem_arena_destroy(&h->arena);
    free(h->slab);
    free(h);
}

/* stream discontinuity: what the three stage threads do on a marker chunk (pipeline.c:458-571) */
void iqref_reset(void *hv)
{
    iqref_t *h = (iqref_t *)hv;
    pre_processor_reset(&h->res);
    resampler_reset(h->res.resampler);
    post_processor_reset(&h->res);
}

void iqref_set_capture(void *hv, int stage, float *buf, int64_t cap_samples)
{
    iqref_t *h = (iqref_t *)hv;
    if (stage < 0 || stage > 2) return;
    h->cap_buf[stage] = (complex_float_t *)buf;
    h->cap_cap[stage] = cap_samples;
    h->cap_len[stage] = 0;
}
int64_t iqref_get_capture_len(void *hv, int stage) { return ((iqref_t *)hv)->cap_len[stage]; }
void iqref_set_trace(void *hv, uint32_t *buf, int64_t cap) { iqref_t *h = hv; h->trace = buf; h->trace_cap = cap; h->trace_len = 0; }
int64_t iqref_get_trace_len(void *hv) { return ((iqref_t *)hv)->trace_len; }

static void capture(iqref_t *h, int stage, const complex_float_t *p, size_t n)
{
    if (!h->cap_buf[stage]) return;
    int64_t room = h->cap_cap[stage] - h->cap_len[stage];
    if ((int64_t)n > room) n = room > 0 ? (size_t)room : 0;
    memcpy(h->cap_buf[stage] + h->cap_len[stage], p, n * sizeof(complex_float_t));
    h->cap_len[stage] += (int64_t)n;
}

/* ---- the three stage bodies, as the stage threads perform them ---- */
static void stage_pre(iqref_t *h, SampleChunk *it)
{
    pre_processor_apply_chain(&h->res, it);                       /* pipeline.c:466 */
}
static void stage_resample(iqref_t *h, SampleChunk *it)
{
    AppResources *r = &h->res;
    it->current_input_buffer = it->complex_sample_buffer_a;       /* pipeline.c:512-513 */
    it->current_output_buffer = it->complex_sample_buffer_b;
    unsigned int nout = 0;
    if (r->is_passthrough) {
        nout = (unsigned int)it->frames_read;
        memcpy(it->current_output_buffer, it->current_input_buffer, nout * sizeof(complex_float_t));
    } else {
        resampler_execute(r->resampler, it->current_input_buffer, (unsigned int)it->frames_read,
                          it->current_output_buffer, &nout);      /* pipeline.c:521 */
    }
    it->frames_to_write = nout;
    it->current_input_buffer = it->complex_sample_buffer_b;       /* pipeline.c:527-528 */
    it->current_output_buffer = it->complex_sample_buffer_a;
}
static void stage_post(iqref_t *h, SampleChunk *it)
{
    post_processor_apply_chain(&h->res, it);                      /* pipeline.c:573 */
}

/* Process n_frames of raw input; append converted output to `out`. Returns 0 on success. */
int iqref_process(void *hv, const void *raw_in, int64_t n_frames, void *out, int64_t out_cap_bytes,
                  int64_t *out_frames)
{
    iqref_t *h = (iqref_t *)hv;
    SampleChunk *it = &h->chunk[0];
    const char *src = (const char *)raw_in;
    char *dst = (char *)out;
    int64_t done = 0, written = 0;
    g_fatal_count = 0;
    while (done < n_frames) {
        int64_t n = n_frames - done;
        if (n > PIPELINE_CHUNK_BASE_SAMPLES) n = PIPELINE_CHUNK_BASE_SAMPLES;
        memcpy(it->raw_input_data, src + done * h->in_bps, (size_t)n * h->in_bps);
        it->frames_read = n;
        it->frames_to_write = 0;
        it->packet_sample_format = h->res.input_format;
        it->is_last_chunk = false;
        it->stream_discontinuity_event = false;

        stage_pre(h, it);
        if (it->frames_read > 0) {
            capture(h, 0, it->complex_sample_buffer_a, (size_t)it->frames_read);
            stage_resample(h, it);
            capture(h, 1, it->current_input_buffer, it->frames_to_write);
            stage_post(h, it);
        }
        if (h->trace && h->trace_len < h->trace_cap) h->trace[h->trace_len++] = it->frames_to_write;
        if (it->frames_to_write > 0) {
            size_t nb = (size_t)it->frames_to_write * h->out_bps;
            if (written * (int64_t)h->out_bps + (int64_t)nb > out_cap_bytes) return -2;
            memcpy(dst + written * h->out_bps, it->final_output_data, nb);
            written += it->frames_to_write;
        }
        done += n;
    }
    *out_frames = written;
    return g_fatal_count ? -1 : 0;
}

/* ---------------------------------------------------------------------------------------
 * Threaded runner: Reader(caller) -> Pre -> Resampler -> Post -> Writer(caller-side copy),
 * one pthread per compute stage like pipeline.c:99-116.  Used only for CPU-baseline timing.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    SampleChunk *q[IQREF_POOL + 1];
    int head, tail, count, closed;
    pthread_mutex_t mu; pthread_cond_t ne, nf;
} bq_t;
static void bq_init(bq_t *q) { memset(q, 0, sizeof(*q)); pthread_mutex_init(&q->mu, NULL); pthread_cond_init(&q->ne, NULL); pthread_cond_init(&q->nf, NULL); }
static void bq_put(bq_t *q, SampleChunk *c)
{
    pthread_mutex_lock(&q->mu);
    while (q->count == IQREF_POOL + 1) pthread_cond_wait(&q->nf, &q->mu);
    q->q[q->tail] = c; q->tail = (q->tail + 1) % (IQREF_POOL + 1); q->count++;
    pthread_cond_signal(&q->ne);
    pthread_mutex_unlock(&q->mu);
}
static SampleChunk *bq_get(bq_t *q)
{
    pthread_mutex_lock(&q->mu);
    while (q->count == 0) pthread_cond_wait(&q->ne, &q->mu);
    SampleChunk *c = q->q[q->head]; q->head = (q->head + 1) % (IQREF_POOL + 1); q->count--;
    pthread_cond_signal(&q->nf);
    pthread_mutex_unlock(&q->mu);
    return c;
}
typedef struct { iqref_t *h; bq_t *in, *out; int stage; } worker_t;
static void *worker(void *arg)
{
    worker_t *w = (worker_t *)arg;
    for (;;) {
        SampleChunk *it = bq_get(w->in);
        if (it->is_last_chunk) { bq_put(w->out, it); break; }
        if (w->stage == 0) stage_pre(w->h, it);
        else if (w->stage == 1) { if (it->frames_read > 0) stage_resample(w->h, it); }
        else { if (it->frames_read > 0) stage_post(w->h, it); }
        bq_put(w->out, it);
    }
    return NULL;
}

int iqref_process_threaded(void *hv, const void *raw_in, int64_t n_frames, void *out, int64_t out_cap_bytes,
                           int64_t *out_frames)
{
    iqref_t *h = (iqref_t *)hv;
    bq_t freeq, q0, q1, q2, q3;
    bq_init(&freeq); bq_init(&q0); bq_init(&q1); bq_init(&q2); bq_init(&q3);
    SampleChunk marker; memset(&marker, 0, sizeof(marker)); marker.is_last_chunk = true;
    for (int i = 0; i < IQREF_POOL; i++) bq_put(&freeq, &h->chunk[i]);
    worker_t w[3] = {{h, &q0, &q1, 0}, {h, &q1, &q2, 1}, {h, &q2, &q3, 2}};
    pthread_t th[3];
    for (int i = 0; i < 3; i++) pthread_create(&th[i], NULL, worker, &w[i]);

    const char *src = (const char *)raw_in;
    char *dst = (char *)out;
    int64_t fed = 0, written = 0, inflight = 0;
    int rc = 0, eos_sent = 0, eos_seen = 0;
    g_fatal_count = 0;
    while (!eos_seen) {
        /* reader side: keep the pipe full */
        while (!eos_sent && inflight < IQREF_POOL) {
            if (fed >= n_frames) { bq_put(&q0, &marker); eos_sent = 1; break; }
            SampleChunk *it = bq_get(&freeq);
            int64_t n = n_frames - fed;
            if (n > PIPELINE_CHUNK_BASE_SAMPLES) n = PIPELINE_CHUNK_BASE_SAMPLES;
            memcpy(it->raw_input_data, src + fed * h->in_bps, (size_t)n * h->in_bps);
            it->frames_read = n; it->frames_to_write = 0;
            it->packet_sample_format = h->res.input_format;
            it->is_last_chunk = false; it->stream_discontinuity_event = false;
            bq_put(&q0, it);
            fed += n; inflight++;
        }
        /* writer side */
        SampleChunk *it = bq_get(&q3);
        if (it->is_last_chunk) { eos_seen = 1; break; }
        if (it->frames_to_write > 0) {
            size_t nb = (size_t)it->frames_to_write * h->out_bps;
            if (written * (int64_t)h->out_bps + (int64_t)nb > out_cap_bytes) rc = -2;
            else { memcpy(dst + written * h->out_bps, it->final_output_data, nb); written += it->frames_to_write; }
        }
        bq_put(&freeq, it); inflight--;
    }
    for (int i = 0; i < 3; i++) pthread_join(th[i], NULL);
    *out_frames = written;
    if (g_fatal_count) rc = -1;
    return rc;
}

/* ---------------------------------------------------------------------------------------
 * Introspection / stage-level entry points
 * ------------------------------------------------------------------------------------- */
typedef struct {
    float    ratio;
    int32_t  filter_impl;          /* FilterImplementationType */
    int32_t  filter_post_resample;
    uint32_t filter_block_size;
    uint32_t filter_num_taps;
    uint32_t nco_dtheta;           /* pre or post NCO, 0 if none */
    int32_t  nco_is_post;
    uint32_t cap_samples;
    uint32_t agc_locked;
    float    agc_gain, agc_peak_memory;
    uint64_t agc_samples_seen;
} iqref_info;

void iqref_get_info(void *hv, iqref_info *o)
{
    iqref_t *h = (iqref_t *)hv;
    AppResources *r = &h->res;
    memset(o, 0, sizeof(*o));
    o->ratio = h->ratio;
    o->filter_impl = r->user_filter_type_actual;
    o->filter_post_resample = h->config.apply_user_filter_post_resample;
    o->filter_block_size = r->user_filter_block_size;
#ifndef IQ_HARNESS_NO_LIQUID
    void *nco = r->pre_resample_nco ? r->pre_resample_nco : r->post_resample_nco;
    if (nco) o->nco_dtheta = liquid_compat_nco_get_dtheta((nco_crcf)nco);
#endif
    o->nco_is_post = r->post_resample_nco != NULL;
    o->cap_samples = (uint32_t)h->cap;
    o->agc_locked = r->agc_is_locked;
    o->agc_gain = r->agc_current_gain;
    o->agc_peak_memory = r->agc_peak_memory;
    o->agc_samples_seen = r->agc_samples_seen;
#ifndef IQ_HARNESS_NO_LIQUID
    if (r->user_filter_object) {
        switch (r->user_filter_type_actual) {
            case FILTER_IMPL_FIR_SYMMETRIC:  o->filter_num_taps = liquid_compat_firfilt_crcf_get_taps(r->user_filter_object, NULL, 0); break;
            case FILTER_IMPL_FIR_ASYMMETRIC: o->filter_num_taps = liquid_compat_firfilt_cccf_get_taps(r->user_filter_object, NULL, 0); break;
            default: o->filter_num_taps = liquid_compat_fftfilt_get_taps(r->user_filter_object, NULL, 0); break;
        }
    }
#endif
}
/* master taps as interleaved complex floats */
uint32_t iqref_get_filter_taps(void *hv, float *out, uint32_t cap)
{
    iqref_t *h = (iqref_t *)hv;
    AppResources *r = &h->res;
    if (!r->user_filter_object) return 0;
#ifdef IQ_HARNESS_NO_LIQUID
    (void)out; (void)cap;
    return 0;
#else
    liquid_float_complex *o = (liquid_float_complex *)out;
    if (r->user_filter_type_actual == FILTER_IMPL_FIR_SYMMETRIC) {
        uint32_t n = liquid_compat_firfilt_crcf_get_taps(r->user_filter_object, NULL, 0);
        float *t = (float *)malloc(n * sizeof(float));
        liquid_compat_firfilt_crcf_get_taps(r->user_filter_object, t, n);
        for (uint32_t i = 0; i < n && i < cap; i++) o[i] = t[i];
        free(t);
        return n;
    }
    if (r->user_filter_type_actual == FILTER_IMPL_FIR_ASYMMETRIC)
        return liquid_compat_firfilt_cccf_get_taps(r->user_filter_object, o, cap);
    return liquid_compat_fftfilt_get_taps(r->user_filter_object, o, cap);
#endif
}
void *iqref_get_msresamp(void *hv) { return ((iqref_t *)hv)->res.resampler; }

/* sample_convert.c entry points (true reference code, no liquid involved) */
size_t iqref_get_bytes_per_sample(int fmt) { return get_bytes_per_sample((format_t)fmt); }
int iqref_convert_block_to_cf32(const void *in, float *out, size_t n, int fmt, float gain)
{
    return convert_block_to_cf32(in, (complex_float_t *)out, n, (format_t)fmt, gain) ? 0 : -1;
}
int iqref_convert_cf32_to_block(const float *in, void *out, size_t n, int fmt)
{
    return convert_cf32_to_block((const complex_float_t *)in, out, n, (format_t)fmt) ? 0 : -1;
}

/* iq_correct.c optimizer with a seeded rand() so the hill-climb is reproducible */
int iqref_iq_optimize(void *hv, const float *block1024, unsigned int seed, float *mag, float *phase,
                      float *avg_power, float *power_range)
{
    iqref_t *h = (iqref_t *)hv;
    AppResources *r = &h->res;
    if (!h->config.iq_correction.enable) return -1;
    r->iq_correction.last_optimization_time = -1e9;
    srand(seed);
    iq_correct_run_optimization(r, (const complex_float_t *)block1024);
    int a = r->iq_correction.active_buffer_idx;
    *mag = r->iq_correction.factors_buffer[a].mag;
    *phase = r->iq_correction.factors_buffer[a].phase;
    *avg_power = r->iq_correction.average_power;
    *power_range = r->iq_correction.power_range;
    return 0;
}
