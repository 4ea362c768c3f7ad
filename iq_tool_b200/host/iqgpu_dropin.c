/*
 * iqgpu_dropin.c — context registry, configuration snapshot and the FUSED chunk-train engine of
 * the drop-in host layer.  See iqgpu_dropin.h for the model and INTEGRATION.md for the build.
 *
 * Thread model (unchanged from the reference, src/pipeline.c:99-116): one Pre-Processor thread,
 * one Resampler thread, one Post-Processor thread, connected by the reference's queues.  Here the
 * pre thread is the only producer of staged chunks, the post thread the only executor/consumer;
 * `mu` guards the hand-over between them.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include "iqgpu_dropin.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "constants.h"        /* reference: PIPELINE_CHUNK_BASE_SAMPLES ... */
#include "log.h"              /* reference logger */
#include "signal_handler.h"   /* reference: handle_fatal_thread_error */

#define DROPIN_SLOTS 16
static IqGpuDropin    *g_slots[DROPIN_SLOTS];
static pthread_mutex_t g_reg_mu = PTHREAD_MUTEX_INITIALIZER;

IqGpuDropin *iqgpu_dropin_find(const AppResources *res)
{
    IqGpuDropin *d = NULL;
    pthread_mutex_lock(&g_reg_mu);
    for (int i = 0; i < DROPIN_SLOTS; i++)
        if (g_slots[i] && g_slots[i]->res == res) { d = g_slots[i]; break; }
    pthread_mutex_unlock(&g_reg_mu);
    return d;
}

IqGpuDropin *iqgpu_dropin_get(AppResources *res)
{
    IqGpuDropin *d = iqgpu_dropin_find(res);
    if (d) return d;
    d = (IqGpuDropin *)calloc(1, sizeof(*d));
    if (!d) return NULL;
    d->res = res;
    const char *e = getenv("IQGPU_DROPIN_EAGER");
    d->eager = (e && *e && *e != '0');
    const char *dev = getenv("IQGPU_DEVICE");
    d->device = dev ? atoi(dev) : 0;
    const char *seed = getenv("IQGPU_IQ_SEED");
    d->iq_seed = seed ? (uint32_t)strtoul(seed, NULL, 0) : 20261017u;
    pthread_mutex_init(&d->mu, NULL);
    pthread_cond_init(&d->room, NULL);
    pthread_mutex_lock(&g_reg_mu);
    int placed = 0;
    for (int i = 0; i < DROPIN_SLOTS && !placed; i++)
        if (!g_slots[i]) { g_slots[i] = d; placed = 1; }
    pthread_mutex_unlock(&g_reg_mu);
    if (!placed) { free(d); return NULL; }
    return d;
}

void iqgpu_dropin_addref(AppResources *res)
{
    IqGpuDropin *d = iqgpu_dropin_get(res);
    if (!d) return;
    pthread_mutex_lock(&d->mu);
    d->refs++;
    d->cfg_ready = false;              /* a *_create ran: rebuild the snapshot on next use */
    pthread_mutex_unlock(&d->mu);
}

static void destroy_chains(IqGpuDropin *d)
{
    iqgpu_chain **all[] = {&d->plan, &d->fused, &d->mod_dc, &d->mod_iq, &d->mod_nco, &d->mod_rs, &d->mod_filter, &d->mod_agc};
    for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); i++)
        if (*all[i]) { iqgpu_chain_destroy(*all[i]); *all[i] = NULL; }
}

void iqgpu_dropin_release(AppResources *res)
{
    IqGpuDropin *d = iqgpu_dropin_find(res);
    if (!d) return;
    pthread_mutex_lock(&d->mu);
    const int left = --d->refs;
    pthread_mutex_unlock(&d->mu);
    if (left > 0) return;
    pthread_mutex_lock(&g_reg_mu);
    for (int i = 0; i < DROPIN_SLOTS; i++)
        if (g_slots[i] == d) g_slots[i] = NULL;
    pthread_mutex_unlock(&g_reg_mu);
    destroy_chains(d);
    for (int i = 0; i < 2; i++) iqgpu_host_free(d->stage[i].raw);
    iqgpu_host_free(d->out);
    pthread_mutex_destroy(&d->mu);
    pthread_cond_destroy(&d->room);
    free(d);
}

void iqgpu_dropin_fatal(AppResources *res, const char *what)
{
    char msg[512];
    snprintf(msg, sizeof(msg), "%s (%s)", what, iqgpu_last_error());
    handle_fatal_thread_error(msg, res);    /* sets error_occurred + request_shutdown, signal_handler.c:149 */
}

/* AppConfig/AppResources -> iqgpu_chain_config, the way src/config.c and src/setup.c resolved them */
static void snapshot_config(IqGpuDropin *d)
{
    const AppResources *r = d->res;
    const AppConfig *c = r->config;
    iqgpu_chain_config *g = &d->cfg;
    memset(g, 0, sizeof(*g));
    g->input_format = (int32_t)r->input_format;
    g->output_format = (int32_t)c->output_format;
    g->input_rate_hz = (double)r->source_info.samplerate;
    g->target_rate_hz = c->target_rate;
    g->gain = c->gain;
    g->dc_block_enable = c->dc_block.enable ? 1 : 0;
    g->iq_correction_enable = c->iq_correction.enable ? 1 : 0;
    if (c->iq_correction.enable) {
        const int a = r->iq_correction.active_buffer_idx;
        g->iq_mag = r->iq_correction.factors_buffer[a].mag;
        g->iq_phase = r->iq_correction.factors_buffer[a].phase;
    }
    g->shift_after_resample = c->shift_after_resample;
    /* freq_shift_create resolves nco_shift_hz (frequency_shift.c:33-35); before it ran, the CLI value */
    g->freq_shift_hz = (r->nco_shift_hz != 0.0) ? r->nco_shift_hz : (double)c->freq_shift_hz_arg;
    g->no_resample = (c->no_resample || r->is_passthrough) ? 1 : 0;
    g->num_filter_requests = c->num_filter_requests;
    for (int i = 0; i < c->num_filter_requests && i < IQGPU_MAX_FILTER_CHAIN; i++) {
        g->filter_requests[i].type = (int32_t)c->filter_requests[i].type;
        g->filter_requests[i].freq1_hz = c->filter_requests[i].freq1_hz;
        g->filter_requests[i].freq2_hz = c->filter_requests[i].freq2_hz;
    }
    g->transition_width_hz = c->transition_width_hz_arg;
    g->filter_taps = c->filter_taps_arg;
    g->attenuation_db = c->attenuation_db_arg;
    g->filter_type_request = (c->filter_type_str_arg == NULL) ? IQGPU_FILTER_REQ_AUTO
                           : (c->filter_type_request == FILTER_TYPE_FIR ? IQGPU_FILTER_REQ_FIR : IQGPU_FILTER_REQ_FFT);
    g->filter_fft_size = c->filter_fft_size_arg;
    g->agc_enable = c->output_agc.enable ? 1 : 0;
    g->agc_profile = (int32_t)c->output_agc.profile;
    g->agc_target_level_arg = c->output_agc.target_level_arg > 0 ? c->output_agc.target_level : 0.0f;
}

bool iqgpu_dropin_configure(IqGpuDropin *d)
{
    if (d->cfg_ready) return true;
    snapshot_config(d);
    destroy_chains(d);
    if (iqgpu_chain_create(&d->cfg, -1, &d->plan) != IQGPU_OK) {
        log_fatal("GPU chain: configuration rejected: %s", iqgpu_last_error());
        return false;
    }
    d->cfg_ready = true;
    return true;
}

iqgpu_chain *iqgpu_dropin_module(IqGpuDropin *d, int stage)
{
    if (!iqgpu_dropin_configure(d)) return NULL;
    iqgpu_chain **slot = stage == IQGPU_STAGE_DC ? &d->mod_dc : stage == IQGPU_STAGE_IQ ? &d->mod_iq
                       : stage == IQGPU_STAGE_NCO ? &d->mod_nco : stage == IQGPU_STAGE_RESAMPLER ? &d->mod_rs
                       : stage == IQGPU_STAGE_FILTER ? &d->mod_filter : &d->mod_agc;
    if (!*slot) {
        iqgpu_chain_config g = d->cfg;
        g.stage_select = stage;
        if (iqgpu_chain_create(&g, d->device, slot) != IQGPU_OK) {
            log_fatal("GPU chain: module-level chain (stage %d) failed: %s", stage, iqgpu_last_error());
            *slot = NULL;
        }
        /* the module-level DC blocker reproduces liquid's fp32 state rounding (serial evaluation: on one 16384-frame
         * chunk it costs what a kernel launch costs); IQGPU_DC_EXACT=1 selects the exact-arithmetic scan instead */
        const char *ex = getenv("IQGPU_DC_EXACT");
        if (*slot && stage == IQGPU_STAGE_DC && !(ex && *ex && *ex != '0')) iqgpu_chain_set_option(*slot, "dc_mode", 1);
    }
    return *slot;
}

/* ------------------------------------------------------------------------------------------
 * FUSED engine
 * ---------------------------------------------------------------------------------------- */
static bool ensure_stage_room(IqGpuStageBuf *b, size_t extra)
{
    if (b->bytes + extra <= b->cap_bytes) return true;
    size_t ncap = b->cap_bytes ? b->cap_bytes * 2 : (size_t)512 * PIPELINE_CHUNK_BASE_SAMPLES * 4;
    while (ncap < b->bytes + extra) ncap *= 2;
    unsigned char *p = (unsigned char *)iqgpu_host_alloc(ncap);
    if (!p) return false;
    if (b->bytes) memcpy(p, b->raw, b->bytes);
    iqgpu_host_free(b->raw);
    b->raw = p; b->cap_bytes = ncap;
    return true;
}

static bool ensure_fused(IqGpuDropin *d)
{
    if (!iqgpu_dropin_configure(d)) return false;
    if (d->fused) return true;
    if (iqgpu_chain_create(&d->cfg, d->device, &d->fused) != IQGPU_OK) { d->fused = NULL; return false; }
    /* chunk trains evaluate the DC blocker in exact arithmetic (DC blocker); IQGPU_DC_REFERENCE=1 trades the
     * fused front for liquid's literal fp32 recurrence (serial, ~250 Msamples/s) when bit-level agreement matters more */
    const char *lit = getenv("IQGPU_DC_REFERENCE");
    if (lit && *lit && *lit != '0') iqgpu_chain_set_option(d->fused, "dc_mode", 1);
    return true;
}

bool iqgpu_dropin_stage_chunk(IqGpuDropin *d, SampleChunk *item)
{
    pthread_mutex_lock(&d->mu);
    if (!ensure_fused(d)) { pthread_mutex_unlock(&d->mu); return false; }
    if ((int32_t)item->packet_sample_format != d->cfg.input_format) {
        pthread_mutex_unlock(&d->mu);
        log_error("GPU chain: chunk sample format %d differs from the configured input format %d",
                  (int)item->packet_sample_format, (int)d->cfg.input_format);
        return false;
    }
    while (d->stage[d->fill].n == IQGPU_DROPIN_MAX_PENDING) pthread_cond_wait(&d->room, &d->mu);
    IqGpuStageBuf *b = &d->stage[d->fill];
    const size_t bps = iqgpu_get_bytes_per_sample(d->cfg.input_format);
    const size_t nbytes = (size_t)item->frames_read * bps;
    if (!ensure_stage_room(b, nbytes)) { pthread_mutex_unlock(&d->mu); return false; }
    memcpy(b->raw + b->bytes, item->raw_input_data, nbytes);
    b->bytes += nbytes;
    IqGpuPending *p = &b->pend[b->n++];
    p->frames = (uint32_t)item->frames_read;
    p->reset_before = d->reset_pending ? 1 : 0;
    if (d->reset_pending) d->pre_fft_rem = 0;
    d->reset_pending = false;
    /* a pre-resample FFT filter quantises frames_read to whole blocks (filter.c:491-526) */
    iqgpu_chain_info inf;
    iqgpu_chain_get_info(d->plan, &inf);
    const bool fft = inf.filter_impl == IQGPU_FILTER_IMPL_FFT_SYM || inf.filter_impl == IQGPU_FILTER_IMPL_FFT_ASYM;
    if (fft && !inf.filter_post_resample && item->frames_read > 0) {
        const uint64_t tot = d->pre_fft_rem + (uint64_t)item->frames_read;
        const uint64_t blocks = tot / inf.filter_block_size;
        item->frames_read = (int64_t)(blocks * inf.filter_block_size);
        d->pre_fft_rem = tot - blocks * inf.filter_block_size;
    }
    p->reaches_post = item->frames_read > 0;
    pthread_mutex_unlock(&d->mu);
    return true;
}

unsigned iqgpu_dropin_resampler_count(IqGpuDropin *d, unsigned n_in)
{
    uint64_t o0 = 0, o1 = 0;
    pthread_mutex_lock(&d->mu);
    if (iqgpu_dropin_configure(d)) {
        iqgpu_chain_resampler_outputs_after(d->plan, d->rs_pos, &o0);
        d->rs_pos += n_in;
        iqgpu_chain_resampler_outputs_after(d->plan, d->rs_pos, &o1);
    }
    pthread_mutex_unlock(&d->mu);
    return (unsigned)(o1 - o0);
}

void iqgpu_dropin_mark_reset(IqGpuDropin *d)
{
    pthread_mutex_lock(&d->mu);
    d->reset_pending = true;
    pthread_mutex_unlock(&d->mu);
}

/* run every chunk of `b` through the fused chain; called by the post thread only, `mu` NOT held */
static bool execute_train(IqGpuDropin *d, IqGpuStageBuf *b)
{
    AppResources *r = d->res;
    const size_t in_bps = iqgpu_get_bytes_per_sample(d->cfg.input_format);
    const size_t out_bps = iqgpu_get_bytes_per_sample(d->cfg.output_format);
    iqgpu_chain_info inf;
    iqgpu_chain_get_info(d->plan, &inf);
    uint64_t total_frames = 0;
    for (size_t i = 0; i < b->n; i++) total_frames += b->pend[i].frames;
    const double ratio = d->cfg.no_resample ? 1.0 : (double)inf.ratio;
    const size_t need = ((size_t)ceil((double)total_frames * fmax(1.0, ratio)) + 2 * (size_t)inf.filter_block_size +
                         (size_t)b->n * 4 + 65536) * out_bps;
    if (need > d->out_cap_bytes) {
        iqgpu_host_free(d->out);
        d->out = (unsigned char *)iqgpu_host_alloc(need);
        d->out_cap_bytes = d->out ? need : 0;
        if (!d->out) return false;
    }
    if (d->cfg.iq_correction_enable) {      /* iq_correct.c:146-149: snapshot the active factors */
        pthread_mutex_lock(&r->iq_correction.iq_factors_mutex);
        const int a = r->iq_correction.active_buffer_idx;
        const float mag = r->iq_correction.factors_buffer[a].mag, ph = r->iq_correction.factors_buffer[a].phase;
        pthread_mutex_unlock(&r->iq_correction.iq_factors_mutex);
        iqgpu_chain_set_iq_factors(d->fused, mag, ph);
    }
    static uint32_t frames[IQGPU_DROPIN_MAX_PENDING];   /* post thread only */
    size_t i = 0, in_off = 0, out_off = 0;
    while (i < b->n) {
        if (b->pend[i].reset_before && iqgpu_chain_reset(d->fused) != IQGPU_OK) return false;
        size_t j = i, seg_frames = 0;
        do { frames[j - i] = b->pend[j].frames; seg_frames += b->pend[j].frames; j++; } while (j < b->n && !b->pend[j].reset_before);
        size_t produced = 0;
        if (seg_frames) {
            if (iqgpu_chain_process(d->fused, b->raw + in_off, seg_frames, frames, j - i, d->out + out_off,
                                    d->out_cap_bytes - out_off, &produced, d->out_counts + i) != IQGPU_OK)
                return false;
        } else {
            for (size_t k = i; k < j; k++) d->out_counts[k] = 0;
        }
        in_off += seg_frames * in_bps;
        out_off += produced * out_bps;
        i = j;
    }
    for (size_t k = 0; k < b->n; k++) d->out_reaches_post[k] = b->pend[k].reaches_post;
    d->out_n = b->n; d->out_next = 0; d->out_off_bytes = 0;
    /* AGC state mirrored into AppResources for observers (app_context.h:226-231) */
    iqgpu_chain_info fi;
    if (iqgpu_chain_get_info(d->fused, &fi) == IQGPU_OK) {
        r->agc_is_locked = fi.agc_locked != 0;
        r->agc_current_gain = fi.agc_gain;
        r->agc_peak_memory = fi.agc_peak_memory;
        r->agc_samples_seen = fi.agc_samples_seen;
    }
    b->n = 0; b->bytes = 0;
    return true;
}

bool iqgpu_dropin_finish_chunk(IqGpuDropin *d, SampleChunk *item)
{
    const size_t out_bps = iqgpu_get_bytes_per_sample(d->cfg.output_format);
    for (;;) {
        /* chunks the pre stage dropped (frames_read == 0) never reach the post stage: skip them */
        while (d->out_next < d->out_n && !d->out_reaches_post[d->out_next]) {
            d->out_off_bytes += (size_t)d->out_counts[d->out_next] * out_bps;
            d->out_next++;
        }
        if (d->out_next < d->out_n) break;
        pthread_mutex_lock(&d->mu);
        IqGpuStageBuf *b = &d->stage[d->fill];
        if (b->n == 0) {
            pthread_mutex_unlock(&d->mu);
            log_error("GPU chain: post stage called for a chunk the pre stage never staged");
            return false;
        }
        d->fill ^= 1;                      /* the pre thread carries on in the other buffer */
        pthread_cond_broadcast(&d->room);
        pthread_mutex_unlock(&d->mu);
        if (!execute_train(d, b)) return false;
    }
    const uint32_t count = d->out_counts[d->out_next];
    const size_t nbytes = (size_t)count * out_bps;
    if (nbytes > item->final_output_capacity_bytes) {
        log_error("GPU chain: chunk output (%zu bytes) exceeds final_output_capacity_bytes", nbytes);
        return false;
    }
    memcpy(item->final_output_data, d->out + d->out_off_bytes, nbytes);
    item->frames_to_write = count;
    d->out_off_bytes += nbytes;
    d->out_next++;
    return true;
}
