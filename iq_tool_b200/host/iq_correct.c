/*
 * iq_correct.c (GPU drop-in) — replaces reference src/iq_correct.c
 * (include/iq_correct.h:32,44,56,65,76).  The correction itself (re' = re (1+mag),
 * im' = im + phase re) is part of K1; one optimiser pass (1024-point windowed FFT asymmetry
 * metric, 1 + 25 evaluations, 5 % smoothing; iq_correct.c:154-235, 315-393) is kernel K6.
 * The double-buffered factors and their mutex stay in AppResources exactly as in the reference,
 * so the optimiser thread (src/utility_threads.c:35-47) needs no change.
 */
#include "iq_correct.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "constants.h"
#include "iqgpu_dropin.h"
#include "log.h"
#include "pre_processor.h"
#include "utils.h"          /* reference: get_monotonic_time_sec */

bool iq_correct_init(AppConfig *config, AppResources *resources, MemoryArena *arena)
{
    (void)arena;
    memset(&resources->iq_correction.factors_buffer, 0, sizeof(resources->iq_correction.factors_buffer));
    resources->iq_correction.fft_plan = NULL;
    if (!config->iq_correction.enable) return true;
    srand((unsigned int)time(NULL));                        /* iq_correct.c:92 */
    if (pthread_mutex_init(&resources->iq_correction.iq_factors_mutex, NULL) != 0) {
        log_fatal("Failed to initialize I/Q correction mutex.");
        return false;
    }
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) { log_fatal("Failed to create GPU I/Q correction object."); return false; }
    iqgpu_dropin_addref(resources);
    resources->iq_correction.fft_plan = d;                  /* opaque slot: "initialised" */
    resources->iq_correction.active_buffer_idx = 0;
    resources->iq_correction.average_power = 0.0f;
    resources->iq_correction.power_range = 0.0f;
    resources->iq_correction.samples_in_accum = 0;
    resources->iq_correction.last_optimization_time = -1e9;     /* sample clock starts at 0: the first probe passes the gate, as in the reference */
    log_info("I/Q Correction enabled");
    return true;
}

void iq_correct_apply(AppResources *resources, complex_float_t *samples, int num_samples)
{
    if (!resources->config->iq_correction.enable || !resources->iq_correction.fft_plan || num_samples <= 0) return;
    IqGpuDropin *d = (IqGpuDropin *)resources->iq_correction.fft_plan;
    pthread_mutex_lock(&resources->iq_correction.iq_factors_mutex);
    const IqCorrectionFactors f = resources->iq_correction.factors_buffer[resources->iq_correction.active_buffer_idx];
    pthread_mutex_unlock(&resources->iq_correction.iq_factors_mutex);
    iqgpu_chain *c = iqgpu_dropin_module(d, IQGPU_STAGE_IQ);
    size_t n_out = 0;
    uint32_t one = (uint32_t)num_samples;
    if (!c || iqgpu_chain_set_iq_factors(c, f.mag, f.phase) != IQGPU_OK ||
        iqgpu_chain_process(c, samples, (size_t)num_samples, &one, 1, samples, (size_t)num_samples * 8, &n_out, NULL) != IQGPU_OK)
        iqgpu_dropin_fatal(resources, "I/Q correction: GPU execution failed");
}

void iq_correct_run_optimization(AppResources *resources, const complex_float_t *optimization_data)
{
    if (!resources->config->iq_correction.enable || !resources->iq_correction.fft_plan) return;
    IqCorrectionResources *q = &resources->iq_correction;
    IqGpuDropin *d = (IqGpuDropin *)q->fft_plan;
    /* :157-162 paces the passes with the wall clock and :391 draws the step directions from rand() seeded with time(): the
     * result then depends on how fast the machine is and on when it was started.  Deliberate divergence:
     * the clock is the SAMPLE clock — frames that have entered the pre-processor stage / input rate — and the directions
     * come from a counter-based generator (seed: IQGPU_IQ_SEED), so a capture gives the same factors every time. */
    const double rate = (double)resources->source_info.samplerate;
    const double now = rate > 0.0 ? (double)__atomic_load_n(&d->frames_pre_total, __ATOMIC_RELAXED) / rate : 0.0;
    if ((now - q->last_optimization_time) * 1000.0 < IQ_CORRECTION_INTERVAL_MS) return;

    float dirs[2 * IQ_MAX_PASSES];
    const uint64_t attempt = d->iq_attempts++;
    /* IQGPU_IQ_LIBC_RAND=1: the reference's own source of directions (rand(), :391) — for parity runs against it */
    const char *libc_rand = getenv("IQGPU_IQ_LIBC_RAND");
    const int use_rand = libc_rand && *libc_rand && *libc_rand != '0';
    for (int i = 0; i < 2 * IQ_MAX_PASSES; i++)
        dirs[i] = use_rand ? ((rand() > (RAND_MAX / 2)) ? 1.0f : -1.0f) : iqgpu_iq_direction(d->iq_seed, attempt, (uint32_t)i);
    pthread_mutex_lock(&q->iq_factors_mutex);
    const int active = q->active_buffer_idx;
    float mag = q->factors_buffer[active].mag, phase = q->factors_buffer[active].phase;
    pthread_mutex_unlock(&q->iq_factors_mutex);
    float avg = 0.f, range = 0.f;
    if (iqgpu_iq_optimize((const float *)optimization_data, dirs, &mag, &phase, &avg, &range) != IQGPU_OK) {
        log_error("I/Q optimisation pass failed on the GPU: %s", iqgpu_last_error());
        return;
    }
    q->average_power = avg;
    q->power_range = range;
    if (range < IQ_CORRECTION_POWER_THRESHOLD_DB) return;   /* :168-171: too weak, factors untouched */
    q->last_optimization_time = now;
    pthread_mutex_lock(&q->iq_factors_mutex);                /* :206-216: publish into the inactive slot, swap */
    const int inactive = 1 - q->active_buffer_idx;
    q->factors_buffer[inactive].mag = mag;
    q->factors_buffer[inactive].phase = phase;
    q->active_buffer_idx = inactive;
    pthread_mutex_unlock(&q->iq_factors_mutex);
}

void iq_correct_destroy(AppResources *resources)
{
    if (resources->iq_correction.fft_plan) {
        pthread_mutex_destroy(&resources->iq_correction.iq_factors_mutex);
        resources->iq_correction.fft_plan = NULL;
        iqgpu_dropin_release(resources);
    }
}

/* The reference calls this from initialize_application BEFORE iq_correct_init has run
 * (setup.c:291 vs pipeline.c:140) and crashes on a NULL fft_buffer.  The drop-in
 * declines politely in that situation; when it IS initialised it calibrates on the first block
 * of the file through the eager pre-processor path (the block must not enter the stream). */
void iqgpu_dropin_pre_eager(AppResources *resources, SampleChunk *item);
bool iq_correct_run_initial_calibration(ModuleContext *ctx, SNDFILE *infile)
{
    AppResources *resources = ctx->resources;
    if (!infile) { log_warn("Cannot perform initial I/Q correction without a valid file handle."); return true; }
    if (!resources->iq_correction.fft_plan) {
        log_warn("Initial I/Q calibration requested before the I/Q corrector exists; skipping.");
        return true;
    }
    if (resources->source_info.frames < IQ_CORRECTION_FFT_SIZE) {
        log_warn("Input file is too short for I/Q calibration. Skipping.");
        return true;
    }
    const size_t raw_bytes = IQ_CORRECTION_FFT_SIZE * resources->input_bytes_per_sample_pair;
    void *raw = malloc(raw_bytes);
    complex_float_t *cf = (complex_float_t *)malloc(IQ_CORRECTION_FFT_SIZE * sizeof(complex_float_t));
    bool ok = true;
    if (!raw || !cf) { log_fatal("Failed to allocate temporary buffers for I/Q calibration."); ok = false; }
    else if (sf_read_raw(infile, raw, (sf_count_t)raw_bytes) < (sf_count_t)raw_bytes) {
        log_warn("Failed to read enough samples for I/Q calibration. Skipping.");
        sf_seek(infile, 0, SEEK_SET);
    } else {
        SampleChunk tmp;
        memset(&tmp, 0, sizeof(tmp));
        tmp.raw_input_data = raw;
        tmp.frames_read = IQ_CORRECTION_FFT_SIZE;
        tmp.packet_sample_format = resources->input_format;
        tmp.complex_sample_buffer_a = cf;
        tmp.complex_buffer_capacity_samples = IQ_CORRECTION_FFT_SIZE;
        iqgpu_dropin_pre_eager(resources, &tmp);
        pre_processor_reset(resources);                      /* the probe block must leave no state behind */
        resources->iq_correction.last_optimization_time = -1e9;
        iq_correct_run_optimization(resources, cf);
        resources->iq_correction.last_optimization_time = 0.0;          /* sample clock: the stream starts here */
        if (sf_seek(infile, 0, SEEK_SET) < 0) { log_fatal("Failed to rewind input file after I/Q calibration."); ok = false; }
        else log_info("Initial I/Q calibration complete.");
    }
    free(raw); free(cf);
    return ok;
}
