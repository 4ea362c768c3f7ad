/*
 * pre_processor.c (GPU drop-in) — replaces reference src/pre_processor.c
 * (include/pre_processor.h:24,34).  Stage order of the reference (:10-55): convert -> DC block ->
 * I/Q apply -> pre-resample NCO -> pre-resample filter, all landing in complex_sample_buffer_a.
 *
 * FUSED mode: the chunk's raw input is staged for the GPU train and nothing else happens here;
 * the cf32 ping-pong buffers of the chunk are never written (they need not exist on the host at
 * all, SURVEY 8(b) "Ownership") except for the first 1024 frames when the I/Q optimiser is on,
 * because the pre thread copies those to the optimiser queue (src/pipeline.c:468-476).
 */
#include "pre_processor.h"

#include "constants.h"
#include "dc_block.h"
#include "filter.h"
#include "frequency_shift.h"
#include "iq_correct.h"
#include "iqgpu_dropin.h"
#include "log.h"
#include "sample_convert.h"
#include "signal_handler.h"

/* the reference's literal sequence on host buffers, every step through a module-level GPU chain */
void iqgpu_dropin_pre_eager(AppResources *resources, SampleChunk *item)
{
    AppConfig *config = (AppConfig *)resources->config;
    item->current_input_buffer = item->complex_sample_buffer_a;
    item->current_output_buffer = item->complex_sample_buffer_a;
    if (!convert_block_to_cf32(item->raw_input_data, item->current_output_buffer, (size_t)item->frames_read,
                               item->packet_sample_format, config->gain)) {
        handle_fatal_thread_error("Pre-Processor: Failed to convert samples.", resources);
        item->frames_read = 0;
        return;
    }
    if (config->dc_block.enable) dc_block_apply(resources, item->current_output_buffer, (int)item->frames_read);
    if (config->iq_correction.enable) iq_correct_apply(resources, item->current_output_buffer, (int)item->frames_read);
    if (resources->pre_resample_nco)
        freq_shift_apply(resources->pre_resample_nco, resources->nco_shift_hz, item->current_output_buffer,
                         item->current_output_buffer, (unsigned int)item->frames_read);
    if (resources->user_filter_object && !config->apply_user_filter_post_resample)
        item->frames_read = filter_apply(resources, item, false);
}

/* optimiser probe: convert + I/Q apply + pre-NCO of the chunk's first 1024 frames.  The DC
 * blocker is left out (its state belongs to the stream; the metric ignores bins within 5 % of DC,
 * iq_correct.c:345-346). */
static void iq_probe(IqGpuDropin *d, AppResources *resources, SampleChunk *item)
{
    const AppConfig *config = resources->config;
    if (!convert_block_to_cf32(item->raw_input_data, item->complex_sample_buffer_a, IQ_CORRECTION_FFT_SIZE,
                               item->packet_sample_format, config->gain))
        return;
    iq_correct_apply(resources, item->complex_sample_buffer_a, IQ_CORRECTION_FFT_SIZE);
    if (resources->pre_resample_nco)
        freq_shift_apply(resources->pre_resample_nco, resources->nco_shift_hz, item->complex_sample_buffer_a,
                         item->complex_sample_buffer_a, IQ_CORRECTION_FFT_SIZE);
    (void)d;
}

void pre_processor_apply_chain(AppResources *resources, SampleChunk *item)
{
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) {
        handle_fatal_thread_error("Pre-Processor: no GPU context.", resources);
        item->frames_read = 0;
        return;
    }
    if (item->frames_read > 0) __atomic_fetch_add(&d->frames_pre_total, (uint64_t)item->frames_read, __ATOMIC_RELAXED);   /* the sample clock */
    if (d->eager) { iqgpu_dropin_pre_eager(resources, item); return; }
    item->current_input_buffer = item->complex_sample_buffer_a;
    item->current_output_buffer = item->complex_sample_buffer_a;
    const int64_t staged_frames = item->frames_read;
    if (!iqgpu_dropin_stage_chunk(d, item)) {
        iqgpu_dropin_fatal(resources, "Pre-Processor: Failed to stage samples for the GPU chain.");
        item->frames_read = 0;
        return;
    }
    if (resources->config->iq_correction.enable && staged_frames >= IQ_CORRECTION_FFT_SIZE && item->complex_sample_buffer_a)
        iq_probe(d, resources, item);
}

void pre_processor_reset(AppResources *resources)
{
    dc_block_reset(resources);
    freq_shift_reset_nco(resources->pre_resample_nco);
    filter_reset(resources);
    IqGpuDropin *d = iqgpu_dropin_find(resources);
    if (d) iqgpu_dropin_mark_reset(d);      /* FUSED: the whole GPU chain restarts at this point of the stream */
}
