/*
 * post_processor.c (GPU drop-in) — replaces reference src/post_processor.c
 * (include/post_processor.h:24,34).  Stage order of the reference (:9-70): post-resample filter ->
 * post-resample NCO (to the other ping-pong buffer) -> AGC -> convert to the output format.
 *
 * FUSED mode: this is where the GPU runs.  Every chunk staged so far goes through the fused chain
 * as one train; the chunk receives its own output bytes and its own frames_to_write (which, with
 * an FFT filter, is the block-quantised count the reference's filter_apply would return).
 */
#include "post_processor.h"

#include "agc.h"
#include "filter.h"
#include "frequency_shift.h"
#include "iqgpu_dropin.h"
#include "log.h"
#include "sample_convert.h"
#include "signal_handler.h"

static void post_eager(AppResources *resources, SampleChunk *item)
{
    AppConfig *config = (AppConfig *)resources->config;
    if (item->frames_to_write == 0) return;
    complex_float_t *cur = item->current_input_buffer;
    if (resources->user_filter_object && config->apply_user_filter_post_resample) {
        const bool fft = resources->user_filter_type_actual == FILTER_IMPL_FFT_SYMMETRIC ||
                         resources->user_filter_type_actual == FILTER_IMPL_FFT_ASYMMETRIC;
        item->frames_to_write = filter_apply(resources, item, true);
        if (fft) cur = item->current_output_buffer;
    }
    if (resources->post_resample_nco) {
        complex_float_t *dst = (cur == item->complex_sample_buffer_a) ? item->complex_sample_buffer_b : item->complex_sample_buffer_a;
        freq_shift_apply(resources->post_resample_nco, resources->nco_shift_hz, cur, dst, item->frames_to_write);
        cur = dst;
    }
    agc_apply(resources, cur, item->frames_to_write);
    if (!convert_cf32_to_block(cur, item->final_output_data, item->frames_to_write, config->output_format)) {
        handle_fatal_thread_error("Post-Processor: Failed to convert samples.", resources);
        item->frames_to_write = 0;
    }
}

void post_processor_apply_chain(AppResources *resources, SampleChunk *item)
{
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) {
        handle_fatal_thread_error("Post-Processor: no GPU context.", resources);
        item->frames_to_write = 0;
        return;
    }
    if (d->eager) { post_eager(resources, item); return; }
    if (!iqgpu_dropin_finish_chunk(d, item)) {
        iqgpu_dropin_fatal(resources, "Post-Processor: GPU chain execution failed.");
        item->frames_to_write = 0;
    }
}

void post_processor_reset(AppResources *resources)
{
    freq_shift_reset_nco(resources->post_resample_nco);
    filter_reset(resources);
    agc_reset(resources);
    /* FUSED: nothing more to do -- the marker the pre stage recorded restarts the whole chain
     * before the first chunk that follows the discontinuity. */
}
