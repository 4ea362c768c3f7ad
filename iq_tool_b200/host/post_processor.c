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

/* EAGER mode: one module-level GPU chain per stage on the chunk's host buffers.  `at` follows the samples through the
 * ping-pong pair: a stage that works out of place leaves them in the other buffer. */
static complex_float_t *other_of(const SampleChunk *item, const complex_float_t *buf)
{
    return buf == item->complex_sample_buffer_a ? item->complex_sample_buffer_b : item->complex_sample_buffer_a;
}

static void post_eager(AppResources *resources, SampleChunk *item)
{
    const AppConfig *config = resources->config;
    if (!item->frames_to_write) return;
    complex_float_t *at = item->current_input_buffer;
    const bool filter_here = resources->user_filter_object != NULL && config->apply_user_filter_post_resample;
    if (filter_here) {
        const FilterImplementationType impl = resources->user_filter_type_actual;
        item->frames_to_write = filter_apply(resources, item, true);
        /* the FFT forms write whole blocks into the chunk's output buffer, the FIR forms work in place */
        if (impl == FILTER_IMPL_FFT_SYMMETRIC || impl == FILTER_IMPL_FFT_ASYMMETRIC) at = item->current_output_buffer;
    }
    if (resources->post_resample_nco) {
        complex_float_t *mixed = other_of(item, at);
        freq_shift_apply(resources->post_resample_nco, resources->nco_shift_hz, at, mixed, item->frames_to_write);
        at = mixed;
    }
    agc_apply(resources, at, item->frames_to_write);
    if (convert_cf32_to_block(at, item->final_output_data, item->frames_to_write, config->output_format)) return;
    handle_fatal_thread_error("Post-Processor: Failed to convert samples.", resources);
    item->frames_to_write = 0;
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
