/*
 * filter.c (GPU drop-in) — replaces reference src/filter.c (include/filter.h:36,47,54,69).
 * Tap design (Kaiser stages, spectral inversion, LUT-NCO band-pass modulation, convolution of up
 * to five stages, peak/DC normalisation, FIR-vs-FFT choice, pre-vs-post placement,
 * filter.c:43-393) is restated on the host inside libiqgpu (csrc/design.cpp); the filters
 * themselves are K3 (tiled FIR) and K4 (hand-written FFT block filter).
 */
#include "filter.h"

#include "iqgpu_dropin.h"
#include "log.h"

bool filter_create(AppConfig *config, AppResources *resources, MemoryArena *arena)
{
    (void)arena;   /* design scratch lives inside libiqgpu; nothing is carved from the setup arena */
    resources->user_filter_object = NULL;
    resources->user_filter_type_actual = FILTER_IMPL_NONE;
    resources->user_filter_block_size = 0;
    resources->pre_fft_remainder_buffer = resources->post_fft_remainder_buffer = NULL;
    resources->pre_fft_remainder_len = resources->post_fft_remainder_len = 0;
    config->apply_user_filter_post_resample = false;
    if (config->num_filter_requests == 0) return true;

    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) { log_fatal("Failed to create final combined filter object."); return false; }
    iqgpu_dropin_addref(resources);
    log_info("Designing filter coefficients (this may be slow for large filters)...");
    if (!iqgpu_dropin_configure(d)) {      /* logs the reason (incompatible with the output rate, fft size too small ...) */
        iqgpu_dropin_release(resources);
        return false;
    }
    iqgpu_chain_info inf;
    iqgpu_chain_get_info(d->plan, &inf);
    resources->user_filter_type_actual = (FilterImplementationType)inf.filter_impl;
    resources->user_filter_block_size = inf.filter_block_size;
    config->apply_user_filter_post_resample = inf.filter_post_resample != 0;
    resources->user_filter_object = d;
    return true;
}

void filter_reset(AppResources *resources)
{
    if (!resources->user_filter_object) return;
    IqGpuDropin *d = (IqGpuDropin *)resources->user_filter_object;
    if (d->mod_filter) iqgpu_chain_reset(d->mod_filter);
}

void filter_destroy(AppResources *resources)
{
    if (resources->user_filter_object) {
        resources->user_filter_object = NULL;
        iqgpu_dropin_release(resources);
    }
}

unsigned int filter_apply(AppResources *resources, SampleChunk *item, bool is_post_resample)
{
    if (!resources->user_filter_object)
        return is_post_resample ? item->frames_to_write : (unsigned int)item->frames_read;
    const unsigned int frames_in = is_post_resample ? item->frames_to_write : (unsigned int)item->frames_read;
    if (frames_in == 0) return 0;
    IqGpuDropin *d = (IqGpuDropin *)resources->user_filter_object;
    iqgpu_chain *c = iqgpu_dropin_module(d, IQGPU_STAGE_FILTER);
    const bool fft = resources->user_filter_type_actual == FILTER_IMPL_FFT_SYMMETRIC ||
                     resources->user_filter_type_actual == FILTER_IMPL_FFT_ASYMMETRIC;
    /* FIR: in place on current_input_buffer (filter.c:449-462); FFT: input -> current_output_buffer (:464-483) */
    complex_float_t *dst = fft ? item->current_output_buffer : item->current_input_buffer;
    size_t n_out = 0;
    uint32_t one = frames_in;
    const size_t cap = item->complex_buffer_capacity_samples * sizeof(complex_float_t);
    if (!c || iqgpu_chain_process(c, item->current_input_buffer, frames_in, &one, 1, dst, cap, &n_out, NULL) != IQGPU_OK) {
        iqgpu_dropin_fatal(resources, "Filter: GPU execution failed");
        return 0;
    }
    return (unsigned int)n_out;
}
