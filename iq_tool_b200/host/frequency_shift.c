/*
 * frequency_shift.c (GPU drop-in) — replaces reference src/frequency_shift.c
 * (include/frequency_shift.h:31,42,52,58).  The mixer is liquid's LIQUID_NCO restated on the GPU
 * (32-bit phase accumulator, 1024-entry sine table, nearest entry) inside K1 / K5.
 * The NCO "objects" stored in AppResources are small handles naming the owning context.
 */
#include "frequency_shift.h"

#include <math.h>
#include <stdlib.h>

#include "constants.h"
#include "iqgpu_dropin.h"
#include "log.h"

typedef struct { IqGpuDropin *d; int is_post; } NcoHandle;

bool freq_shift_create(AppConfig *config, AppResources *resources)
{
    if (!config || !resources) return false;
    resources->pre_resample_nco = NULL;
    resources->post_resample_nco = NULL;
    /* frequency_shift.c:33-46: resolve the shift, validate dependent options */
    if (resources->nco_shift_hz == 0.0 && config->freq_shift_hz_arg != 0.0f)
        resources->nco_shift_hz = (double)config->freq_shift_hz_arg;
    if (config->shift_after_resample && fabs(resources->nco_shift_hz) < 1e-9) {
        log_fatal("Option --shift-after-resample was used, but no effective frequency shift was requested or calculated.");
        return false;
    }
    if (fabs(resources->nco_shift_hz) < 1e-9) return true;
    const double rate = config->shift_after_resample ? config->target_rate : (double)resources->source_info.samplerate;
    if (fabs(resources->nco_shift_hz) > (SHIFT_FACTOR_LIMIT * rate)) {
        log_error("Requested frequency shift %.2f Hz exceeds sanity limit for the %s-resample rate of %.1f Hz.",
                  resources->nco_shift_hz, config->shift_after_resample ? "post" : "pre", rate);
        return false;
    }
    NcoHandle *h = (NcoHandle *)calloc(1, sizeof(*h));
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!h || !d) { free(h); log_error("Failed to create GPU NCO (frequency shifter)."); return false; }
    iqgpu_dropin_addref(resources);
    h->d = d; h->is_post = config->shift_after_resample != 0;
    if (h->is_post) resources->post_resample_nco = h; else resources->pre_resample_nco = h;
    return true;
}

void freq_shift_apply(void *nco, double shift_hz, complex_float_t *input_buffer, complex_float_t *output_buffer,
                      unsigned int num_frames)
{
    (void)shift_hz;   /* the sign is part of the chain design (mix up for >= 0, down otherwise, :91-95) */
    if (!nco || num_frames == 0) return;
    NcoHandle *h = (NcoHandle *)nco;
    iqgpu_chain *c = iqgpu_dropin_module(h->d, IQGPU_STAGE_NCO);
    size_t n_out = 0;
    uint32_t one = num_frames;
    if (!c || iqgpu_chain_process(c, input_buffer, num_frames, &one, 1, output_buffer, (size_t)num_frames * 8, &n_out, NULL) != IQGPU_OK)
        iqgpu_dropin_fatal(h->d->res, "Frequency shift: GPU execution failed");
}

void freq_shift_reset_nco(void *nco)
{
    if (!nco) return;
    NcoHandle *h = (NcoHandle *)nco;
    if (h->d->mod_nco) iqgpu_chain_reset(h->d->mod_nco);   /* phase <- 0, frequency kept (:105) */
}

void freq_shift_destroy_ncos(AppResources *resources)
{
    if (!resources) return;
    void **slots[2] = {&resources->pre_resample_nco, &resources->post_resample_nco};
    for (int i = 0; i < 2; i++) {
        if (*slots[i]) {
            free(*slots[i]);
            *slots[i] = NULL;
            iqgpu_dropin_release(resources);
        }
    }
}
