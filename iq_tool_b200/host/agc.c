/*
 * agc.c (GPU drop-in) — replaces reference src/agc.c (include/agc.h:31,45,54,61).
 * DX/LOCAL: liquid's agc_crcf recurrence; DIGITAL: the reference's own peak/lock/ratchet state
 * machine (agc.c:105-222) as per-chunk peaks -> one-warp scan -> scale, all inside K5.
 * Wall-clock reads of the reference (agc.c:176,202-207) are replaced by the sample clock.
 */
#include "agc.h"

#include "iqgpu_dropin.h"
#include "log.h"

bool agc_create(AppConfig *config, AppResources *resources)
{
    resources->output_agc_object = NULL;
    if (!config->output_agc.enable) return true;
    resources->agc_is_locked = false;
    resources->agc_current_gain = 1.0f;
    resources->agc_samples_seen = 0;
    resources->agc_last_strong_peak_time = 0.0;
    resources->agc_peak_memory = (config->output_agc.profile != AGC_PROFILE_DIGITAL) ? 0.001f : 0.05f;   /* :66,78 */
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) { log_fatal("Failed to create GPU AGC object."); return false; }
    iqgpu_dropin_addref(resources);
    /* the reference keeps a liquid object only for DX/LOCAL; the slot doubles as "we hold a reference" */
    resources->output_agc_object = d;
    log_info("Output AGC enabled.");
    return true;
}

void agc_apply(AppResources *resources, complex_float_t *samples, unsigned int num_samples)
{
    if (!resources->config->output_agc.enable || num_samples == 0 || !resources->output_agc_object) return;
    IqGpuDropin *d = (IqGpuDropin *)resources->output_agc_object;
    iqgpu_chain *c = iqgpu_dropin_module(d, IQGPU_STAGE_AGC);
    size_t n_out = 0;
    uint32_t one = num_samples;
    if (!c || iqgpu_chain_process(c, samples, num_samples, &one, 1, samples, (size_t)num_samples * 8, &n_out, NULL) != IQGPU_OK) {
        iqgpu_dropin_fatal(resources, "AGC: GPU execution failed");
        return;
    }
    iqgpu_chain_info inf;
    if (iqgpu_chain_get_info(c, &inf) == IQGPU_OK) {
        resources->agc_is_locked = inf.agc_locked != 0;
        resources->agc_current_gain = inf.agc_gain;
        resources->agc_peak_memory = inf.agc_peak_memory;
        resources->agc_samples_seen = inf.agc_samples_seen;
    }
}

void agc_reset(AppResources *resources)
{
    if (resources->output_agc_object) {
        IqGpuDropin *d = (IqGpuDropin *)resources->output_agc_object;
        if (d->mod_agc) iqgpu_chain_reset(d->mod_agc);
    }
    resources->agc_is_locked = false;            /* agc.c:232-236 */
    resources->agc_samples_seen = 0;
    resources->agc_peak_memory = 0.05f;
    resources->agc_current_gain = 1.0f;
    resources->agc_last_strong_peak_time = 0.0;
}

void agc_destroy(AppResources *resources)
{
    if (resources->output_agc_object) {
        resources->output_agc_object = NULL;
        iqgpu_dropin_release(resources);
    }
}
