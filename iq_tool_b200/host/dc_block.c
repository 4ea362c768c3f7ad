/*
 * dc_block.c (GPU drop-in) — replaces reference src/dc_block.c (include/dc_block.h:29,39,51,60).
 * The DC blocker itself lives in the GPU chain (blocked affine scan inside K1 / the fused front);
 * this file keeps the reference's object protocol: dc_block_filter is non-NULL iff enabled.
 */
#include "dc_block.h"

#include <math.h>

#include "constants.h"
#include "iqgpu_dropin.h"
#include "log.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

bool dc_block_create(AppConfig *config, AppResources *resources)
{
    if (!config->dc_block.enable) {
        resources->dc_block.dc_block_filter = NULL;
        return true;
    }
    /* dc_block.c:32-39: alpha = 2 pi fc / Fs must be positive */
    const float alpha = (float)(2.0 * M_PI * DC_BLOCK_CUTOFF_HZ / resources->source_info.samplerate);
    if (alpha <= 0.0f) {
        log_fatal("DC Block: Calculated normalized alpha (%.6f) is invalid. Ensure DC_BLOCK_CUTOFF_HZ > 0.", alpha);
        return false;
    }
    IqGpuDropin *d = iqgpu_dropin_get(resources);
    if (!d) { log_fatal("Failed to create GPU DC block object."); return false; }
    iqgpu_dropin_addref(resources);
    resources->dc_block.dc_block_filter = d;
    log_info("DC Block enabled");
    return true;
}

void dc_block_reset(AppResources *resources)
{
    if (!resources->config->dc_block.enable || !resources->dc_block.dc_block_filter) return;
    IqGpuDropin *d = (IqGpuDropin *)resources->dc_block.dc_block_filter;
    if (d->mod_dc) iqgpu_chain_reset(d->mod_dc);      /* module-level state; the fused chain resets as a whole */
}

void dc_block_apply(AppResources *resources, complex_float_t *samples, int num_samples)
{
    if (!resources->config->dc_block.enable || !resources->dc_block.dc_block_filter || num_samples <= 0) return;
    IqGpuDropin *d = (IqGpuDropin *)resources->dc_block.dc_block_filter;
    iqgpu_chain *c = iqgpu_dropin_module(d, IQGPU_STAGE_DC);
    size_t n_out = 0;
    uint32_t one = (uint32_t)num_samples;
    if (!c || iqgpu_chain_process(c, samples, (size_t)num_samples, &one, 1, samples, (size_t)num_samples * 8, &n_out, NULL) != IQGPU_OK)
        iqgpu_dropin_fatal(resources, "DC block: GPU execution failed");
}

void dc_block_destroy(AppResources *resources)
{
    if (resources->dc_block.dc_block_filter) {
        resources->dc_block.dc_block_filter = NULL;
        iqgpu_dropin_release(resources);
    }
}
