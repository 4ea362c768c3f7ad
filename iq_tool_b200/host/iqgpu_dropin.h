/*
 * iqgpu_dropin.h — shared state of the drop-in host layer (the .c files of iq_tool_b200/host/).
 *
 * These translation units REPLACE the reference's src/{pre_processor,resampler,post_processor,
 * filter,frequency_shift,sample_convert,dc_block,iq_correct,agc}.c inside an iq_tool build: same
 * prototypes (the reference's own headers are included, never copied), same error behaviour,
 * but every sample is processed by libiqgpu.so (include/iqgpu.h).  See INTEGRATION.md.
 *
 * Two execution modes, same results:
 *   FUSED (default)  The three stage entry points cooperate.  pre_processor_apply_chain only
 *                    STAGES the chunk's raw input (pinned memory); resampler_execute returns the
 *                    chunk's frame count in closed form; post_processor_apply_chain EXECUTES
 *                    every chunk staged so far as one train through the fused GPU chain
 *                    (iqgpu_chain_process) and hands the chunk its own slice of the output.
 *                    Under load the reference's queues fill up and trains grow to hundreds of
 *                    chunks per launch; an idle live stream degenerates to one chunk per call.
 *   EAGER            (IQGPU_DROPIN_EAGER=1) every module function does its own work on the host
 *                    buffers it is given, through a module-level chain (stage_select): the literal
 *                    per-stage drop-in.  Slow (PCIe round trip per stage) but exact, and what a
 *                    direct caller of dc_block_apply()/freq_shift_apply()/... gets in either mode.
 */
#ifndef IQGPU_DROPIN_H
#define IQGPU_DROPIN_H

#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>

#include "app_context.h"      /* reference: AppConfig, AppResources */
#include "pipeline_types.h"   /* reference: SampleChunk */

#include "iqgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IQGPU_DROPIN_MAX_PENDING 2048   /* chunks staged and not yet executed (reference pool: 512) */

typedef struct IqGpuPending {
    uint32_t frames;          /* raw frames staged for this chunk */
    uint8_t  reset_before;    /* a stream discontinuity precedes this chunk */
    uint8_t  reaches_post;    /* the stage threads will call post_processor_apply_chain for it */
} IqGpuPending;

typedef struct IqGpuStageBuf {
    unsigned char *raw;       /* pinned */
    size_t         cap_bytes, bytes;
    IqGpuPending   pend[IQGPU_DROPIN_MAX_PENDING];
    size_t         n;
} IqGpuStageBuf;

typedef struct IqGpuDropin {
    AppResources  *res;
    int            refs;
    int            eager;
    int            device;
    pthread_mutex_t mu;
    pthread_cond_t  room;

    /* configuration snapshot (built on first use, after every *_create ran) */
    bool               cfg_ready;
    iqgpu_chain_config cfg;
    iqgpu_chain       *plan;        /* plan-only chain: design introspection + closed forms */
    iqgpu_chain       *fused;       /* the whole chain on the GPU */
    iqgpu_chain       *mod_dc, *mod_iq, *mod_nco, *mod_rs, *mod_filter, *mod_agc;   /* module-level */

    /* FUSED mode */
    IqGpuStageBuf  stage[2];
    int            fill;            /* buffer the pre thread appends to */
    bool           reset_pending;   /* next staged chunk starts a new stream */
    uint64_t       pre_fft_rem;     /* pre-resample FFT remainder (closed form) */
    uint64_t       rs_pos;          /* frames handed to resampler_execute since its reset */
    /* executed train, waiting to be handed out chunk by chunk */
    unsigned char *out;             /* pinned */
    size_t         out_cap_bytes;
    uint32_t       out_counts[IQGPU_DROPIN_MAX_PENDING];
    uint8_t        out_reaches_post[IQGPU_DROPIN_MAX_PENDING];
    size_t         out_n, out_next;
    size_t         out_off_bytes;
    float          probe[2 * 1024]; /* I/Q optimiser probe block */
    /* I/Q optimiser pacing and directions (sample clock and a seeded generator instead of the reference's
     * wall clock and rand()): frames that have entered the pre stage, probed blocks so far, generator seed */
    volatile uint64_t frames_pre_total;
    uint64_t       iq_attempts;
    uint32_t       iq_seed;
} IqGpuDropin;

IqGpuDropin *iqgpu_dropin_get(AppResources *res);         /* creates on first use, never NULL unless OOM */
IqGpuDropin *iqgpu_dropin_find(const AppResources *res);   /* NULL if none */
void         iqgpu_dropin_addref(AppResources *res);
void         iqgpu_dropin_release(AppResources *res);      /* frees everything at refcount 0 */
bool         iqgpu_dropin_configure(IqGpuDropin *d);       /* builds cfg + plan; false on invalid config */
iqgpu_chain *iqgpu_dropin_module(IqGpuDropin *d, int stage);   /* module-level chain (lazy) */
void         iqgpu_dropin_fatal(AppResources *res, const char *what);

/* FUSED engine */
bool     iqgpu_dropin_stage_chunk(IqGpuDropin *d, SampleChunk *item);
unsigned iqgpu_dropin_resampler_count(IqGpuDropin *d, unsigned n_in);
bool     iqgpu_dropin_finish_chunk(IqGpuDropin *d, SampleChunk *item);
void     iqgpu_dropin_mark_reset(IqGpuDropin *d);

#ifdef __cplusplus
}
#endif
#endif
