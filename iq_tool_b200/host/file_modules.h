/* file_modules.h — the host-side file endpoints behind the drop-in file modules (input_wav.c, input_rawfile.c,
 * output_wav_common.c): one "capture" object for both input modules and one "sink" object for the WAV / RF64
 * writers, on libiqgpu's host-only container code (iqgpu_wav_probe / iqgpu_wav_build_header) and plain FILEs
 * (sndfile_min.h).  The module files themselves only hold their option tables and v-tables. */
#ifndef IQGPU_FILE_MODULES_H
#define IQGPU_FILE_MODULES_H
#include <stdbool.h>
#include <stddef.h>

#include "module.h"
#include "iqgpu.h"
#include "sndfile_min.h"

/* ---- captures (input side) ---- */
typedef struct {
    SNDFILE       *handle;         /* FILE on the sample bytes */
    iqgpu_wav_info wav;            /* header + SDR metadata (WAV captures only) */
    const char    *kind;           /* "WAV" / "RAW": log texts */
    const char    *raw_format;     /* --raw-file-input-sample-format as given */
    double         raw_rate_hz;    /* --raw-file-input-rate */
} IqCapture;

/* allocate the module state in the setup arena and open the capture; false where the reference aborts setup */
bool  iqcap_open_wav(ModuleContext *ctx, float center_target_hz_arg);
bool  iqcap_open_raw(ModuleContext *ctx, const char *format_name, double rate_hz);
void *iqcap_stream(ModuleContext *ctx);                              /* the Reader thread body */
void  iqcap_close(ModuleContext *ctx);
void  iqcap_describe(const ModuleContext *ctx, InputSummaryInfo *info);
bool  iqcap_calibrate_before_streaming(ModuleContext *ctx);

/* ---- WAV / RF64 sink (output side) ---- */
bool   iqsink_format_allowed(struct AppConfig *config);
bool   iqsink_open(ModuleContext *ctx, int sf_format_flag);
void  *iqsink_drain_ring(ModuleContext *ctx);                        /* the Writer thread body */
size_t iqsink_put(ModuleContext *ctx, const void *bytes, size_t count);
void   iqsink_close(ModuleContext *ctx);
#endif
