/* output_wav_common.c — drop-in replacement for the reference's src/output_wav_common.c (the implementation
 * shared by its WAV and RF64 output modules, include/output_wav_common.h:20-24): the same five functions with the
 * same behaviour towards the Writer thread and the ring buffer, the same refusals and texts — with the container
 * written by libiqgpu's host-only header code instead of libsndfile.  The reference's own wrappers src/output_wav.c /
 * src/output_wav_rf64.c compile unchanged on top of it (they only pass SF_FORMAT_WAV / SF_FORMAT_RF64 down;
 * host/compat/sndfile.h supplies the constants).  The sink object itself is in file_modules.c.  SURVEY.md 8(f) rank 4. */
#include "output_wav_common.h"

#include "file_modules.h"

bool   wav_common_validate_options(struct AppConfig *config) { return iqsink_format_allowed(config); }
bool   wav_common_initialize(ModuleContext *ctx, int sf_format_flag) { return iqsink_open(ctx, sf_format_flag); }
void  *wav_common_run_writer(ModuleContext *ctx) { return iqsink_drain_ring(ctx); }
size_t wav_common_write_chunk(ModuleContext *ctx, const void *buffer, size_t bytes_to_write) { return iqsink_put(ctx, buffer, bytes_to_write); }
void   wav_common_finalize_output(ModuleContext *ctx) { iqsink_close(ctx); }
