/* output_wav_common.c — drop-in replacement for the reference's src/output_wav_common.c (the implementation
 * shared by its WAV and RF64 output modules, include/output_wav_common.h:20-24): same five functions, same
 * behaviour towards the Writer thread and the ring buffer, same refusals and texts — with the container written
 * by libiqgpu's host-only header code (iqgpu_wav_build_header through host/sndfile_min.c) instead of libsndfile.
 * The reference's own wrappers src/output_wav.c / src/output_wav_rf64.c compile unchanged on top of it (they only
 * pass SF_FORMAT_WAV / SF_FORMAT_RF64 down; host/compat/sndfile.h supplies the constants).  SURVEY.md 8(f) rank 4. */
#include "output_wav_common.h"

#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include "app_context.h"
#include "constants.h"
#include "log.h"
#include "memory_arena.h"
#include "ring_buffer.h"
#include "signal_handler.h"
#include "utils.h"

#include "iqgpu.h"
#include "sndfile_min.h"

static WavCommonData *state_of(ModuleContext *ctx) { return (WavCommonData *)ctx->resources->output_module_private_data; }

/* the y/n question of src/output_wav_common.c:27-41 */
static bool overwrite_confirmed(const char *shown_path)
{
    fprintf(stderr, "\nOutput file %s exists.\nOverwrite? (y/n): ", shown_path);
    const int raw = getchar();
    const bool line_pending = raw != '\n' && raw != EOF;
    if (line_pending) clear_stdin_buffer();
    if (tolower(raw) == 'y') return true;
    if (line_pending) log_debug("Operation cancelled by user.");
    return false;
}

/* :46-52 */
bool wav_common_validate_options(AppConfig *config)
{
    if (config->output_format == CS16 || config->output_format == CU8) return true;
    log_fatal("Invalid sample format '%s' for WAV/RF64 container. Only 'cs16' and 'cu8' are supported.", config->output_sample_format_name);
    return false;
}

/* :54-118 */
bool wav_common_initialize(ModuleContext *ctx, int sf_format_flag)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    WavCommonData *data = (WavCommonData *)mem_arena_alloc(&resources->setup_arena, sizeof(WavCommonData), true);
    if (!data) return false;
    resources->output_module_private_data = data;

    const char *path = config->effective_output_filename;
    struct stat sb;
    if (lstat(path, &sb) == 0) {
        if (!S_ISREG(sb.st_mode)) { log_fatal("Output path '%s' exists but is not a regular file. Aborting.", path); return false; }
        if (!overwrite_confirmed(path)) return false;
    }
    if (config->output_format != CS16 && config->output_format != CU8) return false;     /* validation should have caught it */
    const int container = (sf_format_flag & SF_FORMAT_TYPEMASK) == SF_FORMAT_RF64 ? IQGPU_CONTAINER_RF64 : IQGPU_CONTAINER_WAV;
    const int rate = (int)config->target_rate;
    unsigned char probe[80];
    if (iqgpu_wav_build_header(container, (int)config->output_format, rate, 0, probe, sizeof(probe)) != IQGPU_OK) {
        log_fatal("The requested container format is not supported (Rate: %d, Format: 0x%08X).", rate, sf_format_flag);
        return false;
    }
    data->handle = sfmin_create(path, container, (int)config->output_format, rate);
    if (!data->handle) { log_fatal("Error opening output WAV file %s", path); return false; }
    return true;
}

/* :120-157 — the Writer thread: ring buffer -> file in IO_OUTPUT_WRITER_CHUNK_SIZE pieces */
void *wav_common_run_writer(ModuleContext *ctx)
{
    AppResources *resources = ctx->resources;
    WavCommonData *data = state_of(ctx);
    unsigned char *staging = (unsigned char *)resources->writer_local_buffer;
    if (!staging) { handle_fatal_thread_error("WAV writer: Local buffer is NULL.", resources); return NULL; }

    for (;;) {
        const size_t n = ring_buffer_read(resources->writer_input_buffer, staging, IO_OUTPUT_WRITER_CHUNK_SIZE);
        if (n == 0) break;                                  /* end of stream or shutdown */
        const sf_count_t done = sfmin_write_raw(data->handle, staging, (sf_count_t)n);
        if (done > 0) data->total_bytes_written += done;
        if ((size_t)done != n) {
            handle_fatal_thread_error("WAV writer: File write error.", resources);
            break;
        }
        if (resources->progress_callback) {
            const unsigned long long frames = (unsigned long long)data->total_bytes_written / resources->output_bytes_per_sample_pair;
            pthread_mutex_lock(&resources->progress_mutex);
            resources->total_output_frames = frames;
            pthread_mutex_unlock(&resources->progress_mutex);
            resources->progress_callback(frames, resources->expected_total_output_frames, (unsigned long long)data->total_bytes_written,
                                         resources->progress_callback_udata);
        }
    }
    log_debug("Common WAV writer thread is exiting.");
    return NULL;
}

/* :159-166 */
size_t wav_common_write_chunk(ModuleContext *ctx, const void *buffer, size_t bytes_to_write)
{
    WavCommonData *data = state_of(ctx);
    if (!data || !data->handle || bytes_to_write == 0) return 0;
    const sf_count_t done = sfmin_write_raw(data->handle, buffer, (sf_count_t)bytes_to_write);
    if (done > 0) data->total_bytes_written += done;
    return (size_t)done;
}

/* :168-174 — closing is what puts the sizes into the header */
void wav_common_finalize_output(ModuleContext *ctx)
{
    WavCommonData *data = state_of(ctx);
    if (!data) return;
    if (data->handle) {
        if (sfmin_finish(data->handle) != 0) log_warn("Could not finalise the WAV header of the output file.");
        data->handle = NULL;
    }
    ctx->resources->final_output_size_bytes = data->total_bytes_written;
}
