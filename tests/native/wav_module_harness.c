/* wav_module_harness.c — TEST DRIVER for iq_tool_b200/host/input_wav.c (the drop-in WAV input module).
 * Plays the part of the reference's pipeline around an InputModuleInterface (src/pipeline.c: the reader thread
 * calls start_stream, the pre-processor thread dequeues reader_output_queue and recycles chunks through
 * free_sample_chunk_queue): initialize -> start_stream on its own thread -> drain -> summary -> cleanup.
 * The reference's own queue / arena / log / utils / signal sources are compiled in place from /root/reference. */
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>

#include "app_context.h"
#include "argparse.h"
#include "input_rawfile.h"
#include "input_wav.h"
#include "iq_correct.h"
#include "memory_arena.h"
#include "module.h"
#include "output_wav.h"
#include "output_wav_rf64.h"
#include "ring_buffer.h"
#include "pipeline_types.h"
#include "queue.h"
#include "signal_handler.h"

pthread_mutex_t g_console_mutex = PTHREAD_MUTEX_INITIALIZER;     /* src/main.c:70 */

typedef struct {
    int      initialized;
    int      input_format;
    int      samplerate;
    int64_t  source_frames;
    double   nco_shift_hz;
    uint64_t total_frames_read;
    uint64_t chunks;               /* non-final chunks delivered */
    uint64_t bytes;                /* raw bytes delivered */
    uint64_t largest_chunk_frames;
    int      saw_last_chunk;
    int      bytes_per_pair;
    int      has_known_length;
    int      summary_count;
    char     summary_label[16][64];
    char     summary_value[16][128];
} wavmod_result;

/* stand-in for the calibration service (src/iq_correct.c:237-300): reads the first block through the module's handle
 * and rewinds, like the service does; what it read is handed back for inspection */
static unsigned char g_calibration_block[1024 * 8];
static long g_calibration_bytes = -2;
bool iq_correct_run_initial_calibration(ModuleContext *ctx, SNDFILE *infile)
{
    const size_t want = 1024 * ctx->resources->input_bytes_per_sample_pair;
    g_calibration_bytes = (long)sf_read_raw(infile, g_calibration_block, (sf_count_t)want);
    return sf_seek(infile, 0, SEEK_SET) == 0;
}
long wavmod_calibration_block(unsigned char *dst, size_t capacity)
{
    if (g_calibration_bytes > 0) memcpy(dst, g_calibration_block, (size_t)g_calibration_bytes < capacity ? (size_t)g_calibration_bytes : capacity);
    return g_calibration_bytes;
}

static InputModuleInterface *g_api;     /* the module under test */

static void *reader_main(void *arg)
{
    ModuleContext *ctx = (ModuleContext *)arg;
    return g_api->start_stream(ctx);
}

static int drive_input(const char *path, float freq_shift_hz_arg, int iq_correction, unsigned pool_chunks, unsigned chunk_frames,
                       unsigned char *sink, size_t sink_capacity, wavmod_result *out)
{
    memset(out, 0, sizeof(*out));
    reset_shutdown_flag();
    AppConfig *config = (AppConfig *)calloc(1, sizeof(AppConfig));
    AppResources *resources = (AppResources *)calloc(1, sizeof(AppResources));
    if (!config || !resources || !mem_arena_init(&resources->setup_arena, 4u << 20)) return -1;
    config->input_filename_arg = (char *)path;
    config->effective_input_filename = (char *)path;
    config->freq_shift_hz_arg = freq_shift_hz_arg;
    config->iq_correction.enable = iq_correction != 0;
    g_calibration_bytes = -2;
    pthread_mutex_init(&resources->progress_mutex, NULL);

    ModuleContext ctx = {config, resources};
    InputModuleInterface *api = g_api;
    out->has_known_length = api->has_known_length() ? 1 : 0;
    int rc = 0;
    if (api->validate_options && !api->validate_options(config)) { rc = 7; goto done; }
    if (!api->initialize(&ctx)) { rc = 1; goto done; }
    out->initialized = 1;
    out->input_format = (int)resources->input_format;
    out->samplerate = resources->source_info.samplerate;
    out->source_frames = resources->source_info.frames;
    out->nco_shift_hz = resources->nco_shift_hz;
    out->bytes_per_pair = (int)resources->input_bytes_per_sample_pair;
    if (!api->pre_stream_iq_correction(&ctx)) { rc = 6; goto done; }

    Queue reader_out, free_q;
    if (!queue_init(&reader_out, pool_chunks + 1, &resources->setup_arena) || !queue_init(&free_q, pool_chunks + 1, &resources->setup_arena)) { rc = -1; goto done; }
    resources->reader_output_queue = &reader_out;
    resources->free_sample_chunk_queue = &free_q;
    SampleChunk *pool = (SampleChunk *)calloc(pool_chunks, sizeof(SampleChunk));
    for (unsigned i = 0; i < pool_chunks; i++) {
        pool[i].raw_input_capacity_bytes = (size_t)chunk_frames * resources->input_bytes_per_sample_pair;
        pool[i].raw_input_data = malloc(pool[i].raw_input_capacity_bytes);
        pool[i].stream_discontinuity_event = true;      /* the module must clear it */
        queue_enqueue(&free_q, &pool[i]);
    }

    pthread_t reader;
    pthread_create(&reader, NULL, reader_main, &ctx);
    for (;;) {
        SampleChunk *c = (SampleChunk *)queue_dequeue(&reader_out);
        if (!c) { rc = 2; break; }
        if (c->stream_discontinuity_event || c->packet_sample_format != resources->input_format) { rc = 3; break; }
        if (c->is_last_chunk) { out->saw_last_chunk = 1; break; }
        const size_t n = (size_t)c->frames_read * resources->input_bytes_per_sample_pair;
        if (out->bytes + n > sink_capacity) { rc = 4; break; }
        memcpy(sink + out->bytes, c->raw_input_data, n);
        out->bytes += n;
        out->chunks++;
        if ((uint64_t)c->frames_read > out->largest_chunk_frames) out->largest_chunk_frames = (uint64_t)c->frames_read;
        c->stream_discontinuity_event = true;
        queue_enqueue(&free_q, c);
    }
    if (rc) { request_shutdown(); queue_signal_shutdown(&free_q); queue_signal_shutdown(&reader_out); }
    pthread_join(reader, NULL);
    out->total_frames_read = resources->total_frames_read;

    InputSummaryInfo summary;
    memset(&summary, 0, sizeof(summary));
    api->get_summary_info(&ctx, &summary);
    out->summary_count = summary.count;
    for (int i = 0; i < summary.count && i < 16; i++) {
        strncpy(out->summary_label[i], summary.items[i].label, 63);
        strncpy(out->summary_value[i], summary.items[i].value, 127);
    }
    for (unsigned i = 0; i < pool_chunks; i++) free(pool[i].raw_input_data);
    free(pool);
    queue_destroy(&reader_out);
    queue_destroy(&free_q);
done:
    api->cleanup(&ctx);
    if (resources->input_module_private_data) rc = rc ? rc : 5;      /* cleanup must drop the private state */
    mem_arena_destroy(&resources->setup_arena);
    pthread_mutex_destroy(&resources->progress_mutex);
    free(config);
    free(resources);
    return rc;
}

/* options reach a module the way argparse delivers them: through the value pointers of its option table */
static void set_option(const struct argparse_option *opts, int n, const char *long_name, float f, const char *str)
{
    for (int i = 0; i < n; i++) {
        if (!opts[i].long_name || strcmp(opts[i].long_name, long_name) != 0) continue;
        if (opts[i].type == ARGPARSE_OPT_FLOAT) *(float *)opts[i].value = f;
        if (opts[i].type == ARGPARSE_OPT_STRING) *(const char **)opts[i].value = str;
    }
}

int wavmod_run(const char *path, float center_target_hz, float freq_shift_hz_arg, int iq_correction, unsigned pool_chunks, unsigned chunk_frames,
               unsigned char *sink, size_t sink_capacity, wavmod_result *out)
{
    int n = 0;
    const struct argparse_option *opts = wav_get_cli_options(&n);
    set_option(opts, n, "wav-center-target-freq", center_target_hz, NULL);
    g_api = get_wav_input_module_api();
    return drive_input(path, freq_shift_hz_arg, iq_correction, pool_chunks, chunk_frames, sink, sink_capacity, out);
}

/* src/input_rawfile.c's module: rate <= 0 / format NULL = option not given */
int rawmod_run(const char *path, float rate_hz, const char *format, int iq_correction, unsigned pool_chunks, unsigned chunk_frames,
               unsigned char *sink, size_t sink_capacity, wavmod_result *out)
{
    int n = 0;
    const struct argparse_option *opts = rawfile_get_cli_options(&n);
    set_option(opts, n, "raw-file-input-rate", rate_hz, NULL);
    set_option(opts, n, "raw-file-input-sample-format", 0.0f, format);
    g_api = get_raw_file_input_module_api();
    return drive_input(path, 0.0f, iq_correction, pool_chunks, chunk_frames, sink, sink_capacity, out);
}

/* ---- output side: the reference's own WAV / RF64 wrappers (src/output_wav.c, src/output_wav_rf64.c, compiled in
 * place) on top of the drop-in output_wav_common.c, driven like the Writer thread side of src/pipeline.c ----------- */
typedef struct {
    int       validated, initialized;
    long long final_output_size_bytes;
    unsigned long long total_output_frames;
    unsigned long long progress_calls, progress_last_bytes;
    int       summary_count;
    char      summary_label[4][64];
    char      summary_value[4][128];
} wavout_result;

static void on_progress(unsigned long long frames, long long total, unsigned long long bytes, void *udata)
{
    wavout_result *r = (wavout_result *)udata;
    (void)frames; (void)total;
    r->progress_calls++;
    r->progress_last_bytes = bytes;
}

static void *writer_main(void *arg)
{
    ModuleContext *ctx = (ModuleContext *)((void **)arg)[0];
    OutputModuleInterface *api = (OutputModuleInterface *)((void **)arg)[1];
    return api->run_writer(ctx);
}

/* mode 0: Writer thread fed through the ring buffer in `piece`-byte writes; mode 1: write_chunk called directly */
int wavout_run(const char *path, int rf64, int output_format, double target_rate, const unsigned char *data, size_t bytes,
               size_t piece, int mode, wavout_result *out)
{
    memset(out, 0, sizeof(*out));
    reset_shutdown_flag();
    AppConfig *config = (AppConfig *)calloc(1, sizeof(AppConfig));
    AppResources *resources = (AppResources *)calloc(1, sizeof(AppResources));
    if (!config || !resources || !mem_arena_init(&resources->setup_arena, 1u << 20)) return -1;
    config->effective_output_filename = (char *)path;
    config->output_format = (format_t)output_format;
    config->output_sample_format_name = (char *)"as given";
    config->target_rate = target_rate;
    pthread_mutex_init(&resources->progress_mutex, NULL);
    resources->output_bytes_per_sample_pair = output_format == CS16 ? 4 : 2;
    resources->progress_callback = on_progress;
    resources->progress_callback_udata = out;
    resources->expected_total_output_frames = -1;

    ModuleContext ctx = {config, resources};
    OutputModuleInterface *api = rf64 ? get_wav_rf64_output_module_api() : get_wav_output_module_api();
    int rc = 0;
    if (!api->validate_options(config)) { rc = 1; goto done; }
    out->validated = 1;
    if (!api->initialize(&ctx)) { rc = 2; goto done; }
    out->initialized = 1;
    if (mode == 0) {
        resources->writer_input_buffer = ring_buffer_create(3u << 20);
        resources->writer_local_buffer = malloc(IO_OUTPUT_WRITER_CHUNK_SIZE);
        void *args[2] = {&ctx, api};
        pthread_t writer;
        pthread_create(&writer, NULL, writer_main, args);
        for (size_t off = 0; off < bytes;) {
            /* ring_buffer_write never blocks: it takes what fits (the reference paces its reader on the fill level) */
            const size_t n = bytes - off < piece ? bytes - off : piece;
            const size_t taken = ring_buffer_write(resources->writer_input_buffer, data + off, n);
            off += taken;
            if (taken < n) usleep(200);
        }
        ring_buffer_signal_end_of_stream(resources->writer_input_buffer);
        pthread_join(writer, NULL);
        free(resources->writer_local_buffer);
        ring_buffer_destroy(resources->writer_input_buffer);
    } else {
        for (size_t off = 0; off < bytes;) {
            const size_t n = bytes - off < piece ? bytes - off : piece;
            if (api->write_chunk(&ctx, data + off, n) != n) { rc = 3; break; }
            off += n;
        }
    }
    api->finalize_output(&ctx);
    out->final_output_size_bytes = resources->final_output_size_bytes;
    out->total_output_frames = resources->total_output_frames;
    OutputSummaryInfo summary;
    memset(&summary, 0, sizeof(summary));
    api->get_summary_info(&ctx, &summary);
    out->summary_count = summary.count;
    for (int i = 0; i < summary.count && i < 4; i++) {
        strncpy(out->summary_label[i], summary.items[i].label, 63);
        strncpy(out->summary_value[i], summary.items[i].value, 127);
    }
done:
    mem_arena_destroy(&resources->setup_arena);
    pthread_mutex_destroy(&resources->progress_mutex);
    free(config);
    free(resources);
    return rc;
}
