/* file_modules.c — see file_modules.h.  Behaviour follows the reference's file modules (file:line under
 * /root/reference): src/input_wav.c:472-632,701-729, src/input_rawfile.c:83-171,252-302, their Reader loops
 * (:634-699 / :173-250, in file_reader.c) and src/output_wav_common.c:27-174. */
#include "file_modules.h"

#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "app_context.h"
#include "constants.h"
#include "iq_correct.h"
#include "log.h"
#include "memory_arena.h"
#include "output_wav_common.h"
#include "ring_buffer.h"
#include "sample_convert.h"
#include "signal_handler.h"
#include "utils.h"

#include "file_reader.h"

/* ================================================================================================ captures */
static IqCapture *capture_of(const ModuleContext *ctx) { return (IqCapture *)ctx->resources->input_module_private_data; }

static IqCapture *new_capture(ModuleContext *ctx, const char *kind)
{
    IqCapture *cap = (IqCapture *)mem_arena_alloc(&ctx->resources->setup_arena, sizeof(IqCapture), true);
    if (cap) cap->kind = kind;
    ctx->resources->input_module_private_data = cap;
    return cap;
}

static void adopt_format(AppResources *resources, format_t fmt)
{
    resources->input_format = fmt;
    resources->input_bytes_per_sample_pair = get_bytes_per_sample(fmt);
}

bool iqcap_open_wav(ModuleContext *ctx, float center_target_hz_arg)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    IqCapture *cap = new_capture(ctx, "WAV");
    if (!cap) return false;

    const char *path = config->effective_input_filename;
    log_info("Opening WAV input file: %s", path);
    if (iqgpu_wav_probe(path, &cap->wav) != IQGPU_OK) {         /* not a WAV, != 2 channels, PCM subtype, rate */
        log_fatal("%s", iqgpu_rawfile_last_error());
        return false;
    }
    adopt_format(resources, (format_t)cap->wav.sample_format);  /* CS16 or CU8 */
    if (cap->wav.frames == 0) log_warn("Warning: Input file appears to be empty (0 frames).");
    resources->source_info.samplerate = cap->wav.sample_rate_hz;
    resources->source_info.frames = (int64_t)cap->wav.frames;

    double shift = 0.0;
    if (iqgpu_wav_center_target_shift(&cap->wav, center_target_hz_arg, (double)config->freq_shift_hz_arg, &shift) != IQGPU_OK) {
        log_fatal("%s", iqgpu_rawfile_last_error());
        return false;
    }
    if (center_target_hz_arg != 0.0f) resources->nco_shift_hz = shift;

    cap->handle = sfmin_open(path, cap->wav.data_offset, cap->wav.data_bytes, (uint32_t)resources->input_bytes_per_sample_pair);
    if (!cap->handle) log_fatal("Error opening input file: %s", path);
    return cap->handle != NULL;
}

bool iqcap_open_raw(ModuleContext *ctx, const char *format_name, double rate_hz)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    IqCapture *cap = new_capture(ctx, "RAW");
    if (!cap) return false;
    cap->raw_format = format_name;
    cap->raw_rate_hz = rate_hz;

    const format_t fmt = utils_get_format_from_string(format_name);
    if (fmt == FORMAT_UNKNOWN) {
        log_fatal("Invalid RAW input format '%s'. See --help for valid formats.", format_name);
        return false;
    }
    adopt_format(resources, fmt);
    if (resources->input_bytes_per_sample_pair == 0) {
        log_fatal("Internal error: could not determine sample size for format '%s'.", format_name);
        return false;
    }
    /* the formats the reference can open as a headerless stream; cs24 is not among them */
    static const format_t openable[] = {SC16Q11, CS16, CU16, CS8, CU8, CS32, CU32, CF32};
    bool known = false;
    for (size_t i = 0; i < sizeof(openable) / sizeof(openable[0]); i++) known |= openable[i] == fmt;
    if (!known) { log_fatal("Internal error: unhandled format enum in rawfile_initialize."); return false; }

    const char *path = config->effective_input_filename;
    log_info("Opening RAW input file: %s", path);
    struct stat sb;
    const bool regular = stat(path, &sb) == 0 && S_ISREG(sb.st_mode);
    if (regular) cap->handle = sfmin_open(path, 0, (uint64_t)sb.st_size, (uint32_t)resources->input_bytes_per_sample_pair);
    if (!cap->handle) {
        log_fatal("Error opening RAW input file '%s'.", config->input_filename_arg);
        return false;
    }
    resources->source_info.samplerate = (int)rate_hz;
    resources->source_info.frames = (int64_t)((uint64_t)sb.st_size / resources->input_bytes_per_sample_pair);
    return true;
}

void *iqcap_stream(ModuleContext *ctx)
{
    IqCapture *cap = capture_of(ctx);
    iqgpu_file_reader_loop(ctx, cap->handle, cap->kind);
    return NULL;
}

void iqcap_close(ModuleContext *ctx)
{
    IqCapture *cap = capture_of(ctx);
    if (!cap) return;
    if (cap->handle) {
        log_info("Closing %s input file.", cap->kind);
        sfmin_close(cap->handle);
        cap->handle = NULL;
    }
    ctx->resources->input_module_private_data = NULL;
}

bool iqcap_calibrate_before_streaming(ModuleContext *ctx)
{
    /* the calibration service reads the first block through the module's handle and rewinds it */
    if (!ctx->config->iq_correction.enable) return true;
    return iq_correct_run_initial_calibration(ctx, capture_of(ctx)->handle);
}

static void describe_wav_metadata(const iqgpu_wav_info *wi, InputSummaryInfo *info)
{
    if (!wi->metadata_present) return;
    if (wi->timestamp_unix_present) {
        const time_t when = (time_t)wi->timestamp_unix;
        struct tm utc;
        char text[64];
        if (gmtime_r(&when, &utc)) {
            strftime(text, sizeof(text), "%Y-%m-%d %H:%M:%S UTC", &utc);
            add_summary_item(info, "Timestamp", "%s", text);
        }
    } else if (wi->timestamp_str_present) {
        add_summary_item(info, "Timestamp", "%s", wi->timestamp_str);
    }
    if (wi->center_freq_hz_present) add_summary_item(info, "Center Frequency", "%.0f Hz", wi->center_freq_hz);
    if (wi->software_name_present) {
        char text[130];
        snprintf(text, sizeof(text), "%s %s", wi->software_name, wi->software_version_present ? wi->software_version : "");
        add_summary_item(info, "SDR Software", "%s", text);
    }
    if (wi->radio_model_present) add_summary_item(info, "Radio Model", "%s", wi->radio_model);
}

void iqcap_describe(const ModuleContext *ctx, InputSummaryInfo *info)
{
    const AppResources *resources = ctx->resources;
    const IqCapture *cap = capture_of(ctx);
    const char *shown = ctx->config->input_filename_arg;
    const bool is_wav = cap->kind[0] == 'W';
    char size_text[40];
    long long size_bytes;

    add_summary_item(info, "Input File", "%s", shown);
    if (is_wav) {
        add_summary_item(info, "Input Format", "%s", resources->input_format == CS16 ? "16-bit Signed Complex PCM (cs16)"
                                                   : resources->input_format == CU8 ? "8-bit Unsigned Complex PCM (cu8)" : "Unknown PCM");
        add_summary_item(info, "Input Rate", "%.0f Hz", (double)resources->source_info.samplerate);
        struct stat sb;
        size_bytes = stat(shown, &sb) == 0 ? (long long)sb.st_size : -1LL;           /* the file, header and all */
    } else {
        add_summary_item(info, "Input Type", "RAW FILE");
        add_summary_item(info, "Input Format", "%s", cap->raw_format);
        add_summary_item(info, "Input Rate", "%.0f Hz", cap->raw_rate_hz);
        size_bytes = (long long)(resources->source_info.frames * (int64_t)resources->input_bytes_per_sample_pair);
    }
    add_summary_item(info, "Input File Size", "%s", format_file_size(size_bytes, size_text, sizeof(size_text)));
    if (is_wav) describe_wav_metadata(&cap->wav, info);
}

/* ================================================================================================ WAV / RF64 sink */
static WavCommonData *sink_of(ModuleContext *ctx) { return (WavCommonData *)ctx->resources->output_module_private_data; }

static void count_written(WavCommonData *sink, sf_count_t done)
{
    if (done > 0) sink->total_bytes_written += done;
}

/* the y/n question asked when the output file exists */
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

bool iqsink_format_allowed(AppConfig *config)
{
    const bool ok = config->output_format == CS16 || config->output_format == CU8;
    if (!ok) log_fatal("Invalid sample format '%s' for WAV/RF64 container. Only 'cs16' and 'cu8' are supported.", config->output_sample_format_name);
    return ok;
}

bool iqsink_open(ModuleContext *ctx, int sf_format_flag)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    WavCommonData *sink = (WavCommonData *)mem_arena_alloc(&resources->setup_arena, sizeof(WavCommonData), true);
    if (!sink) return false;
    resources->output_module_private_data = sink;

    const char *path = config->effective_output_filename;
    struct stat sb;
    if (lstat(path, &sb) == 0) {
        if (!S_ISREG(sb.st_mode)) { log_fatal("Output path '%s' exists but is not a regular file. Aborting.", path); return false; }
        if (!overwrite_confirmed(path)) return false;
    }
    if (config->output_format != CS16 && config->output_format != CU8) return false;     /* validation should have caught it */
    const int container = (sf_format_flag & SF_FORMAT_TYPEMASK) == SF_FORMAT_RF64 ? IQGPU_CONTAINER_RF64 : IQGPU_CONTAINER_WAV;
    const int rate = (int)config->target_rate;
    unsigned char trial[80];
    if (iqgpu_wav_build_header(container, (int)config->output_format, rate, 0, trial, sizeof(trial)) != IQGPU_OK) {
        log_fatal("The requested container format is not supported (Rate: %d, Format: 0x%08X).", rate, sf_format_flag);
        return false;
    }
    sink->handle = sfmin_create(path, container, (int)config->output_format, rate);
    if (!sink->handle) log_fatal("Error opening output WAV file %s", path);
    return sink->handle != NULL;
}

/* ring buffer -> file in IO_OUTPUT_WRITER_CHUNK_SIZE pieces, progress after every piece */
void *iqsink_drain_ring(ModuleContext *ctx)
{
    AppResources *resources = ctx->resources;
    WavCommonData *sink = sink_of(ctx);
    unsigned char *staging = (unsigned char *)resources->writer_local_buffer;
    if (!staging) { handle_fatal_thread_error("WAV writer: Local buffer is NULL.", resources); return NULL; }

    size_t n;
    while ((n = ring_buffer_read(resources->writer_input_buffer, staging, IO_OUTPUT_WRITER_CHUNK_SIZE)) != 0) {   /* 0: end of stream / shutdown */
        const sf_count_t done = sfmin_write_raw(sink->handle, staging, (sf_count_t)n);
        count_written(sink, done);
        if ((size_t)done != n) { handle_fatal_thread_error("WAV writer: File write error.", resources); break; }
        if (!resources->progress_callback) continue;
        const unsigned long long frames = (unsigned long long)sink->total_bytes_written / resources->output_bytes_per_sample_pair;
        pthread_mutex_lock(&resources->progress_mutex);
        resources->total_output_frames = frames;
        pthread_mutex_unlock(&resources->progress_mutex);
        resources->progress_callback(frames, resources->expected_total_output_frames, (unsigned long long)sink->total_bytes_written,
                                     resources->progress_callback_udata);
    }
    log_debug("Common WAV writer thread is exiting.");
    return NULL;
}

size_t iqsink_put(ModuleContext *ctx, const void *bytes, size_t count)
{
    WavCommonData *sink = sink_of(ctx);
    if (!sink || !sink->handle || count == 0) return 0;
    const sf_count_t done = sfmin_write_raw(sink->handle, bytes, (sf_count_t)count);
    count_written(sink, done);
    return (size_t)done;
}

/* closing is what puts the sizes into the header */
void iqsink_close(ModuleContext *ctx)
{
    WavCommonData *sink = sink_of(ctx);
    if (!sink) return;
    if (sink->handle && sfmin_finish(sink->handle) != 0) log_warn("Could not finalise the WAV header of the output file.");
    sink->handle = NULL;
    ctx->resources->final_output_size_bytes = sink->total_bytes_written;
}
