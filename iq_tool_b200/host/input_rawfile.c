/* input_rawfile.c — drop-in replacement for the reference's src/input_rawfile.c (the raw-file input module,
 * include/input_rawfile.h): same exported functions (get_raw_file_input_module_api, rawfile_get_cli_options), the same
 * two required options, refusals, summary lines and Reader-thread behaviour — on a plain FILE (host/sndfile_min.c)
 * instead of libsndfile's SF_FORMAT_RAW reader.  Together with host/input_wav.c and host/output_wav_common.c this
 * removes libsndfile from a GPU build of iq_tool. */
#include "input_rawfile.h"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include "app_context.h"
#include "input_common.h"
#include "iq_correct.h"
#include "log.h"
#include "memory_arena.h"
#include "sample_convert.h"
#include "signal_handler.h"
#include "utils.h"

#include "file_reader.h"
#include "sndfile_min.h"

typedef struct { SNDFILE *capture; } RawModuleState;

/* --raw-file-input-rate / --raw-file-input-sample-format (src/input_rawfile.c:35-58) */
static struct {
    float  rate_arg;
    char  *format_arg;
    double rate_hz;
    bool   rate_given;
} s_raw_options;
static const struct argparse_option s_raw_cli_options[] = {
    OPT_GROUP("Raw File Input Options"),
    OPT_FLOAT(0, "raw-file-input-rate", &s_raw_options.rate_arg, "(Required) The sample rate of the RAW input file.", NULL, 0, 0),
    OPT_STRING(0, "raw-file-input-sample-format", &s_raw_options.format_arg, "(Required) The sample format of the RAW input file.", NULL, 0, 0),
};

const struct argparse_option *rawfile_get_cli_options(int *count)
{
    *count = (int)(sizeof(s_raw_cli_options) / sizeof(s_raw_cli_options[0]));
    return s_raw_cli_options;
}

static RawModuleState *state_of(const ModuleContext *ctx) { return (RawModuleState *)ctx->resources->input_module_private_data; }

/* :83-103 */
static bool rawfile_validate_options(AppConfig *config)
{
    (void)config;
    if (s_raw_options.rate_arg > 0.0f) {
        s_raw_options.rate_hz = (double)s_raw_options.rate_arg;
        s_raw_options.rate_given = true;
    }
    if (!s_raw_options.rate_given) { log_fatal("Missing required option --raw-file-input-rate <hz> for raw file input."); return false; }
    if (!s_raw_options.format_arg) { log_fatal("Missing required option --raw-file-input-sample-format <format> for raw file input."); return false; }
    return true;
}

/* :105-171 */
static bool rawfile_initialize(ModuleContext *ctx)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    RawModuleState *st = (RawModuleState *)mem_arena_alloc(&resources->setup_arena, sizeof(RawModuleState), true);
    if (!st) return false;
    resources->input_module_private_data = st;

    resources->input_format = utils_get_format_from_string(s_raw_options.format_arg);
    if (resources->input_format == FORMAT_UNKNOWN) {
        log_fatal("Invalid RAW input format '%s'. See --help for valid formats.", s_raw_options.format_arg);
        return false;
    }
    resources->input_bytes_per_sample_pair = get_bytes_per_sample(resources->input_format);
    if (resources->input_bytes_per_sample_pair == 0) {
        log_fatal("Internal error: could not determine sample size for format '%s'.", s_raw_options.format_arg);
        return false;
    }
    switch (resources->input_format) {      /* the formats the reference can open as a raw stream; cs24 is not among them */
        case SC16Q11: case CS16: case CU16: case CS8: case CU8: case CS32: case CU32: case CF32: break;
        default: log_fatal("Internal error: unhandled format enum in rawfile_initialize."); return false;
    }

    const char *path = config->effective_input_filename;
    log_info("Opening RAW input file: %s", path);
    struct stat sb;
    if (stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        log_fatal("Error opening RAW input file '%s'.", config->input_filename_arg);
        return false;
    }
    st->capture = sfmin_open(path, 0, (uint64_t)sb.st_size, (uint32_t)resources->input_bytes_per_sample_pair);
    if (!st->capture) {
        log_fatal("Error opening RAW input file '%s'.", config->input_filename_arg);
        return false;
    }
    resources->source_info.samplerate = (int)s_raw_options.rate_hz;
    resources->source_info.frames = (int64_t)((uint64_t)sb.st_size / resources->input_bytes_per_sample_pair);
    return true;
}

/* :173-250 */
static void *rawfile_start_stream(ModuleContext *ctx)
{
    AppResources *resources = ctx->resources;
    const AppConfig *config = ctx->config;
    if (config->raw_passthrough && resources->input_format != config->output_format) {
        char msg[256];
        snprintf(msg, sizeof(msg), "Option --raw-passthrough requires input and output formats to be identical. "
                 "Input format is '%s', output format is '%s'.", s_raw_options.format_arg, config->output_sample_format_name);
        handle_fatal_thread_error(msg, resources);
        return NULL;
    }
    iqgpu_file_reader_loop(ctx, state_of(ctx)->capture, "RAW");
    return NULL;
}

static void rawfile_stop_stream(ModuleContext *ctx) { (void)ctx; }

static void rawfile_cleanup(ModuleContext *ctx)
{
    RawModuleState *st = state_of(ctx);
    if (!st) return;
    if (st->capture) {
        log_info("Closing RAW input file.");
        sfmin_close(st->capture);
        st->capture = NULL;
    }
    ctx->resources->input_module_private_data = NULL;
}

/* :270-288 */
static void rawfile_get_summary_info(const ModuleContext *ctx, InputSummaryInfo *info)
{
    const AppResources *resources = ctx->resources;
    char size_text[40];
    add_summary_item(info, "Input File", "%s", ctx->config->input_filename_arg);
    add_summary_item(info, "Input Type", "RAW FILE");
    add_summary_item(info, "Input Format", "%s", s_raw_options.format_arg);
    add_summary_item(info, "Input Rate", "%.0f Hz", s_raw_options.rate_hz);
    add_summary_item(info, "Input File Size", "%s",
                     format_file_size((long long)(resources->source_info.frames * (int64_t)resources->input_bytes_per_sample_pair), size_text, sizeof(size_text)));
}

/* :290-302 */
static bool rawfile_pre_stream_iq_correction(ModuleContext *ctx)
{
    if (!ctx->config->iq_correction.enable) return true;
    return iq_correct_run_initial_calibration(ctx, state_of(ctx)->capture);
}

static InputModuleInterface s_raw_module = {
    .initialize = rawfile_initialize,
    .start_stream = rawfile_start_stream,
    .stop_stream = rawfile_stop_stream,
    .cleanup = rawfile_cleanup,
    .get_summary_info = rawfile_get_summary_info,
    .validate_options = rawfile_validate_options,
    .validate_generic_options = NULL,
    .has_known_length = _input_source_has_known_length_true,
    .pre_stream_iq_correction = rawfile_pre_stream_iq_correction,
};

InputModuleInterface *get_raw_file_input_module_api(void) { return &s_raw_module; }
