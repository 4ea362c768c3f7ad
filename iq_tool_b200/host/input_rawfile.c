/* input_rawfile.c — drop-in replacement for the reference's src/input_rawfile.c (the raw-file input module,
 * include/input_rawfile.h): same exported functions (get_raw_file_input_module_api, rawfile_get_cli_options), the
 * same two required options, refusals, summary lines and Reader-thread behaviour — on a plain FILE instead of
 * libsndfile's headerless reader.  With input_wav.c and output_wav_common.c this removes libsndfile from a GPU build
 * of iq_tool.  Option table and v-table only; the capture object is in file_modules.c. */
#include "input_rawfile.h"

#include <stdio.h>

#include "app_context.h"
#include "input_common.h"
#include "log.h"
#include "signal_handler.h"

#include "file_modules.h"

/* --raw-file-input-rate / --raw-file-input-sample-format (src/input_rawfile.c:35-58) */
static float  s_rate_arg;
static char  *s_format_arg;
static double s_rate_hz;           /* sticky once a positive rate was seen, like the reference's */
static const struct argparse_option s_options[] = {
    OPT_GROUP("Raw File Input Options"),
    OPT_FLOAT(0, "raw-file-input-rate", &s_rate_arg, "(Required) The sample rate of the RAW input file.", NULL, 0, 0),
    OPT_STRING(0, "raw-file-input-sample-format", &s_format_arg, "(Required) The sample format of the RAW input file.", NULL, 0, 0),
};

const struct argparse_option *rawfile_get_cli_options(int *count)
{
    *count = (int)(sizeof(s_options) / sizeof(s_options[0]));
    return s_options;
}

/* both options are mandatory (src/input_rawfile.c:83-103) */
static bool both_options_given(AppConfig *config)
{
    (void)config;
    if (s_rate_arg > 0.0f) s_rate_hz = (double)s_rate_arg;
    const char *missing = s_rate_hz <= 0.0 ? "--raw-file-input-rate <hz>" : !s_format_arg ? "--raw-file-input-sample-format <format>" : NULL;
    if (missing) log_fatal("Missing required option %s for raw file input.", missing);
    return missing == NULL;
}

static bool open_capture(ModuleContext *ctx) { return iqcap_open_raw(ctx, s_format_arg, s_rate_hz); }
static void nothing_to_stop(ModuleContext *ctx) { (void)ctx; }

/* --raw-passthrough needs identical formats (src/input_rawfile.c:178-186), then the shared Reader loop */
static void *stream(ModuleContext *ctx)
{
    const AppConfig *config = ctx->config;
    if (config->raw_passthrough && ctx->resources->input_format != config->output_format) {
        char msg[256];
        snprintf(msg, sizeof(msg), "Option --raw-passthrough requires input and output formats to be identical. "
                 "Input format is '%s', output format is '%s'.", s_format_arg, config->output_sample_format_name);
        handle_fatal_thread_error(msg, ctx->resources);
        return NULL;
    }
    return iqcap_stream(ctx);
}

InputModuleInterface *get_raw_file_input_module_api(void)
{
    static InputModuleInterface api = {
        .initialize = open_capture,
        .start_stream = stream,
        .stop_stream = nothing_to_stop,
        .cleanup = iqcap_close,
        .get_summary_info = iqcap_describe,
        .validate_options = both_options_given,
        .validate_generic_options = NULL,
        .has_known_length = _input_source_has_known_length_true,
        .pre_stream_iq_correction = iqcap_calibrate_before_streaming,
    };
    return &api;
}
