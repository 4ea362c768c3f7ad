/* input_wav.c — drop-in replacement for the reference's src/input_wav.c (the WAV input module,
 * include/input_wav.h:15,20): same exported functions (get_wav_input_module_api, wav_get_cli_options), same
 * InputModuleInterface behaviour, same log / summary text — but the container and its SDR metadata are read by
 * the host-only entry points of libiqgpu.so (iqgpu_wav_probe, iqgpu_wav_center_target_shift; csrc/wavfile.cpp)
 * instead of libsndfile + expat, which a GPU build therefore does not need.
 *
 * The Reader thread keeps the reference's shape (src/input_wav.c:634-699): one SampleChunk from the free queue
 * per read, frames_read / is_last_chunk bookkeeping, writer back-pressure; the chunks then go through the
 * drop-in stage layer (host/pre_processor.c ...), which gathers them into trains for the GPU. */
#include "input_wav.h"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "app_context.h"
#include "constants.h"
#include "input_common.h"
#include "iq_correct.h"
#include "log.h"
#include "memory_arena.h"
#include "sample_convert.h"
#include "utils.h"

#include "iqgpu.h"
#include "file_reader.h"
#include "sndfile_min.h"

typedef struct {
    SNDFILE       *capture;        /* FILE on the data chunk (sndfile_min.h) */
    iqgpu_wav_info info;           /* header + SdrMetadata as wav_initialize sees them */
} WavModuleState;

/* --wav-center-target-freq, registered with the CLI through wav_get_cli_options (src/input_wav.c:434-447) */
static struct { float center_target_hz_arg; } s_wav_options;
static const struct argparse_option s_wav_cli_options[] = {
    OPT_GROUP("WAV Input Specific Options"),
    OPT_FLOAT(0, "wav-center-target-freq", &s_wav_options.center_target_hz_arg,
              "Shift signal to a new target center frequency (e.g., 97.3e6)", NULL, 0, 0),
};

const struct argparse_option *wav_get_cli_options(int *count)
{
    *count = (int)(sizeof(s_wav_cli_options) / sizeof(s_wav_cli_options[0]));
    return s_wav_cli_options;
}

static WavModuleState *state_of(const ModuleContext *ctx) { return (WavModuleState *)ctx->resources->input_module_private_data; }

/* src/input_wav.c:542-632 */
static bool wav_initialize(ModuleContext *ctx)
{
    const AppConfig *config = ctx->config;
    AppResources *resources = ctx->resources;
    WavModuleState *st = (WavModuleState *)mem_arena_alloc(&resources->setup_arena, sizeof(WavModuleState), true);
    if (!st) return false;
    resources->input_module_private_data = st;

    const char *path = config->effective_input_filename;
    log_info("Opening WAV input file: %s", path);
    if (iqgpu_wav_probe(path, &st->info) != IQGPU_OK) {       /* not a WAV, != 2 channels, PCM subtype, rate */
        log_fatal("%s", iqgpu_rawfile_last_error());
        return false;
    }
    resources->input_format = (format_t)st->info.sample_format;                    /* CS16 or CU8 */
    resources->input_bytes_per_sample_pair = get_bytes_per_sample(resources->input_format);
    if (st->info.frames == 0) log_warn("Warning: Input file appears to be empty (0 frames).");
    resources->source_info.samplerate = st->info.sample_rate_hz;
    resources->source_info.frames = (int64_t)st->info.frames;

    double shift = 0.0;
    if (iqgpu_wav_center_target_shift(&st->info, s_wav_options.center_target_hz_arg, (double)config->freq_shift_hz_arg, &shift) != IQGPU_OK) {
        log_fatal("%s", iqgpu_rawfile_last_error());
        return false;
    }
    if (s_wav_options.center_target_hz_arg != 0.0f) resources->nco_shift_hz = shift;

    st->capture = sfmin_open(path, st->info.data_offset, st->info.data_bytes, (uint32_t)resources->input_bytes_per_sample_pair);
    if (!st->capture) {
        log_fatal("Error opening input file: %s", path);
        return false;
    }
    return true;
}

/* src/input_wav.c:634-699 */
static void *wav_start_stream(ModuleContext *ctx)
{
    iqgpu_file_reader_loop(ctx, state_of(ctx)->capture, "WAV");
    return NULL;
}

static void wav_stop_stream(ModuleContext *ctx) { (void)ctx; }

static void wav_cleanup(ModuleContext *ctx)
{
    WavModuleState *st = state_of(ctx);
    if (!st) return;
    if (st->capture) {
        log_info("Closing WAV input file.");
        sfmin_close(st->capture);
        st->capture = NULL;
    }
    ctx->resources->input_module_private_data = NULL;
}

/* src/input_wav.c:472-540: the lines of the start-up summary */
static void wav_get_summary_info(const ModuleContext *ctx, InputSummaryInfo *info)
{
    const AppConfig *config = ctx->config;
    const AppResources *resources = ctx->resources;
    const iqgpu_wav_info *wi = &state_of(ctx)->info;
    const char *shown = config->input_filename_arg;

    add_summary_item(info, "Input File", "%s", shown);
    add_summary_item(info, "Input Format", "%s", resources->input_format == CS16 ? "16-bit Signed Complex PCM (cs16)"
                                               : resources->input_format == CU8 ? "8-bit Unsigned Complex PCM (cu8)" : "Unknown PCM");
    add_summary_item(info, "Input Rate", "%.0f Hz", (double)resources->source_info.samplerate);
    struct stat sb;
    char size_text[40];
    add_summary_item(info, "Input File Size", "%s", format_file_size(stat(shown, &sb) == 0 ? (long long)sb.st_size : -1LL, size_text, sizeof(size_text)));
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

/* src/input_wav.c:718-729: the calibration service reads the first block through the module's handle */
static bool wav_pre_stream_iq_correction(ModuleContext *ctx)
{
    if (!ctx->config->iq_correction.enable) return true;
    return iq_correct_run_initial_calibration(ctx, state_of(ctx)->capture);
}

static InputModuleInterface s_wav_module = {
    .initialize = wav_initialize,
    .start_stream = wav_start_stream,
    .stop_stream = wav_stop_stream,
    .cleanup = wav_cleanup,
    .get_summary_info = wav_get_summary_info,
    .validate_options = NULL,
    .validate_generic_options = NULL,
    .has_known_length = _input_source_has_known_length_true,
    .pre_stream_iq_correction = wav_pre_stream_iq_correction,
};

InputModuleInterface *get_wav_input_module_api(void) { return &s_wav_module; }
