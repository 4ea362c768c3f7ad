/* input_wav.c — drop-in replacement for the reference's src/input_wav.c (the WAV input module,
 * include/input_wav.h): same exported functions (get_wav_input_module_api, wav_get_cli_options) and the same
 * InputModuleInterface behaviour, log and summary texts — but the container and its SDR metadata are read by the
 * host-only entry points of libiqgpu.so (iqgpu_wav_probe, iqgpu_wav_center_target_shift; csrc/wavfile.cpp) instead
 * of libsndfile + expat, which a GPU build therefore does not need.  This file holds the
 * option table and the v-table; the capture object lives in file_modules.c, the Reader loop in file_reader.c. */
#include "input_wav.h"

#include "input_common.h"

#include "file_modules.h"

/* --wav-center-target-freq reaches the module through this table (src/input_wav.c:434-447) */
static float s_center_target_hz;
static const struct argparse_option s_options[] = {
    OPT_GROUP("WAV Input Specific Options"),
    OPT_FLOAT(0, "wav-center-target-freq", &s_center_target_hz, "Shift signal to a new target center frequency (e.g., 97.3e6)", NULL, 0, 0),
};

const struct argparse_option *wav_get_cli_options(int *count)
{
    *count = (int)(sizeof(s_options) / sizeof(s_options[0]));
    return s_options;
}

static bool open_capture(ModuleContext *ctx) { return iqcap_open_wav(ctx, s_center_target_hz); }
static void nothing_to_stop(ModuleContext *ctx) { (void)ctx; }

InputModuleInterface *get_wav_input_module_api(void)
{
    static InputModuleInterface api = {
        .initialize = open_capture,
        .start_stream = iqcap_stream,
        .stop_stream = nothing_to_stop,
        .cleanup = iqcap_close,
        .get_summary_info = iqcap_describe,
        .validate_options = NULL,
        .validate_generic_options = NULL,
        .has_known_length = _input_source_has_known_length_true,
        .pre_stream_iq_correction = iqcap_calibrate_before_streaming,
    };
    return &api;
}
