/*
 * sample_convert.c (GPU drop-in) — replaces reference src/sample_convert.c.
 * Same prototypes (include/sample_convert.h:19,35,50); conversions run on the GPU through the
 * C ABI (bit exact with the reference, tests/test_gpu_parity.py conversion KATs).
 */
#include "sample_convert.h"

#include "iqgpu.h"
#include "log.h"

size_t get_bytes_per_sample(format_t format)
{
    return iqgpu_get_bytes_per_sample((int)format);            /* sample_convert.c:102-123 */
}

bool convert_block_to_cf32(const void *restrict input_buffer, complex_float_t *restrict output_buffer,
                           size_t num_frames, format_t input_format, float gain)
{
    if (iqgpu_convert_block_to_cf32(input_buffer, (float *)output_buffer, num_frames, (int)input_format, gain) != IQGPU_OK) {
        log_error("convert_block_to_cf32: %s", iqgpu_last_error());   /* "Unhandled input format", :207 */
        return false;
    }
    return true;
}

bool convert_cf32_to_block(const complex_float_t *restrict input_buffer, void *restrict output_buffer,
                           size_t num_frames, format_t output_format)
{
    if (iqgpu_convert_cf32_to_block((const float *)input_buffer, output_buffer, num_frames, (int)output_format) != IQGPU_OK) {
        log_error("convert_cf32_to_block: %s", iqgpu_last_error());   /* "Unhandled output format", :304 */
        return false;
    }
    return true;
}
