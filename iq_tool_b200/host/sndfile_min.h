/* sndfile_min.h — the two libsndfile calls the reference makes on an open capture outside its input
 * modules (iq_correct_run_initial_calibration: sf_read_raw + sf_seek, include/iq_correct.h:76), served from
 * a plain FILE positioned on the data chunk that iqgpu_wav_probe located.  Used by host/input_wav.c so that a
 * GPU build of iq_tool needs neither libsndfile nor expat for WAV captures.  Link sndfile_min.c INSTEAD of
 * libsndfile, never next to it (same symbol names by design). */
#ifndef IQGPU_SNDFILE_MIN_H
#define IQGPU_SNDFILE_MIN_H
#include <stdint.h>
#include <stdio.h>
#include <sndfile.h>     /* the SNDFILE / sf_count_t typedefs (the real header, or oracle/stubs/sndfile.h) */

struct SNDFILE_tag {
    FILE    *file;
    /* reading: a capture positioned on its data chunk */
    uint64_t data_offset;    /* first sample byte */
    uint64_t data_bytes;     /* whole frames only */
    uint64_t position;       /* bytes consumed from the data chunk */
    uint32_t frame_bytes;    /* bytes per I/Q pair */
    /* writing: a WAV / RF64 file whose header is patched on close */
    int      writing;
    int      container;      /* IQGPU_CONTAINER_WAV / _RF64 */
    int      sample_format;  /* IQGPU_FMT_CS16 / _CU8 */
    int      sample_rate_hz;
    uint64_t bytes_written;
};

SNDFILE *sfmin_open(const char *path, uint64_t data_offset, uint64_t data_bytes, uint32_t frame_bytes);
void     sfmin_close(SNDFILE *s);
/* writer side (what SFM_WRITE + sf_write_raw + sf_close amount to for 2-channel PCM) */
SNDFILE *sfmin_create(const char *path, int container, int sample_format, int sample_rate_hz);
sf_count_t sfmin_write_raw(SNDFILE *s, const void *ptr, sf_count_t bytes);
int      sfmin_finish(SNDFILE *s);   /* patches the sizes into the header and closes; 0 on success */
#endif
