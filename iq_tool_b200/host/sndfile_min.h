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
    uint64_t data_offset;    /* first sample byte */
    uint64_t data_bytes;     /* whole frames only */
    uint64_t position;       /* bytes consumed from the data chunk */
    uint32_t frame_bytes;    /* bytes per I/Q pair */
};

SNDFILE *sfmin_open(const char *path, uint64_t data_offset, uint64_t data_bytes, uint32_t frame_bytes);
void     sfmin_close(SNDFILE *s);
#endif
