/* compat/sndfile.h — the sliver of <sndfile.h> that iq_tool's sources need once the WAV modules are served by
 * libiqgpu's host-only container code (host/input_wav.c, host/output_wav_common.c, host/sndfile_min.c): the
 * SNDFILE / sf_count_t types, the two raw calls the calibration service makes on an input handle, and the
 * container / subtype constants the output wrappers (src/output_wav.c, src/output_wav_rf64.c) pass down.
 * Put this directory on the include path ONLY in a build that does not link libsndfile. */
#ifndef IQGPU_COMPAT_SNDFILE_H
#define IQGPU_COMPAT_SNDFILE_H
#include <stdint.h>
#include <stdio.h>

typedef struct SNDFILE_tag SNDFILE;
typedef int64_t sf_count_t;

/* numeric values of libsndfile's public enum (sndfile.h) */
#define SF_FORMAT_WAV      0x010000
#define SF_FORMAT_RF64     0x220000
#define SF_FORMAT_PCM_16   0x0002
#define SF_FORMAT_PCM_U8   0x0005
#define SF_FORMAT_SUBMASK  0x0000FFFF
#define SF_FORMAT_TYPEMASK 0x0FFF0000

sf_count_t sf_read_raw(SNDFILE *sndfile, void *ptr, sf_count_t bytes);
sf_count_t sf_seek(SNDFILE *sndfile, sf_count_t frames, int whence);
#endif
