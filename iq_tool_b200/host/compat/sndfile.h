/* compat/sndfile.h — the sliver of <sndfile.h> the drop-in translation units need to compile without libsndfile
 * installed: the SNDFILE / sf_count_t types and the two raw calls iq_correct_run_initial_calibration makes on an input
 * handle (include/iq_correct.h:76).  A build that links the real libsndfile (the reference's normal build) uses its own
 * header instead; this directory goes on the include path only where the library is absent (the test harness). */
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
