/*
 * Stub <sndfile.h> — TEST INFRASTRUCTURE ONLY (oracle).
 * The reference's include/iq_correct.h:17 includes <sndfile.h> only for the SNDFILE* type in
 * iq_correct_run_initial_calibration() (iq_correct.c:237-300, file-input calibration), which
 * the oracle never calls.  libsndfile is not installed here; these declarations let the
 * reference's src/iq_correct.c compile in place.  The two functions abort if ever reached.
 */
#ifndef ORACLE_STUB_SNDFILE_H
#define ORACLE_STUB_SNDFILE_H
#include <stdint.h>
#include <stdio.h>
typedef struct SNDFILE_tag SNDFILE;
typedef int64_t sf_count_t;
sf_count_t sf_read_raw(SNDFILE *sndfile, void *ptr, sf_count_t bytes);
sf_count_t sf_seek(SNDFILE *sndfile, sf_count_t frames, int whence);
#endif
