/* file_reader.h — the Reader-thread loop shared by the drop-in file input modules (host/input_wav.c,
 * host/input_rawfile.c).  Same shape as the reference's (src/input_wav.c:634-699, src/input_rawfile.c:188-249):
 * one SampleChunk from the free queue per read, frames_read / is_last_chunk bookkeeping, a zero-frame chunk marks
 * the end, and the reader pauses while the writer's ring buffer is above its high-water mark. */
#ifndef IQGPU_FILE_READER_H
#define IQGPU_FILE_READER_H
#include "module.h"
#include "sndfile_min.h"

void iqgpu_file_reader_loop(ModuleContext *ctx, SNDFILE *capture, const char *kind /* "WAV" / "RAW": log text */);
#endif
