// stream_io.hpp — the reader / chain / writer pass shared by the raw-file and WAV-container entry points
// (rawfile.cpp, wavfile.cpp).  Host code only; uses the C ABI of include/iqgpu.h.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/iqgpu.h"

namespace iqio {

// Streams at most `in_limit_bytes` from the current position of `fin` through `chain` into `fout`
// (three overlapped stages on pinned rings, see rawfile.cpp).  The caller owns the chain and both files.
int run_stream(iqgpu_chain* chain, const iqgpu_chain_config* cfg, FILE* fin, uint64_t in_limit_bytes, FILE* fout,
               size_t train_chunks, iqgpu_rawfile_stats* stats, std::string& err);

// sets the calling thread's message behind iqgpu_rawfile_last_error()
void set_last_error(const std::string& msg);

}  // namespace iqio
