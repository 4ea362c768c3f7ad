// stream_stub.cpp — TEST DOUBLE, never part of libiqgpu.so.  Gives rawfile.cpp / wavfile.cpp (the host-side
// reader / chain / writer plumbing around the path) a stand-in for the chain entry points of include/iqgpu.h
// so the streaming logic — read limits, chunk-train cuts, trailing partial frames, header patching — can be
// exercised on a machine without a GPU.  The stand-in is NOT a signal path: it copies every second input
// frame to the output (same format in and out), keeping the phase of that decimation across calls, so that a
// wrong train cut, a re-read or a dropped frame changes the result.
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/iqgpu.h"

struct iqgpu_chain {
    iqgpu_chain_config cfg;
    uint64_t seen = 0;
    uint64_t calls = 0;
    size_t largest_call = 0;
};

static thread_local std::string g_err;
static uint64_t g_calls = 0, g_largest = 0;

extern "C" {

const char* iqgpu_last_error(void) { return g_err.c_str(); }
void* iqgpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void iqgpu_host_free(void* p) { free(p); }

size_t iqgpu_get_bytes_per_sample(int format)
{
    switch (format) {
        case IQGPU_FMT_CU8: case IQGPU_FMT_CS8: return 2;
        case IQGPU_FMT_CS16: case IQGPU_FMT_CU16: case IQGPU_FMT_SC16Q11: return 4;
        case IQGPU_FMT_CF32: case IQGPU_FMT_CS32: case IQGPU_FMT_CU32: return 8;
        default: return 0;
    }
}

int iqgpu_chain_create(const iqgpu_chain_config* cfg, int device, iqgpu_chain** out)
{
    if (device != 0) { g_err = "stub: no such device"; return IQGPU_ENODEVICE; }
    if (cfg->input_format != cfg->output_format) { g_err = "stub: input and output format must agree"; return IQGPU_EINVAL; }
    *out = new iqgpu_chain{*cfg};
    return IQGPU_OK;
}
void iqgpu_chain_destroy(iqgpu_chain* c)
{
    if (c) { g_calls = c->calls; g_largest = c->largest_call; }
    delete c;
}
int iqgpu_chain_set_option(iqgpu_chain*, const char*, int64_t) { return IQGPU_OK; }
int iqgpu_chain_get_info(iqgpu_chain*, iqgpu_chain_info* info)
{
    memset(info, 0, sizeof(*info));
    info->ratio = 0.5f;
    return IQGPU_OK;
}
int iqgpu_chain_process(iqgpu_chain* c, const void* in, size_t n_frames, const uint32_t*, size_t, void* out, size_t out_capacity_bytes,
                        size_t* out_frames, uint32_t*)
{
    const size_t b = iqgpu_get_bytes_per_sample(c->cfg.input_format);
    size_t produced = 0;
    for (size_t i = 0; i < n_frames; i++, c->seen++) {
        if (c->seen & 1) continue;
        if ((produced + 1) * b > out_capacity_bytes) { g_err = "stub: output capacity"; return IQGPU_ECAPACITY; }
        memcpy((char*)out + produced * b, (const char*)in + i * b, b);
        produced++;
    }
    c->calls++;
    if (n_frames > c->largest_call) c->largest_call = n_frames;
    *out_frames = produced;
    return IQGPU_OK;
}

// what the last destroyed stub chain saw
void stub_last_run(uint64_t* calls, uint64_t* largest_call) { *calls = g_calls; *largest_call = g_largest; }

}  // extern "C"
