// rawfile.cpp — raw-file streaming around the chain: the step either side of the
// hot path for file inputs.  Replaces the reference's Reader -> ... -> Writer plumbing for raw files
// (src/input_rawfile.c:188-249: one sf_read_raw of a 16384-frame chunk per loop, src/output_raw_file.c:
// 146-184: 1 MB fwrites out of a ring buffer) with three overlapped stages on pinned memory:
//
//   reader thread   large sequential reads of whole chunk trains into a ring of pinned input buffers
//   caller thread   iqgpu_chain_process on each train (H2D / kernels / D2H pipelined inside the chain)
//   writer thread   one fwrite per train out of a ring of pinned output buffers
//
// Chunk semantics are the reference's: the file is cut into 16384-frame chunks, a short read makes a short
// last chunk, trailing bytes that do not fill a frame are dropped (input_rawfile.c:236), and nothing is
// flushed at end of stream.  Only the C ABI of include/iqgpu.h is used.
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iqgpu.h"
#include "stream_io.hpp"

namespace {

struct Slot {
    void* buf = nullptr;
    size_t bytes = 0;       // valid bytes
    bool last = false;
};

// single-producer single-consumer ring of pinned buffers
struct Ring {
    std::vector<Slot> slots;
    size_t head = 0, tail = 0, count = 0;
    std::mutex mu;
    std::condition_variable cv;
    bool aborted = false;

    Slot* acquire_free()        // producer: next slot to fill (blocks while the ring is full)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return count < slots.size() || aborted; });
        return aborted ? nullptr : &slots[head];
    }
    void publish()
    {
        std::lock_guard<std::mutex> lk(mu);
        head = (head + 1) % slots.size();
        count++;
        cv.notify_all();
    }
    Slot* acquire_full()        // consumer: oldest filled slot (blocks while the ring is empty)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return count > 0 || aborted; });
        return (count > 0) ? &slots[tail] : nullptr;
    }
    void release()
    {
        std::lock_guard<std::mutex> lk(mu);
        tail = (tail + 1) % slots.size();
        count--;
        cv.notify_all();
    }
    void abort()
    {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};

thread_local std::string g_io_err;

}  // namespace

void iqio::set_last_error(const std::string& msg) { g_io_err = msg; }

extern "C" {

const char* iqgpu_rawfile_last_error(void) { return g_io_err.c_str(); }

int iqgpu_rawfile_run(const iqgpu_chain_config* cfg, int device, const char* in_path, const char* out_path,
                      size_t train_chunks, iqgpu_rawfile_stats* stats)
{
    if (!cfg || !in_path || !out_path) { g_io_err = "null argument"; return IQGPU_EINVAL; }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!iqgpu_get_bytes_per_sample(cfg->input_format) || !iqgpu_get_bytes_per_sample(cfg->output_format)) { g_io_err = "unhandled sample format"; return IQGPU_EINVAL; }

    iqgpu_chain* chain = nullptr;
    int rc = iqgpu_chain_create(cfg, device, &chain);
    if (rc) { g_io_err = iqgpu_last_error(); return rc; }
    FILE* fin = fopen(in_path, "rb");
    if (!fin) { g_io_err = std::string("cannot open input file ") + in_path; iqgpu_chain_destroy(chain); return IQGPU_EINVAL; }
    FILE* fout = fopen(out_path, "wb");
    if (!fout) { g_io_err = std::string("cannot open output file ") + out_path; fclose(fin); iqgpu_chain_destroy(chain); return IQGPU_EINVAL; }
    rc = iqio::run_stream(chain, cfg, fin, UINT64_MAX, fout, train_chunks, stats, g_io_err);
    fclose(fin);
    fclose(fout);
    iqgpu_chain_destroy(chain);
    return rc;
}

}  // extern "C"

// One pass of an open input stream through an existing chain into an open output stream: at most `in_limit_bytes`
// are read from the current position of `fin` (UINT64_MAX = to end of file); the caller owns chain and files.
int iqio::run_stream(iqgpu_chain* chain, const iqgpu_chain_config* cfg, FILE* fin, uint64_t in_limit_bytes, FILE* fout,
                     size_t train_chunks, iqgpu_rawfile_stats* stats, std::string& err)
{
    if (train_chunks == 0) train_chunks = 64;      // 1 Mi frames per train: measured best on B200 (profiles/r02_bench_lines.jsonl, file:cfg2)
    const size_t in_bps = iqgpu_get_bytes_per_sample(cfg->input_format), out_bps = iqgpu_get_bytes_per_sample(cfg->output_format);
    if (!in_bps || !out_bps) { err = "unhandled sample format"; return IQGPU_EINVAL; }
    int rc = IQGPU_OK;
    const size_t train_frames = train_chunks * (size_t)IQGPU_CHUNK_SAMPLES;
    iqgpu_chain_set_option(chain, "subtrain_frames", (int64_t)std::min<size_t>(train_frames, (size_t)1 << 24));

    // output capacity of one train: closed form for the worst alignment + one FFT block of slack
    iqgpu_chain_info info{};
    iqgpu_chain_get_info(chain, &info);
    const double r = (cfg->no_resample || info.ratio <= 0.f) ? 1.0 : (double)info.ratio;
    const size_t out_cap_frames = (size_t)((double)train_frames * (r > 1.0 ? r : 1.0)) + 4 * IQGPU_CHUNK_SAMPLES + 2 * (size_t)info.filter_block_size + 4096;

    Ring rin, rout;
    rin.slots.resize(3);
    rout.slots.resize(3);
    bool alloc_ok = true;
    for (auto& s : rin.slots) { s.buf = iqgpu_host_alloc(train_frames * in_bps); alloc_ok &= s.buf != nullptr; }
    for (auto& s : rout.slots) { s.buf = iqgpu_host_alloc(out_cap_frames * out_bps); alloc_ok &= s.buf != nullptr; }
    auto cleanup = [&]() {
        for (auto& s : rin.slots) iqgpu_host_free(s.buf);
        for (auto& s : rout.slots) iqgpu_host_free(s.buf);
    };
    if (!alloc_ok) { err = "pinned host allocation failed (no CUDA device?)"; cleanup(); return IQGPU_ENOMEM; }

    std::string reader_err, writer_err;
    uint64_t bytes_written = 0;
    uint64_t remaining = in_limit_bytes;    // touched by the reader thread only
    // A regular file is read with several pread()s in flight per train (one thread's copy out of the page cache moves
    // ~2 GB/s — less than a tenth of what the chain takes over PCIe); anything else (a pipe) with fread.
    struct stat sb;
    const int fd = fileno(fin);
    const bool regular = fd >= 0 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode);
    off_t file_pos = regular ? ftello(fin) : 0;
    const unsigned n_read_threads = regular ? std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2)) : 1;
    auto parallel_read = [&](char* dst, size_t want, bool& io_error) -> size_t {
        const uint64_t left_in_file = (uint64_t)sb.st_size > (uint64_t)file_pos ? (uint64_t)sb.st_size - (uint64_t)file_pos : 0;
        want = (size_t)std::min<uint64_t>(want, left_in_file);
        if (!want) return 0;
        const size_t slice = ((want + n_read_threads - 1) / n_read_threads + 4095) & ~(size_t)4095;
        std::vector<std::thread> th;
        std::vector<char> bad(n_read_threads, 0);
        for (unsigned k = 0; k < n_read_threads; k++) {
            const size_t lo = (size_t)k * slice;
            if (lo >= want) break;
            const size_t len = std::min(slice, want - lo);
            th.emplace_back([&, k, lo, len] {
                size_t done = 0;
                while (done < len) {
                    const ssize_t r = pread(fd, dst + lo + done, len - done, file_pos + (off_t)(lo + done));
                    if (r <= 0) { bad[k] = 1; return; }
                    done += (size_t)r;
                }
            });
        }
        for (auto& t : th) t.join();
        for (char b : bad) io_error |= b != 0;
        file_pos += (off_t)want;
        return want;
    };
    std::thread reader([&] {
        for (;;) {
            Slot* s = rin.acquire_free();
            if (!s) return;
            const size_t full = train_frames * in_bps;
            const size_t want = (size_t)std::min<uint64_t>(full, remaining);
            bool io_error = false;
            const size_t got = !want ? 0 : (regular ? parallel_read(static_cast<char*>(s->buf), want, io_error) : fread(s->buf, 1, want, fin));
            if (io_error || (got < want && !regular && ferror(fin))) { reader_err = "read error on the input file"; rin.abort(); return; }
            remaining -= got;
            s->bytes = got - got % in_bps;      // a trailing partial frame is dropped (input_rawfile.c:236)
            s->last = got < full;
            rin.publish();
            if (s->last) return;
        }
    });
    std::thread writer([&] {
        for (;;) {
            Slot* s = rout.acquire_full();
            if (!s) return;
            if (s->bytes && fwrite(s->buf, 1, s->bytes, fout) != s->bytes) { writer_err = "write error on the output file"; rout.abort(); rin.abort(); return; }
            bytes_written += s->bytes;
            const bool last = s->last;
            rout.release();
            if (last) return;
        }
    });

    uint64_t frames_in = 0, frames_out = 0, trains = 0;
    rc = IQGPU_OK;
    for (;;) {
        Slot* in = rin.acquire_full();
        if (!in) { rc = IQGPU_EINVAL; break; }
        Slot* out = rout.acquire_free();
        if (!out) { rc = IQGPU_EINVAL; break; }
        const size_t n = in->bytes / in_bps;
        size_t produced = 0;
        if (n) {
            rc = iqgpu_chain_process(chain, in->buf, n, nullptr, 0, out->buf, out_cap_frames * out_bps, &produced, nullptr);
            if (rc) { err = iqgpu_last_error(); break; }
        }
        frames_in += n; frames_out += produced; trains++;
        out->bytes = produced * out_bps;
        out->last = in->last;
        const bool last = in->last;
        rin.release();
        rout.publish();
        if (last) break;
    }
    if (rc) { rin.abort(); rout.abort(); }
    reader.join();
    writer.join();
    // an I/O thread that failed aborted the rings, which is what stopped the loop above: report its message
    if (!reader_err.empty()) { err = reader_err; rc = IQGPU_EINVAL; }
    if (!writer_err.empty()) { err = writer_err; rc = IQGPU_EINVAL; }
    if (stats) { stats->frames_in = frames_in; stats->frames_out = frames_out; stats->bytes_written = bytes_written; stats->trains = trains; }
    cleanup();
    return rc;
}
