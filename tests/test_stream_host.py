"""CPU test of the host-side streaming plumbing around the path (iq_tool_b200/csrc/rawfile.cpp, wavfile.cpp:
reader thread / chain calls / writer thread) against a TEST DOUBLE of the chain
(tests/native/stream_stub.cpp: keeps every second frame, phase carried across calls).  The product library is not
involved: the two host sources are compiled with the stub into a scratch library.  What is checked is what the GPU
cannot tell apart from a kernel bug: chunk-train cuts, the read limit of a data chunk, torn last frames, header
patching, error reporting."""
import ctypes as C
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

from iq_tool_b200.configs import FORMAT_CODES, ChainConfig
from iq_tool_b200.gpu import RawfileStatsC, WavInfoC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHUNK = 16384


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("streamstub") / "libstreamstub.so")
    csrc = os.path.join(ROOT, "iq_tool_b200", "csrc")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wall", "-o", out,
                    os.path.join(ROOT, "tests", "native", "stream_stub.cpp"), os.path.join(csrc, "rawfile.cpp"),
                    os.path.join(csrc, "wavfile.cpp")], check=True)
    lib = C.CDLL(out)
    lib.iqgpu_rawfile_last_error.restype = C.c_char_p
    return lib


def cfg_c(fmt="cs16"):
    return ChainConfig(input_format=fmt, output_format=fmt, input_rate_hz=2.0e6, target_rate_hz=1.0e6).to_c()


def last_run(lib):
    calls, largest = C.c_uint64(), C.c_uint64()
    lib.stub_last_run(C.byref(calls), C.byref(largest))
    return calls.value, largest.value


def riff(chunks: bytes) -> bytes:
    return b"RIFF" + struct.pack("<I", len(chunks) + 4) + b"WAVE" + chunks


def chunk(cid: bytes, body: bytes) -> bytes:
    return cid + struct.pack("<I", len(body)) + body + (b"\0" if len(body) & 1 else b"")


def fmt_chunk(rate, bits):
    return chunk(b"fmt ", struct.pack("<HHIIHH", 1, 2, rate, rate * bits // 4, bits // 4, bits))


@pytest.mark.parametrize("frames,train_chunks,stray", [(0, 4, 0), (1, 4, 0), (4 * CHUNK, 4, 0), (4 * CHUNK + 1, 4, 3),
                                                        (11 * CHUNK + 4321, 3, 1), (5 * CHUNK - 1, 0, 0)])
def test_rawfile_stream_cuts_trains_and_drops_a_torn_frame(stub, frames, train_chunks, stray, tmp_path):
    rng = np.random.default_rng(frames + 1)
    raw = rng.integers(-32768, 32767, size=2 * frames, dtype=np.int16)
    src, dst = tmp_path / "in.cs16", tmp_path / "out.cs16"
    src.write_bytes(raw.tobytes() + b"\x7f" * stray)
    st, c = RawfileStatsC(), cfg_c()
    rc = stub.iqgpu_rawfile_run(C.byref(c), 0, os.fsencode(src), os.fsencode(dst), C.c_size_t(train_chunks), C.byref(st))
    assert rc == 0, stub.iqgpu_rawfile_last_error()
    out = np.fromfile(dst, dtype=np.int16)
    assert np.array_equal(out, raw.reshape(-1, 2)[::2].reshape(-1))
    assert (st.frames_in, st.frames_out, st.bytes_written) == (frames, (frames + 1) // 2, out.size * 2)
    calls, largest = last_run(stub)
    per_train = (train_chunks or 64) * CHUNK
    assert largest <= per_train and calls == -(-frames // per_train)      # an empty last train makes no chain call
    assert st.trains >= calls


def test_rawfile_errors_are_reported(stub, tmp_path):
    c = cfg_c()
    assert stub.iqgpu_rawfile_run(C.byref(c), 0, os.fsencode(tmp_path / "nope"), os.fsencode(tmp_path / "o"), C.c_size_t(0), None) != 0
    assert b"cannot open input" in stub.iqgpu_rawfile_last_error()
    (tmp_path / "in").write_bytes(b"\0" * 64)
    assert stub.iqgpu_rawfile_run(C.byref(c), 0, os.fsencode(tmp_path / "in"), os.fsencode(tmp_path / "no_dir" / "o"), C.c_size_t(0), None) != 0
    assert b"cannot open output" in stub.iqgpu_rawfile_last_error()
    assert stub.iqgpu_rawfile_run(C.byref(c), 1, os.fsencode(tmp_path / "in"), os.fsencode(tmp_path / "o"), C.c_size_t(0), None) == -2
    assert b"no such device" in stub.iqgpu_rawfile_last_error()


@pytest.mark.parametrize("fmt,bits,out_container,frames", [("cs16", 16, 1, 9 * CHUNK + 17), ("cu8", 8, 1, 3 * CHUNK), ("cs16", 16, 2, 2 * CHUNK + 5),
                                                           ("cs16", 16, 0, 7), ("cu8", 8, 2, 0)])
def test_wav_stream_reads_only_the_data_chunk_and_finalises_the_header(stub, fmt, bits, out_container, frames, tmp_path):
    dt = np.int16 if bits == 16 else np.uint8
    rng = np.random.default_rng(frames + bits)
    raw = rng.integers(0, 255, size=2 * frames).astype(dt)
    payload = raw.tobytes()
    # metadata in front, an odd-sized junk chunk and a big LIST chunk behind the samples; a wrong format/rate in the configuration
    blob = riff(fmt_chunk(2_000_000, bits) + chunk(b"auxi", b'<r><Definition RadioCenterFreq="1.5e6"/></r>') + chunk(b"data", payload)
                + chunk(b"junk", b"\xAA" * 33) + chunk(b"LIST", b"\x55" * (3 * CHUNK * 4)))
    src, dst = tmp_path / "rec_20200102_030405Z.wav", tmp_path / "out.bin"
    src.write_bytes(blob)
    c = cfg_c(fmt)
    c.input_rate_hz = 123.0
    st, info = RawfileStatsC(), WavInfoC()
    rc = stub.iqgpu_wavfile_run(C.byref(c), 0, os.fsencode(src), 1, os.fsencode(dst), out_container, C.c_float(0.0), C.c_size_t(2),
                                C.byref(st), C.byref(info))
    assert rc == 0, stub.iqgpu_rawfile_last_error()
    want = raw.reshape(-1, 2)[::2].reshape(-1).tobytes()
    got = dst.read_bytes()
    hb = {0: 0, 1: 44, 2: 80}[out_container]
    assert got[hb:] == want
    assert (st.frames_in, st.frames_out, st.bytes_written) == (frames, (frames + 1) // 2, len(want))
    assert (info.frames, info.sample_rate_hz, info.sample_format, info.center_freq_hz) == (frames, 2_000_000, FORMAT_CODES[fmt], 1.5e6)
    assert info.timestamp_unix == 1577934245
    if out_container == 1:
        with wave.open(str(dst), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, bits // 8, 1_000_000, (frames + 1) // 2)
            assert w.readframes(w.getnframes()) == want
        assert struct.unpack_from("<I", got, 4)[0] == len(got) - 8
    if out_container == 2:
        assert got[:4] == b"RF64" and struct.unpack_from("<QQQ", got, 20) == (len(got) - 8, len(want), (frames + 1) // 2)
    assert last_run(stub)[1] <= 2 * CHUNK


def test_wav_stream_of_a_raw_capture_and_the_shift_option(stub, tmp_path):
    raw = np.arange(2 * 1000, dtype=np.int16)
    src, dst = tmp_path / "c.cs16", tmp_path / "o.wav"
    src.write_bytes(raw.tobytes())
    c = cfg_c()
    st = RawfileStatsC()
    assert stub.iqgpu_wavfile_run(C.byref(c), 0, os.fsencode(src), 0, os.fsencode(dst), 1, C.c_float(0.0), C.c_size_t(0), C.byref(st), None) == 0
    with wave.open(str(dst), "rb") as w:
        assert w.getnframes() == 500 and w.readframes(500) == raw.reshape(-1, 2)[::2].tobytes()
    # a centre target needs a WAV input with the metadata; nothing is written when it is refused
    assert stub.iqgpu_wavfile_run(C.byref(c), 0, os.fsencode(src), 0, os.fsencode(tmp_path / "x.wav"), 1, C.c_float(1e6), C.c_size_t(0), None, None) != 0
    assert b"needs a WAV input" in stub.iqgpu_rawfile_last_error() and not (tmp_path / "x.wav").exists()


def test_wav_to_wav_without_resampling_takes_the_header_rate_from_the_file(stub, tmp_path):
    """--no-resample: setup.c:99 makes the source rate the target rate; for a WAV input that rate is only known after the
    probe, so the output header must carry the FILE's rate, not the configuration's target."""
    raw = np.arange(2 * 640, dtype=np.int16)
    src, dst = tmp_path / "in.wav", tmp_path / "out.wav"
    src.write_bytes(riff(fmt_chunk(1_234_567, 16) + chunk(b"data", raw.tobytes())))
    c = cfg_c()
    c.no_resample = 1
    c.target_rate_hz = 0.0                      # unset on the command line when --no-resample is given
    st = RawfileStatsC()
    rc = stub.iqgpu_wavfile_run(C.byref(c), 0, os.fsencode(src), 1, os.fsencode(dst), 1, C.c_float(0.0), C.c_size_t(0), C.byref(st), None)
    assert rc == 0, stub.iqgpu_rawfile_last_error()
    with wave.open(str(dst), "rb") as w:
        assert w.getframerate() == 1_234_567
