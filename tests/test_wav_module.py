"""The drop-in WAV input module (iq_tool_b200/host/input_wav.c: the reference's `get_wav_input_module_api()` /
`wav_get_cli_options()` without libsndfile or expat) driven through its InputModuleInterface by
tests/native/wav_module_harness.c the way src/pipeline.c drives an input module: initialize, optional pre-stream
calibration, start_stream on a reader thread feeding SampleChunks through the reference's own queues, summary,
cleanup.  CPU only (the module is host code); built by `make -C iq_tool_b200/host wavmodule` where the reference
headers exist, prebuilt elsewhere."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from iq_tool_b200.configs import FORMAT_CODES
from test_wavfile import SDRC_XML, chunk, fmt_chunk, riff, sdruno_auxi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "native", "_build", "libwavmodule_harness.so")
CHUNK = 16384


class Result(C.Structure):
    _fields_ = [("initialized", C.c_int), ("input_format", C.c_int), ("samplerate", C.c_int), ("source_frames", C.c_int64),
                ("nco_shift_hz", C.c_double), ("total_frames_read", C.c_uint64), ("chunks", C.c_uint64), ("bytes", C.c_uint64),
                ("largest_chunk_frames", C.c_uint64), ("saw_last_chunk", C.c_int), ("bytes_per_pair", C.c_int),
                ("has_known_length", C.c_int), ("summary_count", C.c_int), ("summary_label", (C.c_char * 64) * 16),
                ("summary_value", (C.c_char * 128) * 16)]

    def summary(self) -> dict:
        return {bytes(self.summary_label[i]).split(b"\0")[0].decode(): bytes(self.summary_value[i]).split(b"\0")[0].decode()
                for i in range(self.summary_count)}


@pytest.fixture(scope="module")
def mod():
    if os.path.exists("/root/reference/include/app_context.h"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "iq_tool_b200", "host"), "wavmodule"], check=True, stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB):
        pytest.skip("WAV module harness not built (needs the reference headers)")
    lib = C.CDLL(LIB)
    lib.wavmod_run.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_size_t, C.POINTER(Result)]
    lib.wavmod_calibration_block.restype = C.c_long
    lib.wavmod_calibration_block.argtypes = [C.c_void_p, C.c_size_t]
    return lib


def run(mod, path, target=0.0, shift_arg=0.0, iq=0, pool=4, chunk_frames=CHUNK, cap=1 << 22):
    sink = np.zeros(cap, dtype=np.uint8)
    res = Result()
    rc = mod.wavmod_run(os.fsencode(path), target, shift_arg, iq, pool, chunk_frames, sink.ctypes.data, cap, C.byref(res))
    return rc, res, sink[:res.bytes].tobytes()


def test_module_exports_the_reference_interface(mod):
    for name in ("get_wav_input_module_api", "wav_get_cli_options", "sf_read_raw", "sf_seek", "wav_common_validate_options",
                 "wav_common_initialize", "wav_common_run_writer", "wav_common_write_chunk", "wav_common_finalize_output"):
        assert hasattr(mod, name), name


@pytest.mark.parametrize("fmt,bits,frames", [("cs16", 16, 5 * CHUNK + 4321), ("cu8", 8, 3 * CHUNK), ("cs16", 16, 1), ("cs16", 16, 0)])
def test_reader_thread_delivers_exactly_the_data_chunk(mod, fmt, bits, frames, tmp_path):
    dt = np.int16 if bits == 16 else np.uint8
    payload = np.random.default_rng(frames + bits).integers(0, 255, size=2 * frames).astype(dt).tobytes()
    path = tmp_path / "capture.wav"
    path.write_bytes(riff(fmt_chunk(2_000_000, bits) + chunk(b"auxi", SDRC_XML) + chunk(b"data", payload + b"\x01" * (1 if frames else 0))
                          + chunk(b"LIST", b"\x55" * 70_001)))
    rc, res, got = run(mod, path)
    assert rc == 0 and res.initialized and res.has_known_length
    assert (res.input_format, res.samplerate, res.source_frames, res.bytes_per_pair) == (FORMAT_CODES[fmt], 2_000_000, frames, bits // 4)
    assert got == payload and res.total_frames_read == frames and res.saw_last_chunk
    assert res.chunks == -(-frames // CHUNK) and res.largest_chunk_frames == min(frames, CHUNK)
    assert res.nco_shift_hz == 0.0


def test_summary_lines_are_the_reference_s(mod, tmp_path):
    path = tmp_path / "capture.wav"
    path.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"auxi", SDRC_XML) + chunk(b"data", b"\0\0\0\0" * 10)))
    rc, res, _ = run(mod, path)
    assert rc == 0
    assert res.summary() == {
        "Input File": str(path), "Input Format": "16-bit Signed Complex PCM (cs16)", "Input Rate": "2000000 Hz",
        "Input File Size": res.summary()["Input File Size"], "Timestamp": "2015-08-04 20:56:28 UTC",
        "Center Frequency": "97300000 Hz", "SDR Software": "SDR Console Version 3.0 build 1", "Radio Model": "Airspy & SpyVerter"}
    assert res.summary()["Input File Size"].endswith("B") or "bytes" in res.summary()["Input File Size"].lower()
    plain = tmp_path / "plain.wav"
    plain.write_bytes(riff(fmt_chunk(48_000, 8) + chunk(b"data", b"\x80\x80" * 10)))
    rc, res, _ = run(mod, plain)
    assert rc == 0 and list(res.summary()) == ["Input File", "Input Format", "Input Rate", "Input File Size"]
    assert res.summary()["Input Format"] == "8-bit Unsigned Complex PCM (cu8)"


def test_center_target_option_sets_the_shift_and_its_refusals(mod, tmp_path):
    path = tmp_path / "SDRuno_20210309_170559Z_14074kHz.wav"
    path.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"auxi", sdruno_auxi(freq=14_074_000)) + chunk(b"data", b"\1\0\2\0" * 100)))
    rc, res, _ = run(mod, path, target=14.1e6)
    assert rc == 0 and res.nco_shift_hz == 14_074_000.0 - 14_100_000.0
    assert res.summary()["Center Frequency"] == "14074000 Hz" and res.summary()["SDR Software"].strip() == "SDRuno"
    rc, res, _ = run(mod, path, target=14.1e6, shift_arg=1000.0)          # --freq-shift given as well
    assert rc == 1 and not res.initialized
    bare = tmp_path / "bare.wav"
    bare.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"data", b"\1\0\2\0" * 100)))
    rc, res, _ = run(mod, bare, target=14.1e6)                            # no centre frequency to work from
    assert rc == 1 and not res.initialized
    rc, res, _ = run(mod, bare)
    assert rc == 0 and res.nco_shift_hz == 0.0


@pytest.mark.parametrize("blob", [riff(fmt_chunk(48_000, 16, channels=1) + chunk(b"data", b"\0\0" * 8)),
                                  riff(fmt_chunk(48_000, 24) + chunk(b"data", b"\0" * 12)), b"not a wav file at all"])
def test_initialize_refuses_what_the_reference_refuses(mod, blob, tmp_path):
    path = tmp_path / "bad.wav"
    path.write_bytes(blob)
    rc, res, _ = run(mod, path)
    assert rc == 1 and not res.initialized
    rc, res, _ = run(mod, tmp_path / "missing.wav")
    assert rc == 1


def test_pre_stream_calibration_reads_the_first_block_and_rewinds(mod, tmp_path):
    frames = 2 * CHUNK + 10
    payload = np.random.default_rng(5).integers(-3000, 3000, size=2 * frames).astype(np.int16).tobytes()
    path = tmp_path / "capture.wav"
    path.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"data", payload)))
    rc, res, got = run(mod, path, iq=1)
    block = np.zeros(8192, dtype=np.uint8)
    n = mod.wavmod_calibration_block(block.ctypes.data, block.size)
    assert rc == 0 and n == 4096 and block[:n].tobytes() == payload[:4096]
    assert got == payload                                                 # the stream still starts at frame 0
    rc, res, got = run(mod, path, iq=0)
    assert rc == 0 and mod.wavmod_calibration_block(block.ctypes.data, block.size) == -2      # service not called


# ---------------------------------------------------------------------------------------------- output modules
class OutResult(C.Structure):
    _fields_ = [("validated", C.c_int), ("initialized", C.c_int), ("final_output_size_bytes", C.c_longlong),
                ("total_output_frames", C.c_ulonglong), ("progress_calls", C.c_ulonglong), ("progress_last_bytes", C.c_ulonglong),
                ("summary_count", C.c_int), ("summary_label", (C.c_char * 64) * 4), ("summary_value", (C.c_char * 128) * 4)]


def run_out(mod, path, rf64, fmt, rate, data: bytes, piece, mode):
    mod.wavout_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(OutResult)]
    res = OutResult()
    rc = mod.wavout_run(os.fsencode(path), rf64, FORMAT_CODES[fmt], rate, data, len(data), piece, mode, C.byref(res))
    return rc, res


@pytest.mark.parametrize("rf64,fmt,mode,piece", [(0, "cs16", 0, 300_001), (0, "cu8", 0, 1 << 20), (1, "cs16", 0, 77_777), (0, "cs16", 1, 65_536),
                                                 (1, "cu8", 1, 4096)])
def test_output_wrappers_of_the_reference_on_the_drop_in_writer(mod, rf64, fmt, mode, piece, tmp_path):
    """src/output_wav.c / src/output_wav_rf64.c (compiled in place, unmodified) -> drop-in wav_common_*: Writer thread
    fed through the reference's ring buffer, or direct write_chunk calls; closing patches the header."""
    import wave
    from iq_tool_b200 import gpu as G
    width = 2 if fmt == "cs16" else 1
    frames = 1_234_567 if mode == 0 else 50_000
    payload = np.random.default_rng(frames).integers(0, 256, size=2 * width * frames, dtype=np.uint8).tobytes()
    path = tmp_path / ("out.rf64" if rf64 else "out.wav")
    rc, res = run_out(mod, path, rf64, fmt, 744187.5, payload, piece, mode)
    assert rc == 0 and res.validated and res.initialized
    assert res.final_output_size_bytes == len(payload)
    blob = path.read_bytes()
    hb = 80 if rf64 else 44
    assert blob[hb:] == payload
    assert blob[:hb] == G.wav_build_header(G.CONTAINER_RF64 if rf64 else G.CONTAINER_WAV, FORMAT_CODES[fmt], 744187, len(payload))
    info = G.wav_probe(str(path))
    assert (info.frames, info.sample_rate_hz, info.sample_format) == (frames, 744187, FORMAT_CODES[fmt])
    if not rf64:
        with wave.open(str(path), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, width, 744187, frames)
    if mode == 0:       # the Writer thread reports progress per 1 MB piece, in frames of the output format
        assert res.progress_calls >= len(payload) // (1 << 20) and res.progress_last_bytes == len(payload)
        assert res.total_output_frames == frames
    label = bytes(res.summary_label[0]).split(b"\0")[0].decode(), bytes(res.summary_value[0]).split(b"\0")[0].decode()
    assert label[0] == "Output Type" and ("RF64" in label[1]) == bool(rf64)


def test_output_module_refusals(mod, tmp_path):
    rc, res = run_out(mod, tmp_path / "o.wav", 0, "cf32", 48000.0, b"", 1, 1)
    assert rc == 1 and not res.validated                                   # only cs16 / cu8 go into a WAV container
    (tmp_path / "dir.wav").mkdir()
    rc, res = run_out(mod, tmp_path / "dir.wav", 0, "cs16", 48000.0, b"", 1, 1)
    assert rc == 2 and res.validated and not res.initialized               # exists but is not a regular file
    rc, res = run_out(mod, tmp_path / "no_dir" / "o.wav", 1, "cs16", 48000.0, b"", 1, 1)
    assert rc == 2                                                         # cannot be opened
    rc, res = run_out(mod, tmp_path / "empty.wav", 0, "cu8", 48000.0, b"", 1, 0)
    assert rc == 0 and (tmp_path / "empty.wav").stat().st_size == 44 and res.final_output_size_bytes == 0


# ---------------------------------------------------------------------------------------------- raw-file input module
def run_raw(mod, path, rate, fmt, iq=0, pool=3, chunk_frames=CHUNK, cap=1 << 22):
    mod.rawmod_run.argtypes = [C.c_char_p, C.c_float, C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_size_t, C.POINTER(Result)]
    sink = np.zeros(cap, dtype=np.uint8)
    res = Result()
    rc = mod.rawmod_run(os.fsencode(path), rate, fmt.encode() if fmt is not None else None, iq, pool, chunk_frames, sink.ctypes.data, cap, C.byref(res))
    return rc, res, sink[:res.bytes].tobytes()


@pytest.mark.parametrize("fmt,pair,frames,stray", [("cs16", 4, 4 * CHUNK + 99, 3), ("CU8", 2, 2 * CHUNK, 1), ("cf32", 8, CHUNK - 1, 7), ("sc16q11", 4, 10, 0),
                                                   ("cs8", 2, 0, 1)])
def test_raw_file_module_reads_whole_frames_in_chunks(mod, fmt, pair, frames, stray, tmp_path):
    """The drop-in of src/input_rawfile.c (get_raw_file_input_module_api on a plain FILE): format names are matched
    without regard to case, the length is whole frames, a torn last frame never reaches a chunk."""
    payload = np.random.default_rng(frames + pair).integers(0, 256, size=frames * pair, dtype=np.uint8).tobytes()
    path = tmp_path / "capture.raw"
    path.write_bytes(payload + b"\xee" * stray)
    rc, res, got = run_raw(mod, path, 2.4e6, fmt)
    assert rc == 0 and res.initialized and res.has_known_length
    assert (res.input_format, res.samplerate, res.source_frames, res.bytes_per_pair) == (FORMAT_CODES[fmt.lower()], 2_400_000, frames, pair)
    assert got == payload and res.total_frames_read == frames and res.saw_last_chunk and res.chunks == -(-frames // CHUNK)
    s = res.summary()
    assert list(s) == ["Input File", "Input Type", "Input Format", "Input Rate", "Input File Size"]
    assert (s["Input Type"], s["Input Format"], s["Input Rate"]) == ("RAW FILE", fmt, "2400000 Hz")


def test_raw_file_module_refusals_and_calibration(mod, tmp_path):
    path = tmp_path / "capture.raw"
    payload = np.arange(4 * 3000, dtype=np.uint8).tobytes() * 1
    path.write_bytes(payload)
    assert run_raw(mod, path, 2.0e6, None)[0] == 7                  # --raw-file-input-sample-format missing
    assert run_raw(mod, path, 2.0e6, "cs17")[0] == 1                # unknown format name
    assert run_raw(mod, path, 2.0e6, "cs24")[0] == 1                # not a format the raw reader opens
    assert run_raw(mod, tmp_path / "missing.raw", 2.0e6, "cs16")[0] == 1
    rc, res, got = run_raw(mod, path, 2.0e6, "cs16", iq=1)
    block = np.zeros(8192, dtype=np.uint8)
    assert rc == 0 and mod.wavmod_calibration_block(block.ctypes.data, block.size) == 4096
    assert block[:4096].tobytes() == payload[:4096] and got == payload
