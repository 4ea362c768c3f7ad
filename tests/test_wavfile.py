"""WAV / RF64 containers around the path (reference src/input_wav.c, src/output_wav_common.c).

CPU: the product's container reader, auxi / file-name metadata parsers, centre-target shift and header writer
(all host code behind the C ABI) against the restated reference logic in oracle/wav_meta.py — whose XML half is
pyexpat, the library the reference itself links — and against the standard library's `wave` module as an
independent RIFF reader / writer.  GPU: whole file runs through the chain.
"""
import dataclasses
import os
import struct
import wave

import numpy as np
import pytest

from iq_tool_b200 import gpu as G
from iq_tool_b200.configs import FORMAT_CODES, NUMPY_DTYPE
from iq_tool_b200.synth import synth_numpy
from oracle import wav_meta as O

CHUNK = 16384


# ------------------------------------------------------------------------------------------------ file builders
def chunk(cid: bytes, body: bytes, declared=None) -> bytes:
    size = len(body) if declared is None else declared
    return cid + struct.pack("<I", size) + body + (b"\0" if len(body) & 1 else b"")


def fmt_chunk(rate, bits, channels=2, tag=1, extensible=False) -> bytes:
    frame = channels * bits // 8
    base = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, rate, rate * frame, frame, bits)
    if extensible:   # cbSize 22, valid bits, channel mask, sub-format GUID whose first word is the real tag
        base += struct.pack("<HHI", 22, bits, 3) + struct.pack("<H", tag) + bytes.fromhex("000000001000800000aa00389b71")
    return chunk(b"fmt ", base)


def riff(chunks: bytes, magic=b"RIFF", size=None) -> bytes:
    return magic + struct.pack("<I", len(chunks) + 4 if size is None else size) + b"WAVE" + chunks


SDRC_XML = (b'<?xml version="1.0"?>\r\n<SDR-XML-Root xml:lang="EN" Description="Saved recording data" Created="04-Aug-2015 20:56">\r\n'
            b'<Definition CurrentTimeUTC="04-08-2015 20:56:28" Filename="x.wav" RadioModel="Airspy &amp; SpyVerter" '
            b"SoftwareName='SDR Console' SoftwareVersion=\"Version 3.0 build 1\" RadioCenterFreq=\"97300000\" "
            b'SampleRate="2000000" UTCSeconds="1438721788"/>\r\n</SDR-XML-Root>\r\n')


def sdruno_auxi(year=2021, month=3, day=9, hour=17, minute=5, sec=59, freq=14_074_000) -> bytes:
    start = struct.pack("<8H", year, month, 2, day, hour, minute, sec, 250)
    stop = struct.pack("<8H", year, month, 2, day, hour, minute + 1, sec, 0)
    return start + stop + struct.pack("<I", freq) + struct.pack("<IIII", 2_000_000, 0, 0, 0) + b"\0" * 100


def md_of(info: G.WavInfoC) -> dict:
    """The SdrMetadata part of a probe result in the oracle's dict form."""
    d = info.as_dict()
    return {k: v for k, v in d.items() if k in ("source_software", "center_freq_hz", "timestamp_unix", "timestamp_str",
                                                 "software_name", "software_version", "radio_model")}


# ------------------------------------------------------------------------------------------------ container reader
@pytest.mark.parametrize("fmt,width", [("cs16", 2), ("cu8", 1)])
def test_probe_agrees_with_the_stdlib_wave_module(fmt, width, tmp_path):
    rng = np.random.default_rng(7)
    payload = rng.integers(0, 256, size=5000 * 2 * width, dtype=np.uint8).tobytes()
    path = tmp_path / "plain.wav"
    with wave.open(str(path), "wb") as w:          # an independent writer
        w.setnchannels(2); w.setsampwidth(width); w.setframerate(2_400_000)
        w.writeframes(payload)
    info = G.wav_probe(str(path))
    ref = O.read_wav_header(str(path))             # an independent reader
    assert info.container == G.CONTAINER_WAV and info.sample_format == FORMAT_CODES[fmt]
    assert (info.channels, info.bits_per_sample, info.sample_rate_hz, info.frames) == \
           (ref["channels"], ref["bits_per_sample"], ref["sample_rate_hz"], ref["frames"])
    assert info.data_offset == 44 and info.data_bytes == len(payload)
    assert open(path, "rb").read()[info.data_offset:info.data_offset + info.data_bytes] == ref["payload"] == payload
    assert not info.metadata_present and info.source_software == O.SDR_SOFTWARE_UNKNOWN


def test_probe_sdr_console_capture(tmp_path):
    """auxi in front of the data (odd length, NUL padded like SDR Console does), a LIST chunk behind it."""
    payload = bytes(range(256)) * 64
    auxi = SDRC_XML + b"\0" * (36 + len(SDRC_XML) % 2 + 1 - 2 * (len(SDRC_XML) % 2))
    assert len(auxi) & 1
    blob = riff(fmt_chunk(2_000_000, 16) + chunk(b"auxi", auxi) + chunk(b"data", payload) + chunk(b"LIST", b"INFOISFT\x04\0\0\0abc\0"))
    path = tmp_path / "04-Aug-2015 205628.000 97.300MHz.wav"
    path.write_bytes(blob)
    info = G.wav_probe(str(path))
    md = O.new_metadata()
    assert O.parse_auxi(auxi, md)
    O.parse_filename(path.name, md)
    assert md_of(info) == md and info.metadata_present
    assert md["source_software"] == O.SDR_CONSOLE and md["center_freq_hz"] == 97.3e6 and md["timestamp_unix"] == 1438721788
    assert md["radio_model"] == "Airspy & SpyVerter"
    magic, chunks = O.walk_chunks(str(path))
    data = [c for c in chunks if c[0] == b"data"][0]
    assert (info.data_offset, info.data_bytes, info.frames) == (data[1], len(payload), len(payload) // 4)


def test_probe_sdruno_capture_binary_auxi_and_name(tmp_path):
    auxi = sdruno_auxi()
    path = tmp_path / "SDRuno_20210309_170559Z_14074kHz.wav"
    path.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"auxi", auxi) + chunk(b"data", b"\1\0\2\0" * 100)))
    info = G.wav_probe(str(path))
    md = O.new_metadata()
    assert O.parse_auxi(auxi, md)
    assert O.parse_filename(path.name, md)          # "14074k" is not a number: only the SDRuno_ prefix counts
    assert md_of(info) == md
    assert md["center_freq_hz"] == 14_074_000.0 and md["source_software"] == O.SDR_UNO and md["software_name"] == "SDRuno"
    assert md["timestamp_str"] == "2021-03-09 17:05:59 UTC"


def test_probe_rf64_extensible_and_metadata_behind_the_data(tmp_path):
    payload = b"\x80\x7f" * 3001                      # cu8, 3001 frames
    ds64 = struct.pack("<QQQI", 0, len(payload), 3001, 0)
    blob = riff(chunk(b"ds64", ds64) + fmt_chunk(8_000_000, 8, extensible=True) + chunk(b"data", payload)[:4] + struct.pack("<I", 0xFFFFFFFF)
                + payload + chunk(b"auxi", sdruno_auxi(freq=433_920_000)), magic=b"RF64", size=0xFFFFFFFF)
    path = tmp_path / "big.wav"
    path.write_bytes(blob)
    info = G.wav_probe(str(path))
    assert info.container == G.CONTAINER_RF64 and info.sample_format == FORMAT_CODES["cu8"] and info.format_tag == 1
    assert (info.sample_rate_hz, info.frames, info.data_bytes) == (8_000_000, 3001, len(payload))
    assert blob[info.data_offset:info.data_offset + info.data_bytes] == payload
    assert info.center_freq_hz_present and info.center_freq_hz == 433_920_000.0


@pytest.mark.parametrize("riff_size,data_size", [(0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (36, 0), (None, 10_000_000)])
def test_probe_unfinished_or_truncated_recording_runs_to_the_end_of_the_file(riff_size, data_size, tmp_path):
    payload = b"\1\2\3\4" * 777 + b"\5\6\7"          # a torn last frame
    blob = riff(fmt_chunk(1_000_000, 16) + b"data" + struct.pack("<I", data_size) + payload, size=riff_size)
    path = tmp_path / "torn.wav"
    path.write_bytes(blob)
    info = G.wav_probe(str(path))
    assert (info.data_offset, info.frames, info.data_bytes) == (44, 777, 777 * 4)


@pytest.mark.parametrize("blob,needle", [
    (riff(fmt_chunk(48_000, 16, channels=1) + chunk(b"data", b"\0\0" * 8)), "2 channels"),
    (riff(fmt_chunk(48_000, 24) + chunk(b"data", b"\0" * 12)), "unsupported PCM subtype"),
    (riff(fmt_chunk(48_000, 32, tag=3) + chunk(b"data", b"\0" * 16)), "unsupported PCM subtype"),
    (riff(fmt_chunk(0, 16) + chunk(b"data", b"\0" * 16)), "sample rate"),
    (riff(chunk(b"data", b"\0" * 16)), "fmt"),
    (riff(fmt_chunk(48_000, 16)), "no data chunk"),
    (b"RIFX" + b"\0" * 60, "not a RIFF"),
    (b"RIFF", "too short"),
])
def test_probe_refuses_what_the_reference_refuses(blob, needle, tmp_path):
    path = tmp_path / "bad.wav"
    path.write_bytes(blob)
    with pytest.raises(G.IqGpuError, match=needle):
        G.wav_probe(str(path))
    with pytest.raises(G.IqGpuError):
        G.wav_probe(str(tmp_path / "missing.wav"))


# ------------------------------------------------------------------------------------------------ metadata parsers
AUXI_CASES = {
    "sdr_console": SDRC_XML,
    "nul_padded": SDRC_XML + b"\0" * 64,
    "seconds_then_string": b'<r><Definition UTCSeconds="100" CurrentTimeUTC="01-01-2001 00:00:10"/></r>',
    "string_then_seconds": b'<r><Definition CurrentTimeUTC="01-01-2001 00:00:10" UTCSeconds="100"/></r>',
    "two_definitions": b'<r><Definition RadioCenterFreq="1e6"/><x/><Definition RadioCenterFreq="2.5e6" RadioModel="B"/></r>',
    "definition_as_root": b'<Definition SoftwareName="SDR Console V3" RadioCenterFreq=" 7100000.5"/>',
    "bad_numbers": b'<r><Definition RadioCenterFreq="97.3e6Hz" UTCSeconds="12x" SoftwareVersion="only a version"/></r>',
    "trailing_space_number": b'<r><Definition RadioCenterFreq="97.3e6 " RadioModel="R"/></r>',
    "nonfinite": b'<r><Definition RadioCenterFreq="inf" SoftwareName="n"/></r>',
    "long_strings": b'<r><Definition SoftwareName="' + b"N" * 200 + b'" RadioModel="' + b"M" * 300 + b'" SoftwareVersion="' + b"V" * 64 + b'"/></r>',
    "entities": b'<r><Definition RadioModel="A&lt;B&gt;&quot;C&apos;&#65;&#x42;" SoftwareName=\'q"q\'/></r>',
    "duplicate_attribute": b'<r><Definition RadioCenterFreq="1" RadioCenterFreq="2"/></r>',
    "error_after_first": b'<r><Definition RadioCenterFreq="5"/><Definition RadioModel="unterminated></r>',
    "error_before": b'<r><broken attr=novalue/><Definition RadioCenterFreq="5"/></r>',
    "junk_in_front": b'junk<r><Definition RadioCenterFreq="5"/></r>',
    "second_root": b'<r/><Definition RadioCenterFreq="5"/>',
    "comment_and_pi": b'\xef\xbb\xbf<?xml version="1.0" encoding="UTF-8"?><!-- c --><r><?pi x?><![CDATA[<Definition RadioCenterFreq="9"/>]]><Definition RadioCenterFreq="5"/></r>',
    "undefined_entity": b'<r><Definition RadioModel="&nbsp;"/></r>',
    "wrong_element": b'<r><definition RadioCenterFreq="5"/></r>',
    "time_only_string": b'<r><Definition CurrentTimeUTC="not a time"/></r>',
    "latin1_declared": b'<?xml version="1.0" encoding="ISO-8859-1"?><r><Definition RadioModel="caf\xe9 \xb5" RadioCenterFreq="7"/></r>',
    "latin1_undeclared": b'<r><Definition RadioModel="caf\xe9" RadioCenterFreq="7"/></r>',
    "utf8_declared_standalone": b'<?xml version="1.0" encoding="UTF-8" standalone="yes"?><r><Definition RadioModel="caf\xc3\xa9"/></r>',
    "utf16_declared_in_bytes": b'<?xml version="1.0" encoding="UTF-16"?><r><Definition RadioCenterFreq="7"/></r>',
    "declaration_not_first": b' <?xml version="1.0"?><r><Definition RadioCenterFreq="7"/></r>',
    "mismatched_end_tag": b'<r><a></b><Definition RadioCenterFreq="7"/></r>',
    "comment_with_double_dash": b'<r><!-- a -- b --><Definition RadioCenterFreq="7"/></r>',
    "non_ascii_names": b'<r \xc3\xa9l\xc3\xa9ment="1"><Definition RadioCenterFreq="7" \xc3\x97="bad"/></r>',
    "sdruno_binary": sdruno_auxi(),
    "binary_no_freq": sdruno_auxi(freq=0),
    "binary_short": sdruno_auxi()[:35],
    "binary_minimal": sdruno_auxi()[:36],
    "empty_like": b" ",
}


@pytest.mark.parametrize("name", sorted(AUXI_CASES))
def test_auxi_chunk_parser_equals_the_restated_reference(name):
    blob = AUXI_CASES[name]
    md = O.new_metadata()
    want = O.parse_auxi(blob, md)
    got, info = G.wav_parse_auxi(blob)
    assert got == want, name
    assert md_of(info) == md, name


FILENAMES = [
    "SDRSharp_20150804_205628Z_97300000Hz_IQ.wav", "SDRSharp_20150804_205628Z_97300kHz_IQ.wav", "baseband_1090000000Hz.wav",
    "rec_1.0905e9hz_x.wav", "HDSDR_20200101_000000Z_7100kHz_RF.wav", "_20200229_235959Z.wav", "x_20200229_235959Z_a_20210101_000000Z.wav",
    "SDRuno_capture.wav", "SDRconnect_IQ_20240102_030405_7000000HZ.wav", "SDRconnect_.wav", "gqrx_20150804_205628_97300000_2000000_fc.raw",
    "nothing.wav", "Hz.wav", "_Hz.wav", "a_-5Hz.wav", "a_0x10Hz.wav", "a_infHz.wav", "a_ 12Hz.wav", "a_12 Hz.wav",
    "a_" + "1" * 31 + "Hz.wav", "a_" + "1" * 32 + "Hz.wav", "two_100Hz_200Hz.wav", "_2020022_9235959Z.wav", "",
]


@pytest.mark.parametrize("name", FILENAMES)
def test_filename_parser_equals_the_restated_reference(name):
    md = O.new_metadata()
    want = O.parse_filename(name, md)
    got, info = G.wav_parse_filename(name)
    assert got == want, name
    assert md_of(info) == md, name


def test_filename_fills_only_what_the_chunk_left_open():
    """wav_initialize order: auxi first, then the name (src/input_wav.c:600-608)."""
    md = O.new_metadata()
    O.parse_auxi(SDRC_XML, md)
    O.parse_filename("SDRSharp_20200101_000000Z_1000000Hz_IQ.wav", md)
    _, info = G.wav_parse_auxi(SDRC_XML)
    G.wav_parse_filename("SDRSharp_20200101_000000Z_1000000Hz_IQ.wav", info)
    assert md_of(info) == md and md["center_freq_hz"] == 97.3e6 and md["source_software"] == O.SDR_CONSOLE


def test_center_target_shift():
    _, info = G.wav_parse_auxi(SDRC_XML)
    md = O.new_metadata()
    O.parse_auxi(SDRC_XML, md)
    for target in (97.4e6, 97_350_001.0, 1.0, -5.0e5):       # the option is a float: 97 350 001 is not representable
        assert G.wav_center_target_shift(info, target) == O.center_target_shift(md, target)
    assert G.wav_center_target_shift(info, 97_350_001.0) == 97.3e6 - 97_350_000.0
    assert G.wav_center_target_shift(info, 0.0, 1234.5) == 1234.5                 # option not given: --freq-shift stands
    with pytest.raises(G.IqGpuError, match="Conflicting"):
        G.wav_center_target_shift(info, 97.4e6, 1000.0)
    with pytest.raises(G.IqGpuError, match="center frequency metadata"):
        G.wav_center_target_shift(G.WavInfoC(), 97.4e6)


# ------------------------------------------------------------------------------------------------ header writer
@pytest.mark.parametrize("fmt,width", [("cs16", 2), ("cu8", 1)])
def test_wav_header_is_read_back_by_the_stdlib_wave_module_and_by_the_probe(fmt, width, tmp_path):
    payload = np.random.default_rng(3).integers(0, 256, size=2 * width * 1234, dtype=np.uint8).tobytes()
    hdr = G.wav_build_header(G.CONTAINER_WAV, FORMAT_CODES[fmt], 744187, len(payload))
    assert len(hdr) == 44
    path = tmp_path / "out.wav"
    path.write_bytes(hdr + payload)
    ref = O.read_wav_header(str(path))
    assert (ref["channels"], ref["bits_per_sample"], ref["sample_rate_hz"], ref["frames"]) == (2, 8 * width, 744187, 1234)
    assert ref["payload"] == payload
    # byte for byte what the stdlib writer produces for the same stream (the canonical 44-byte PCM header)
    with wave.open(str(tmp_path / "std.wav"), "wb") as w:
        w.setnchannels(2); w.setsampwidth(width); w.setframerate(744187); w.writeframes(payload)
    assert (tmp_path / "std.wav").read_bytes()[:44] == hdr
    info = G.wav_probe(str(path))
    assert (info.sample_format, info.sample_rate_hz, info.frames, info.data_offset) == (FORMAT_CODES[fmt], 744187, 1234, 44)


def test_rf64_header_layout_and_round_trip(tmp_path):
    n = 6 * 2**30 + 4                                  # past 4 GiB: only the ds64 chunk can say so
    hdr = G.wav_build_header(G.CONTAINER_RF64, FORMAT_CODES["cs16"], 10_000_000, n)
    assert len(hdr) == 80 and hdr[:4] == b"RF64" and hdr[8:16] == b"WAVEds64"
    assert struct.unpack_from("<I", hdr, 4)[0] == 0xFFFFFFFF and struct.unpack_from("<I", hdr, 16)[0] == 28
    riff_size, data_size, frames, table = struct.unpack_from("<QQQI", hdr, 20)
    assert (riff_size, data_size, frames, table) == (n + 72, n, n // 4, 0)
    assert hdr[48:52] == b"fmt " and struct.unpack_from("<IHHIIHH", hdr, 52) == (16, 1, 2, 10_000_000, 40_000_000, 4, 16)
    assert hdr[72:76] == b"data" and struct.unpack_from("<I", hdr, 76)[0] == 0xFFFFFFFF
    small = b"\1\0\2\0" * 50
    path = tmp_path / "small.rf64"
    path.write_bytes(G.wav_build_header(G.CONTAINER_RF64, FORMAT_CODES["cs16"], 10_000_000, len(small)) + small)
    info = G.wav_probe(str(path))
    assert (info.container, info.frames, info.data_offset, info.data_bytes) == (G.CONTAINER_RF64, 50, 80, 200)
    # a plain WAV cannot say more than 4 GiB: all ones, "to the end of the file"
    big = G.wav_build_header(G.CONTAINER_WAV, FORMAT_CODES["cu8"], 1_000_000, n)
    assert struct.unpack_from("<I", big, 4)[0] == struct.unpack_from("<I", big, 40)[0] == 0xFFFFFFFF


def test_header_writer_refuses_what_the_output_module_refuses():
    with pytest.raises(G.IqGpuError, match="cs16"):    # wav_common_validate_options
        G.wav_build_header(G.CONTAINER_WAV, FORMAT_CODES["cf32"], 48_000, 0)
    with pytest.raises(G.IqGpuError):
        G.wav_build_header(G.CONTAINER_RAW, FORMAT_CODES["cs16"], 48_000, 0)
    with pytest.raises(G.IqGpuError):
        G.wav_build_header(G.CONTAINER_WAV, FORMAT_CODES["cs16"], 0, 0)


def test_wavfile_run_reports_container_errors_before_it_needs_a_device(tmp_path, workloads):
    cfg = workloads["cfg1"].config
    mono = tmp_path / "mono.wav"
    mono.write_bytes(riff(fmt_chunk(2_000_000, 16, channels=1) + chunk(b"data", b"\0\0" * 64)))
    with pytest.raises(G.IqGpuError, match="2 channels"):
        G.wavfile_run(cfg, str(mono), str(tmp_path / "o.wav"))
    ok = tmp_path / "ok.wav"
    ok.write_bytes(riff(fmt_chunk(2_000_000, 16) + chunk(b"data", b"\0\0\0\0" * 64)))
    with pytest.raises(G.IqGpuError, match="center frequency metadata"):
        G.wavfile_run(dataclasses.replace(cfg, freq_shift_hz=0.0), str(ok), str(tmp_path / "o.wav"), center_target_hz=1e6)
    with pytest.raises(G.IqGpuError, match="cs16"):     # cf32 cannot go into a WAV container
        G.wavfile_run(dataclasses.replace(cfg, output_format="cf32"), str(ok), str(tmp_path / "o.wav"))
    if G.device_count() < 1:
        with pytest.raises(G.IqGpuError):               # no CPU fallback behind the container either
            G.wavfile_run(cfg, str(ok), str(tmp_path / "o.wav"))
        assert not (tmp_path / "o.wav").exists()


# ------------------------------------------------------------------------------------------------ whole file runs
@pytest.mark.gpu
@pytest.mark.parametrize("out_container", [G.CONTAINER_WAV, G.CONTAINER_RF64])
def test_wav_capture_through_the_chain(out_container, tmp_path, gpu, workloads):
    """WAV in (auxi in front, LIST behind the data: only the data chunk may reach the chain; format and rate come
    from the header, not from the configuration) -> container out.  The payload equals one in-memory chain call
    byte for byte and the CPU oracle within the integer bar; the header carries (int)target_rate and the frame count."""
    from oracle.loader import CpuChain
    wl = workloads["cfg1"]
    cfg = wl.config
    frames = 37 * CHUNK + 1234
    raw = synth_numpy(wl, frames)
    src, dst = tmp_path / "SDRSharp_20150804_205628Z_97300000Hz_IQ.wav", tmp_path / "out.wav"
    src.write_bytes(riff(fmt_chunk(int(cfg.input_rate_hz), 16) + chunk(b"auxi", SDRC_XML) + chunk(b"data", raw.tobytes())
                         + chunk(b"LIST", b"INFO" + b"\x55" * 100_000)))
    wrong = dataclasses.replace(cfg, input_format="cu8", input_rate_hz=1.0e6)      # the header must win
    st, info = gpu.wavfile_run(wrong, str(src), str(dst), out_container=out_container, train_chunks=8)
    assert st.frames_in == frames == info.frames and info.center_freq_hz == 97.3e6
    one = gpu.Chain(cfg, 0).process(raw)
    blob = dst.read_bytes()
    hb = 44 if out_container == G.CONTAINER_WAV else 80
    out = np.frombuffer(blob[hb:], dtype=NUMPY_DTYPE[cfg.output_format])
    assert st.bytes_written == len(blob) - hb and st.frames_out * 2 == out.size
    assert np.array_equal(out, one)
    assert blob[:hb] == gpu.wav_build_header(out_container, FORMAT_CODES[cfg.output_format], int(cfg.target_rate_hz), st.bytes_written)
    back = gpu.wav_probe(str(dst))
    assert (back.frames, back.sample_rate_hz, back.sample_format) == (st.frames_out, int(cfg.target_rate_hz), FORMAT_CODES[cfg.output_format])
    if out_container == G.CONTAINER_WAV:
        ref = O.read_wav_header(str(dst))
        assert ref["frames"] == st.frames_out and ref["payload"] == out.tobytes()
    cpu = CpuChain(cfg, "oracle").process(raw)
    assert out.size == cpu.size and int(np.abs(out.astype(np.int32) - cpu.astype(np.int32)).max()) <= 1


@pytest.mark.gpu
def test_center_target_frequency_becomes_the_chain_shift(tmp_path, gpu, workloads):
    """--wav-center-target-freq: the recorded centre (auxi) minus the target is the NCO shift of the run."""
    wl = workloads["cfg1"]
    base = dataclasses.replace(wl.config, freq_shift_hz=0.0)
    raw = synth_numpy(wl, 20 * CHUNK)
    src = tmp_path / "capture.wav"
    src.write_bytes(riff(fmt_chunk(int(base.input_rate_hz), 16) + chunk(b"auxi", SDRC_XML) + chunk(b"data", raw.tobytes())))
    dst = tmp_path / "out.raw"
    st, _ = gpu.wavfile_run(base, str(src), str(dst), out_container=G.CONTAINER_RAW, center_target_hz=97.4e6)
    out = np.fromfile(dst, dtype=NUMPY_DTYPE[base.output_format])
    want = gpu.Chain(dataclasses.replace(base, freq_shift_hz=97.3e6 - 97.4e6), 0).process(raw)
    unshifted = gpu.Chain(base, 0).process(raw)
    assert np.array_equal(out, want) and not np.array_equal(out, unshifted)


@pytest.mark.gpu
def test_raw_capture_into_a_wav_container(tmp_path, gpu, workloads):
    wl = workloads["cfg5"]
    cfg = dataclasses.replace(wl.config, output_format="cs16") if wl.config.output_format not in ("cs16", "cu8") else wl.config
    raw = synth_numpy(wl, 50 * CHUNK + 77)
    src, dst = tmp_path / "capture.cs16", tmp_path / "out.wav"
    src.write_bytes(raw.tobytes())
    st, _ = gpu.wavfile_run(cfg, str(src), str(dst), in_container=G.CONTAINER_RAW, out_container=G.CONTAINER_WAV)
    ref = O.read_wav_header(str(dst))
    one = gpu.Chain(cfg, 0).process(raw)
    assert ref["sample_rate_hz"] == int(cfg.target_rate_hz) and ref["frames"] == st.frames_out == one.size // 2
    assert ref["payload"] == one.tobytes()


def test_probe_rf64_with_an_absurd_ds64_size_is_clipped_to_the_file(tmp_path):
    payload = b"\1\0\2\0" * 321
    ds64 = struct.pack("<QQQI", 2**63, 2**64 - 8, 2**62, 0)
    blob = riff(chunk(b"ds64", ds64) + fmt_chunk(1_000_000, 16) + b"data" + struct.pack("<I", 0xFFFFFFFF) + payload, magic=b"RF64", size=0xFFFFFFFF)
    path = tmp_path / "absurd.rf64"
    path.write_bytes(blob)
    info = G.wav_probe(str(path))
    assert (info.frames, info.data_bytes, info.data_offset) == (321, len(payload), len(blob) - len(payload))


def test_xml_scanner_against_expat_on_mutated_documents():
    """Differential check of the product's XML scanner against pyexpat (the reference's parser): 5000 seeded
    mutations of an SDR Console chunk — byte flips, deletions, insertions of markup characters / NULs / broken
    UTF-8, spliced elements, comments, references, CDATA, processing instructions, byte-order marks."""
    import random
    rng = random.Random(20261017)
    alphabet = b"<>&;\"'=/ ?!-[]ab1#x\0\n\t" + bytes([0xC3, 0xA9, 0xFF])
    fragments = [b'<Definition RadioCenterFreq="5"/>', b"<!-- x -->", b"&amp;", b"&#x41;", b'<a b="c">', b"</a>", b"<?p?>",
                 b"<![CDATA[z]]>", b' RadioModel="Q"', b"\xef\xbb\xbf", b'<?xml version="1.0" encoding="ISO-8859-1"?>', b"\xe9"]
    for it in range(5000):
        b = bytearray(SDRC_XML)
        for _ in range(rng.randint(1, 3)):
            op, pos = rng.randint(0, 3), rng.randrange(len(b))
            if op == 0:
                b[pos] = rng.choice(alphabet)
            elif op == 1:
                del b[pos:pos + rng.randint(1, 6)]
            elif op == 2:
                b.insert(pos, rng.choice(alphabet))
            else:
                b[pos:pos] = rng.choice(fragments)
            if not b:
                b = bytearray(b"<")
        blob = bytes(b)
        md = O.new_metadata()
        want = O.parse_auxi(blob, md)
        got, info = G.wav_parse_auxi(blob)
        assert got == want and md_of(info) == md, (it, blob)


def test_filename_parser_against_the_restated_reference_on_random_names():
    """4000 seeded names from the alphabet the two patterns care about (digits, '_', 'Z', 'Hz' in both cases, signs,
    blanks, exponents), with valid fragments spliced in; the oracle scans with libc's sscanf / strtod / timegm."""
    import random
    rng = random.Random(7)
    alphabet = "0123456789__ZzHh.e+- k"
    pieces = ["_20150804_205628Z", "_97300000Hz", "SDRuno_", "SDRconnect_", "_1.0905e9hz", "_20201301_256199Z", "_IQ", ".wav", "_+2015-804_ 5 628Z"]
    for it in range(4000):
        n = rng.randint(0, 30)
        name = "".join(rng.choice(alphabet) for _ in range(n))
        for _ in range(rng.randint(0, 2)):
            pos = rng.randint(0, len(name))
            name = name[:pos] + rng.choice(pieces) + name[pos:]
        md = O.new_metadata()
        want = O.parse_filename(name, md)
        got, info = G.wav_parse_filename(name)
        assert got == want and md_of(info) == md, (it, name)
