"""CPU restatement of the reference's WAV-input metadata logic — TEST INFRASTRUCTURE ONLY.

Follows src/input_wav.c of the reference (file:line under /root/reference):

  parse_auxi_xml      _parse_auxi_xml_expat + expat_start_element_handler   :345-441   (attribute table :334-342)
  parse_auxi_binary   _parse_binary_auxi_data                                :294-332
  parse_auxi          process_specific_chunk's XML-first / binary-second     :173-179
  parse_filename      parse_sdr_metadata_from_filename                       :190-271
  center_target_shift wav_initialize, --wav-center-target-freq               :612-629
  read_wav_header     what sf_open / SF_INFO hand to wav_initialize          :552-598

The XML form goes through pyexpat, i.e. the very library (expat) the reference links, so the product's own
little XML scanner is checked against the real parser.  The container itself is parsed by libsndfile in the
reference; libsndfile is absent here (**parity unpinned** for the chunk walk): `read_wav_header` uses the
standard library's `wave` module as an independent RIFF reader for plain WAV files and a direct struct walk
for RF64.  Nothing under iq_tool_b200/ may import this module.
"""
from __future__ import annotations

import ctypes
import math
import re
import struct
import wave
import xml.parsers.expat

SDR_SOFTWARE_UNKNOWN, SDR_CONSOLE, SDR_SHARP, SDR_UNO, SDR_CONNECT = range(5)
SOFTWARE_LABEL = {SDR_CONSOLE: "SDR Console", SDR_SHARP: "SDR#", SDR_UNO: "SDRuno", SDR_CONNECT: "SDRconnect"}
MAX_METADATA_CHUNK_SIZE = 1024 * 1024

_libc = ctypes.CDLL(None)
_libc.strtod.restype = ctypes.c_double
_libc.strtod.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p)]


def _strtod_full(text: str):
    """strtod that must consume the whole string (the reference's `*endptr == '\\0'` test); None otherwise."""
    raw = text.encode("utf-8")
    if b"\0" in raw:
        return None
    buf = ctypes.create_string_buffer(raw)
    end = ctypes.c_char_p()
    v = _libc.strtod(buf, ctypes.byref(end))
    consumed = ctypes.cast(end, ctypes.c_void_p).value - ctypes.addressof(buf)
    if consumed != len(raw):
        return None
    return v


class _Tm(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("tm_sec", "tm_min", "tm_hour", "tm_mday", "tm_mon", "tm_year", "tm_wday",
                                            "tm_yday", "tm_isdst")] + [("tm_gmtoff", ctypes.c_long), ("tm_zone", ctypes.c_char_p)]


_libc.timegm.restype = ctypes.c_long
_libc.timegm.argtypes = [ctypes.POINTER(_Tm)]


def _timegm(year, month, day, hour, minute, sec):
    """timegm_portable (:273-292) is mktime under TZ="": libc's timegm, including its normalisation of
    out-of-range fields (a binary parse of text bytes produces month 15422 and the like, and C accepts it)."""
    t = _Tm(tm_sec=sec, tm_min=minute, tm_hour=hour, tm_mday=day, tm_mon=month - 1, tm_year=year - 1900)
    ts = _libc.timegm(ctypes.byref(t))
    return None if ts == -1 else ts


def _sscanf_ints(text: str, fmt: bytes, n: int):
    """libc sscanf of n ints — the reference's own parser for its two time patterns, signs / spaces / short fields included."""
    raw = text.encode("utf-8", "surrogateescape")
    if b"\0" in raw:
        raw = raw.split(b"\0")[0]
    vals = [ctypes.c_int(0) for _ in range(n)]
    got = _libc.sscanf(ctypes.c_char_p(raw), ctypes.c_char_p(fmt), *[ctypes.byref(v) for v in vals])
    return got, [v.value for v in vals]


def new_metadata() -> dict:
    """SdrMetadata after init_sdr_metadata (:140-144); keys appear only once they are 'present'."""
    return {"source_software": SDR_SOFTWARE_UNKNOWN}


def parse_auxi_xml(chunk: bytes, md: dict) -> bool:
    def start(name, attrs):
        if name != "Definition":
            return
        for k, v in attrs.items():
            if k == "SoftwareName":
                md["software_name"] = v.encode("utf-8")[:63].decode("utf-8", "ignore")
            elif k == "SoftwareVersion":
                md["software_version"] = v.encode("utf-8")[:63].decode("utf-8", "ignore")
            elif k == "RadioModel":
                md["radio_model"] = v.encode("utf-8")[:127].decode("utf-8", "ignore")
            elif k == "RadioCenterFreq":
                d = _strtod_full(v)
                if d is not None and math.isfinite(d):          # errno (ERANGE) gives +-inf or a denormal: not in the test set
                    md["center_freq_hz"] = d
            elif k == "UTCSeconds":
                if "timestamp_unix" not in md and re.fullmatch(r"\s*[+-]?\d+", v):
                    md["timestamp_unix"] = int(v)
            elif k == "CurrentTimeUTC":
                md["timestamp_str"] = v.encode("utf-8")[:63].decode("utf-8", "ignore")
                got, (day, month, year, hour, minute, sec) = _sscanf_ints(v, b"%d-%d-%d %d:%d:%d", 6)
                if got == 6:
                    ts = _timegm(year, month, day, hour, minute, sec)
                    if ts is not None:
                        md["timestamp_unix"] = ts

    if not chunk:
        return False
    p = xml.parsers.expat.ParserCreate()
    p.StartElementHandler = start
    try:
        p.Parse(chunk, True)
    except xml.parsers.expat.ExpatError:
        pass                                                    # XML_Parse's status is ignored (:418)
    any_data = any(k in md for k in ("software_name", "radio_model", "center_freq_hz", "timestamp_unix"))
    if any_data and "SDR Console" in md.get("software_name", ""):
        md["source_software"] = SDR_CONSOLE
    return any_data


def parse_auxi_binary(chunk: bytes, md: dict) -> bool:
    if len(chunk) < 16 + 16 + 4:
        return False
    year, month, _dow, day, hour, minute, sec, _ms = struct.unpack_from("<8H", chunk, 0)
    got_time = got_freq = False
    ts = _timegm(year, month, day, hour, minute, sec)
    if ts is not None and ts != -1 and "timestamp_unix" not in md:
        md["timestamp_unix"] = ts
        got_time = True
        if "timestamp_str" not in md:
            md["timestamp_str"] = "%04u-%02u-%02u %02u:%02u:%02u UTC" % (year, month, day, hour, minute, sec)
    (freq,) = struct.unpack_from("<I", chunk, 32)
    if freq > 0 and "center_freq_hz" not in md:
        md["center_freq_hz"] = float(freq)
        got_freq = True
    return got_time or got_freq


def parse_auxi(chunk: bytes, md: dict) -> bool:
    if parse_auxi_xml(chunk, md):
        return True
    return parse_auxi_binary(chunk, md)


def parse_filename(base: str, md: dict) -> bool:
    something = inferred_sdrsharp = False
    if "center_freq_hz" not in md:
        hz = base.lower().find("hz")
        if hz >= 0:
            us = base.rfind("_", 0, hz)
            if us >= 0 and us + 1 < hz and hz - (us + 1) < 32:
                f = _strtod_full(base[us + 1:hz])
                if f is not None and math.isfinite(f) and f > 0:
                    md["center_freq_hz"] = f
                    something = inferred_sdrsharp = True
    if "timestamp_unix" not in md:
        for m in re.finditer(r"_", base):
            s = base[m.start():]
            if len(s.encode("utf-8", "surrogateescape")) < 17 or s[9:10] != "_" or s[16:17] != "Z":
                continue
            got, (y, mo, d, h, mi, se) = _sscanf_ints(s, b"_%4d%2d%2d_%2d%2d%2dZ", 6)
            if got != 6:
                continue
            ts = _timegm(y, mo, d, h, mi, se)
            if ts is None or ts == -1:
                continue
            md["timestamp_unix"] = ts
            if "timestamp_str" not in md:
                md["timestamp_str"] = "%04d-%02d-%02d %02d:%02d:%02d UTC" % (y, mo, d, h, mi, se)
            something = inferred_sdrsharp = True
            break
    if md["source_software"] == SDR_SOFTWARE_UNKNOWN:
        if inferred_sdrsharp:
            md["source_software"] = SDR_SHARP
        elif base.startswith("SDRuno_"):
            md["source_software"] = SDR_UNO
        elif base.startswith("SDRconnect_"):
            md["source_software"] = SDR_CONNECT
        if md["source_software"] != SDR_SOFTWARE_UNKNOWN and "software_name" not in md:
            md["software_name"] = SOFTWARE_LABEL[md["source_software"]]
            something = True
    return something


def center_target_shift(md: dict, center_target_hz: float, freq_shift_hz_arg: float = 0.0) -> float:
    """nco_shift_hz; ValueError where the reference log_fatal()s.  The option is a float (:443)."""
    target = struct.unpack("<f", struct.pack("<f", center_target_hz))[0]
    if target == 0.0:
        return freq_shift_hz_arg
    if freq_shift_hz_arg != 0.0:
        raise ValueError("conflicting shift options")
    if "center_freq_hz" not in md:
        raise ValueError("no centre frequency metadata")
    return md["center_freq_hz"] - target


def read_wav_header(path: str) -> dict:
    """channels / sample width / rate / frames / payload of a plain RIFF WAV through the stdlib `wave` reader."""
    with wave.open(path, "rb") as w:
        n = w.getnframes()
        return {"channels": w.getnchannels(), "bits_per_sample": 8 * w.getsampwidth(), "sample_rate_hz": w.getframerate(),
                "frames": n, "payload": w.readframes(n)}


def walk_chunks(path: str):
    """(id, body offset, declared size) of every top-level chunk; RF64 sizes are left as written."""
    blob = open(path, "rb").read()
    pos, out = 12, []
    ds64_data = None
    while pos + 8 <= len(blob):
        cid, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        if cid == b"ds64":
            ds64_data = struct.unpack_from("<Q", blob, pos + 16)[0]
        if cid == b"data" and size == 0xFFFFFFFF and ds64_data is not None:
            size = ds64_data
        out.append((cid, pos + 8, size))
        pos += 8 + size + (size & 1)
    return blob[:4], out
