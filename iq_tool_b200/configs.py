"""Chain configuration (host side) — mirrors the reference's AppConfig/AppResources fields the
hot path reads (reference include/app_context.h:66-138, 205-283) as one flat C struct
(`include/iqgpu.h: iqgpu_chain_config`), plus the five BASELINE.json workloads restated the way
the reference CLI would resolve them.

Numbers that arrive through the reference CLI are parsed with strtof (src/argparse.c:110), so
rates/shifts/cutoffs are rounded to float here as well.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

# format_t numeric values, reference include/common_types.h:33-37
FORMAT_CODES = {
    "u8": 1, "s8": 2, "u16": 3, "s16": 4, "u32": 5, "s32": 6, "f32": 7,
    "cu8": 8, "cs8": 9, "cu16": 10, "cs16": 11, "cs24": 12, "cu32": 13, "cs32": 14,
    "cf32": 15, "sc16q11": 16,
}
FORMAT_NAMES = {v: k for k, v in FORMAT_CODES.items()}
# bytes per I/Q pair, reference src/sample_convert.c:102-123
BYTES_PER_SAMPLE = {"cs8": 2, "cu8": 2, "cs16": 4, "cu16": 4, "sc16q11": 4, "cs24": 6,
                    "cs32": 8, "cu32": 8, "cf32": 8}
NUMPY_DTYPE = {"cs8": np.int8, "cu8": np.uint8, "cs16": np.int16, "cu16": np.uint16,
               "sc16q11": np.int16, "cs24": np.uint8, "cs32": np.int32, "cu32": np.uint32,
               "cf32": np.float32}

FILTER_NONE, FILTER_LOWPASS, FILTER_HIGHPASS, FILTER_PASSBAND, FILTER_STOPBAND = range(5)
FILTER_REQ_AUTO, FILTER_REQ_FIR, FILTER_REQ_FFT = range(3)
FILTER_IMPL_NONE, FILTER_IMPL_FIR_SYM, FILTER_IMPL_FIR_ASYM, FILTER_IMPL_FFT_SYM, FILTER_IMPL_FFT_ASYM = range(5)
AGC_OFF, AGC_DX, AGC_LOCAL, AGC_DIGITAL = range(4)
STAGE_DC, STAGE_IQ, STAGE_NCO, STAGE_FILTER, STAGE_RESAMPLER, STAGE_AGC = 1, 2, 4, 8, 16, 32
CHUNK_SAMPLES = 16384  # PIPELINE_CHUNK_BASE_SAMPLES, reference include/constants.h:123
MAX_FILTER_CHAIN = 5


class FilterRequestC(C.Structure):
    _fields_ = [("type", C.c_int32), ("freq1_hz", C.c_float), ("freq2_hz", C.c_float)]


class ChainConfigC(C.Structure):
    """Binary layout shared by include/iqgpu.h (iqgpu_chain_config) and oracle/iq_chain_cfg.h."""
    _fields_ = [
        ("input_format", C.c_int32), ("output_format", C.c_int32),
        ("input_rate_hz", C.c_double), ("target_rate_hz", C.c_double),
        ("gain", C.c_float), ("dc_block_enable", C.c_int32),
        ("iq_correction_enable", C.c_int32), ("iq_mag", C.c_float), ("iq_phase", C.c_float),
        ("shift_after_resample", C.c_int32), ("freq_shift_hz", C.c_double),
        ("no_resample", C.c_int32), ("num_filter_requests", C.c_int32),
        ("filter_requests", FilterRequestC * MAX_FILTER_CHAIN),
        ("transition_width_hz", C.c_float), ("filter_taps", C.c_int32),
        ("attenuation_db", C.c_float), ("filter_type_request", C.c_int32),
        ("filter_fft_size", C.c_int32), ("agc_enable", C.c_int32), ("agc_profile", C.c_int32),
        ("agc_target_level_arg", C.c_float), ("stage_select", C.c_int32),
    ]


assert C.sizeof(ChainConfigC) == 160


def f32(x: float) -> float:
    return float(np.float32(x))


@dataclass
class ChainConfig:
    input_format: str = "cs16"
    output_format: str = "cs16"
    input_rate_hz: float = 2.0e6
    target_rate_hz: float = 744187.5
    gain: float = 1.0
    dc_block: bool = False
    iq_correction: bool = False
    iq_mag: float = 0.0
    iq_phase: float = 0.0
    freq_shift_hz: float = 0.0
    # False: --freq-shift, a float CLI argument (argparse.c:110); True: the double AppResources.nco_shift_hz that the WAV
    # input derives from centre-frequency metadata (input_wav.c:614-628), generally not a float
    freq_shift_is_double: bool = False
    shift_after_resample: bool = False
    no_resample: bool = False
    # (type, freq1_hz, freq2_hz) as built by reference src/config.c:192-216
    filters: List[Tuple[int, float, float]] = field(default_factory=list)
    transition_width_hz: float = 0.0
    filter_taps: int = 0
    attenuation_db: float = 0.0
    filter_type_request: int = FILTER_REQ_AUTO
    filter_fft_size: int = 0
    agc_enable: bool = False
    agc_profile: int = AGC_OFF
    agc_target_level_arg: float = 0.0
    stage_select: int = 0   # IQGPU_STAGE_* mask: module-level chain (cf32 in/out), 0 = whole chain

    @property
    def in_bytes(self) -> int:
        return BYTES_PER_SAMPLE[self.input_format]

    @property
    def out_bytes(self) -> int:
        return BYTES_PER_SAMPLE[self.output_format]

    @property
    def ratio(self) -> float:
        """float r = (float)(target/input), reference src/setup.c:107"""
        return f32(self.target_rate_hz / float(int(self.input_rate_hz)))

    def to_c(self) -> ChainConfigC:
        c = ChainConfigC()
        c.input_format = FORMAT_CODES[self.input_format]
        c.output_format = FORMAT_CODES[self.output_format]
        c.input_rate_hz = float(int(self.input_rate_hz))
        c.target_rate_hz = float(f32(self.target_rate_hz))
        c.gain = self.gain
        c.dc_block_enable = int(self.dc_block)
        c.iq_correction_enable = int(self.iq_correction)
        c.iq_mag, c.iq_phase = self.iq_mag, self.iq_phase
        c.shift_after_resample = int(self.shift_after_resample)
        c.freq_shift_hz = float(self.freq_shift_hz) if self.freq_shift_is_double else f32(self.freq_shift_hz)
        c.no_resample = int(self.no_resample)
        c.num_filter_requests = len(self.filters)
        for i, (t, a, b) in enumerate(self.filters):
            c.filter_requests[i].type = t
            c.filter_requests[i].freq1_hz = a
            c.filter_requests[i].freq2_hz = b
        c.transition_width_hz = self.transition_width_hz
        c.filter_taps = self.filter_taps
        c.attenuation_db = self.attenuation_db
        c.filter_type_request = self.filter_type_request
        c.filter_fft_size = self.filter_fft_size
        c.agc_enable = int(self.agc_enable)
        c.agc_profile = self.agc_profile if self.agc_enable else AGC_OFF
        c.agc_target_level_arg = self.agc_target_level_arg
        c.stage_select = self.stage_select
        return c


def lowpass(cutoff_hz: float):
    return (FILTER_LOWPASS, f32(cutoff_hz), 0.0)


def highpass(cutoff_hz: float):
    return (FILTER_HIGHPASS, f32(cutoff_hz), 0.0)


def pass_range(start_hz: float, end_hz: float):
    """--pass-range a:b -> centre + bandwidth in float, reference src/config.c:202-207"""
    s, e = np.float32(start_hz), np.float32(end_hz)
    bw = np.float32(e - s)
    centre = np.float32(s + np.float32(bw / np.float32(2.0)))
    return (FILTER_PASSBAND, float(centre), float(bw))


def stopband(start_hz: float, end_hz: float):
    """--stopband a:b, reference src/config.c:209-214 (centre is ignored by filter.c:237-241)"""
    s, e = np.float32(start_hz), np.float32(end_hz)
    bw = np.float32(e - s)
    centre = np.float32(s + np.float32(bw / np.float32(2.0)))
    return (FILTER_STOPBAND, float(centre), float(bw))


@dataclass
class Workload:
    name: str
    config: ChainConfig
    tones: List[Tuple[float, float]]      # (amplitude, frequency Hz)
    dc: float
    sigma: float
    parity_samples: int
    throughput_samples: int
    iq_imbalance: bool = False
    description: str = ""


def baseline_workloads() -> dict:
    """The five BASELINE.json configs, as resolved in SURVEY.md §8(d)."""
    w = {}
    # cfg1: NRSC-5 preset (iq_tool_presets.conf:216-222: cs16-fm-nrsc5 -> 744187.5 sps, agc digital),
    # output overridden to cu8, pre-resample shift -100 kHz
    w["cfg1"] = Workload(
        "cfg1",
        ChainConfig(input_format="cs16", output_format="cu8", input_rate_hz=2.0e6,
                    target_rate_hz=744187.5, freq_shift_hz=-100e3,
                    agc_enable=True, agc_profile=AGC_DIGITAL),
        tones=[(0.30, 100e3), (0.10, -150e3), (0.05, 400e3)], dc=0.0, sigma=0.01,
        parity_samples=1 << 22, throughput_samples=120_000_000,
        description="NRSC-5 preset: cs16 @ 2 Msps -> 744187.5 sps cu8, pre-resample shift, digital AGC")
    # cfg2: cs16 @ 20 Msps -> 744187.5 sps, 255-tap low-pass FIR (time domain) + DC block
    w["cfg2"] = Workload(
        "cfg2",
        ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=20.0e6,
                    target_rate_hz=744187.5, dc_block=True, filters=[lowpass(100e3)],
                    filter_taps=255, filter_type_request=FILTER_REQ_FIR),
        tones=[(0.30, 50e3), (0.10, -200e3), (0.05, 3e6)], dc=0.02, sigma=0.01,
        parity_samples=1 << 24, throughput_samples=600_000_000,
        description="cs16 @ 20 Msps -> 744187.5 sps, DC block, 255-tap low-pass FIR post-resample")
    # cfg3: cs16 @ 61.44 Msps -> 10 Msps, 4095-tap band-pass via FFT filter, post-resample shift
    w["cfg3"] = Workload(
        "cfg3",
        ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=61.44e6,
                    target_rate_hz=10e6, filters=[pass_range(1e6, 3e6)], filter_taps=4095,
                    freq_shift_hz=-2e6, shift_after_resample=True),
        tones=[(0.30, 1.5e6), (0.10, 2.5e6), (0.05, -3e6)], dc=0.0, sigma=0.01,
        parity_samples=1 << 25, throughput_samples=1_228_800_000,
        description="cs16 @ 61.44 Msps -> 10 Msps, 4095-tap complex band-pass (FFT overlap), post shift")
    # cfg4: full chain cu8 @ 2.4 Msps -> cf32 @ 1 Msps (stage API; the reference CLI cannot run it)
    w["cfg4"] = Workload(
        "cfg4",
        ChainConfig(input_format="cu8", output_format="cf32", input_rate_hz=2.4e6,
                    target_rate_hz=1e6, freq_shift_hz=50e3, dc_block=True, iq_correction=True,
                    iq_mag=-0.0476, iq_phase=-0.0524, filters=[stopband(-5e3, 5e3)],
                    agc_enable=True, agc_profile=AGC_LOCAL),
        tones=[(0.30, 200e3), (0.10, -300e3)], dc=0.03, sigma=0.01,
        parity_samples=1 << 22, throughput_samples=288_000_000, iq_imbalance=True,
        description="cu8 @ 2.4 Msps -> cf32 @ 1 Msps: shift, DC block, I/Q apply, notch FIR, LOCAL AGC")
    # cfg5: long capture @ 61.44 Msps, NRSC-5 resample + shift, time-sharded across GPUs
    w["cfg5"] = Workload(
        "cfg5",
        ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=61.44e6,
                    target_rate_hz=744187.5, freq_shift_hz=-100e3,
                    agc_enable=True, agc_profile=AGC_DIGITAL),
        tones=[(0.30, 100e3), (0.10, -150e3), (0.05, 400e3)], dc=0.0, sigma=0.01,
        parity_samples=1 << 24, throughput_samples=1 << 30,
        description="cs16 @ 61.44 Msps -> 744187.5 sps, pre shift, digital AGC; time-sharded")
    return w


def stage_workloads() -> dict:
    """Single-stage workloads for the per-kernel roofline rows of SURVEY.md 8(d) (not BASELINE configs)."""
    w = {}
    # K1 alone: convert + pre-resample LUT-NCO shift, cs16 -> cf32 at the native rate (4 + 8 B per frame, HBM bound)
    w["k1"] = Workload(
        "k1",
        ChainConfig(input_format="cs16", output_format="cf32", input_rate_hz=20.0e6, target_rate_hz=20.0e6,
                    no_resample=True, freq_shift_hz=-100e3),
        tones=[(0.30, 100e3), (0.10, -150e3)], dc=0.0, sigma=0.01,
        parity_samples=1 << 22, throughput_samples=1 << 28,
        description="stage K1 alone: cs16 -> convert + LUT-NCO shift -> cf32 (no resampling)")
    return w
