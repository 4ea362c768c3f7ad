"""CPU tests of the oracle itself (no GPU): the restated oracle (oracle/iq_oracle.c) against the
golden vectors produced by the reference's own code, and sanity checks of the restated liquid
layer against independent float64 math (catches a self-consistent but wrong recollection)."""
import hashlib
import os

import numpy as np
import pytest

from iq_tool_b200.configs import (AGC_DIGITAL, BYTES_PER_SAMPLE, FORMAT_CODES, NUMPY_DTYPE, ChainConfig,
                                  lowpass)
from oracle import loader
from oracle.loader import CpuChain, convert_from_cf32, convert_to_cf32

CFGS = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
KINDS = ["oracle"] + (["ref"] if os.path.isdir(loader.REF_ROOT) or loader.have_ref() else [])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", CFGS)
def test_chain_matches_golden_bit_exact(name, kind, workloads, golden):
    meta, data = golden
    g = data[name]
    wl = workloads[name]
    ch = CpuChain(wl.config, kind)
    n = meta[name]["frames_in"]
    ch.capture(0, n + 16)
    ch.capture(1, n + 16)
    ch.trace(n // 16384 + 2)
    out = ch.process(g["raw"])
    assert out.size // 2 == meta[name]["frames_out"]
    assert hashlib.sha256(out.tobytes()).hexdigest() == meta[name]["out_sha256"]
    assert np.array_equal(out.view(np.uint8), g["out"].view(np.uint8))
    assert np.array_equal(ch.captured(0)[:8192].view(np.uint32), g["pre_head"].view(np.uint32))
    assert np.array_equal(ch.captured(1).view(np.uint32), g["rs"].view(np.uint32))
    assert np.array_equal(ch.traced(), g["counts"])
    assert np.array_equal(ch.filter_taps().view(np.uint32), g["filter_taps"].view(np.uint32))
    info, mi = ch.info(), ch.msresamp_info()
    assert info.nco_dtheta == meta[name]["nco_dtheta"]
    assert mi.step == meta[name]["arb_step"]
    assert list(mi.m_stage[: mi.num_halfband]) == meta[name]["halfband_m"]
    assert info.filter_impl == meta[name]["filter_impl"]
    assert info.filter_block_size == meta[name]["filter_block_size"]


def _kat_inputs():
    codes16 = np.arange(-32768, 32768, dtype=np.int32)
    i16 = np.empty(2 * codes16.size, dtype=np.int16)
    i16[0::2] = codes16.astype(np.int16)
    i16[1::2] = codes16[::-1].astype(np.int16)
    ins = {"cs16": i16, "sc16q11": i16, "cu16": i16.view(np.uint16),
           "cs8": np.arange(-128, 128, dtype=np.int16).astype(np.int8).repeat(2),
           "cu8": np.arange(0, 256, dtype=np.int16).astype(np.uint8).repeat(2)}
    rng = np.random.Generator(np.random.PCG64(777))
    grid = np.concatenate([
        np.linspace(-1.25, 1.25, 20001), (np.arange(-300, 301) + 0.5) / 127.0, (np.arange(-300, 301) + 0.5) / 32767.0,
        rng.uniform(-1.1, 1.1, 50000), np.array([0.0, -0.0, 1.0, -1.0, 1e-9, -1e-9])]).astype(np.float32)
    x = (grid + 1j * grid[::-1]).astype(np.complex64)
    return ins, x


def test_conversion_kats_match_reference_hashes(golden):
    """Every cs16 / 8-bit code x gains and cf32 -> all integer formats incl. clamp edges and ties:
    the restated converters reproduce the reference's sample_convert.c bit for bit."""
    meta, _ = golden
    kats = meta["conversion_kats"]
    ins, x = _kat_inputs()
    for fmt, raw in ins.items():
        for g in (1.0, 0.5, 1.2345):
            y = convert_to_cf32("oracle", raw, FORMAT_CODES[fmt], raw.size // 2, g)
            assert hashlib.sha256(y.tobytes()).hexdigest() == kats[f"to_cf32/{fmt}/gain={g}"], (fmt, g)
    for fmt in ("cs8", "cu8", "cs16", "cu16", "sc16q11", "cs24", "cs32", "cu32", "cf32"):
        y = convert_from_cf32("oracle", x, FORMAT_CODES[fmt], NUMPY_DTYPE[fmt], BYTES_PER_SAMPLE[fmt])
        assert hashlib.sha256(y.tobytes()).hexdigest() == kats[f"from_cf32/{fmt}"], fmt


def test_conversion_known_values():
    """Hand-checkable values from sample_convert.c: cu8 of 0.0 is 128, clamps, sign handling."""
    x = np.array([0.0 + 0.0j, 1.0 - 1.0j, 2.0 - 2.0j, 0.5 / 127 + 0j], dtype=np.complex64)
    cu8 = convert_from_cf32("oracle", x, FORMAT_CODES["cu8"], np.uint8, 2)
    assert list(cu8[:6]) == [128, 128, 255, 1, 255, 0]
    cs16 = convert_from_cf32("oracle", x, FORMAT_CODES["cs16"], np.int16, 4)
    assert list(cs16[:6]) == [0, 0, 32767, -32767, 32767, -32768]
    raw = np.array([-32768, 32767, 0, 1], dtype=np.int16)
    y = convert_to_cf32("oracle", raw, FORMAT_CODES["cs16"], 2, 1.0)
    assert y[0] == np.complex64(-1.0 + (32767 / 32768) * 1j) and y[1] == np.complex64(0 + (1 / 32768) * 1j)
    assert loader.get_lib("oracle")[0].iqo_get_bytes_per_sample(FORMAT_CODES["cs24"]) == 6


def test_nco_constrain_and_phase_closed_form():
    lib, _ = loader.get_lib("oracle")
    assert lib.liquid_compat_nco_constrain(np.float32(np.pi / 2)) in (1 << 30, (1 << 30) - 256, (1 << 30) + 256)
    assert lib.liquid_compat_nco_constrain(np.float32(0.0)) == 0
    # negative angles wrap into [0, 2^32)
    assert abs(int(lib.liquid_compat_nco_constrain(np.float32(-np.pi / 2))) - 3 * (1 << 30)) <= 512
    # shift -100 kHz @ 2 Msps: d_theta = 0.05 * 2^32 (float-precision)
    w = np.float32(2.0 * np.pi * 100e3 / 2e6)
    assert abs(int(lib.liquid_compat_nco_constrain(w)) - int(0.05 * 2 ** 32)) <= 256


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg5"])
def test_resampler_output_count_closed_form(name, workloads):
    """Output count after N inputs is ceil(floor(N/2^S) * 2^24 / step) for any chunking."""
    wl = workloads[name]
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=wl.config.input_rate_hz,
                      target_rate_hz=wl.config.target_rate_hz)
    ch = CpuChain(cfg, "oracle")
    mi = ch.msresamp_info()
    S, step = mi.num_halfband, mi.step
    total_in = total_out = 0
    rng = np.random.Generator(np.random.PCG64(5))
    for n in (1, 16383, 16384, 100000, 7, 3 * 16384 + 1):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * 0.1
        y = ch.process(x.view(np.float32))
        total_in += n
        total_out += y.size // 2
        expect = -((-((total_in >> S) << 24)) // step)
        assert total_out == expect, (n, total_out, expect)


def test_halfband_and_polyphase_design_properties(workloads):
    ch = CpuChain(workloads["cfg5"].config, "oracle")
    mi = ch.msresamp_info()
    assert mi.num_halfband == 6 and mi.npfb == 256 and mi.arb_m == 7
    for i in range(mi.num_halfband):
        h = ch.halfband_taps(i).astype(np.float64)
        m = mi.m_stage[i]
        assert h.size == 4 * m + 1
        assert abs(h[2 * m] - 1.0) < 1e-6                      # centre tap
        assert np.allclose(h, h[::-1], atol=1e-6)              # linear phase
        even = np.delete(h[0::2], m)                           # even taps other than the centre vanish
        assert np.abs(even).max() < 1e-6
        assert abs(h[1::2].sum() - 1.0) < 5e-3                 # odd branch DC gain ~ 1 (stage gain 2)
        # stop band of the halfband prototype is at least ~60 dB down
        H = np.abs(np.fft.fft(h / 2.0, 8192))
        assert H[4096] < 1e-3
    taps = ch.arb_taps().astype(np.float64)
    assert taps.size == 2 * 7 * 256
    bank_gain = taps.reshape(14, 256).sum(axis=0)              # each polyphase arm has unit DC gain
    assert np.abs(bank_gain - 1.0).max() < 2e-3


def test_kaiser_lowpass_against_scipy(workloads):
    """liquid_firdes_kaiser as restated == windowed sinc with scipy's Kaiser window (float64)."""
    from scipy.signal.windows import kaiser
    ch = CpuChain(workloads["cfg2"].config, "oracle")
    h = ch.filter_taps().real.astype(np.float64)
    n = 255
    fc = np.float32(100e3) / np.float32(744187.5)
    beta = 0.1102 * (60.0 - 8.7)
    t = np.arange(n) - (n - 1) / 2
    ref = np.sinc(2 * fc * t) * kaiser(n, beta)
    ref /= ref.sum()
    assert np.abs(h - ref).max() < 2e-6
    assert abs(h.sum() - 1.0) < 1e-5


def test_fir_matches_float64_convolution(workloads):
    wl = workloads["cfg2"]
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=744187.0, target_rate_hz=744187.0,
                      no_resample=True, filters=[lowpass(100e3)], filter_taps=255, filter_type_request=1)
    ch = CpuChain(cfg, "oracle")
    rng = np.random.Generator(np.random.PCG64(11))
    x = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)).astype(np.complex64) * 0.2
    y = ch.process(x.view(np.float32)).view(np.complex64)
    h = ch.filter_taps().astype(np.complex128)
    ref = np.convolve(x.astype(np.complex128), h)[: x.size]
    assert np.sqrt(np.mean(np.abs(y - ref) ** 2)) < 1e-6


def test_fft_filter_equals_fir_and_quantises_output(workloads):
    """fftfilt (overlap-add, block n) == firfilt on the same taps; output length is a multiple of
    the block and the tail is withheld."""
    base = dict(input_format="cf32", output_format="cf32", input_rate_hz=1e6, target_rate_hz=1e6,
                no_resample=True, filters=[lowpass(100e3)], filter_taps=255)
    fir = CpuChain(ChainConfig(filter_type_request=1, **base), "oracle")
    fft = CpuChain(ChainConfig(filter_type_request=2, **base), "oracle")
    block = fft.info().filter_block_size
    assert block == 512
    rng = np.random.Generator(np.random.PCG64(12))
    n = 5 * 16384 + 100
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * 0.2
    a = fir.process(x.view(np.float32)).view(np.complex64)
    b = fft.process(x.view(np.float32)).view(np.complex64)
    assert b.size == (n // block) * block and a.size == n
    assert np.sqrt(np.mean(np.abs(a[: b.size] - b) ** 2)) < 2e-6


def test_dc_blocker_against_float64_recurrence():
    """The restated DC blocker follows v[n]=x[n]+c v[n-1], y=v[n]-v[n-1]; its deviation from the
    float64 recurrence is bounded by the float rounding of the integrator state (ulp(|v|))."""
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=2.0e6, target_rate_hz=2.0e6,
                      no_resample=True, dc_block=True)
    ch = CpuChain(cfg, "oracle")
    n = 200000
    rng = np.random.Generator(np.random.PCG64(13))
    x = (0.02 + 0.1 * rng.standard_normal(n) + 1j * (0.01 + 0.1 * rng.standard_normal(n))).astype(np.complex64)
    y = ch.process(x.view(np.float32)).view(np.complex64)
    alpha = np.float32(2.0 * np.pi * 10.0 / 2000000)
    c = np.float64(-(np.float32(-1.0) + alpha))
    from scipy.signal import lfilter
    ref = lfilter([1.0, -1.0], [1.0, -c], x.astype(np.complex128))
    v_max = np.abs(lfilter([1.0], [1.0, -c], x.astype(np.complex128))).max()
    assert np.abs(y - ref).max() < 4 * np.spacing(np.float32(v_max))


def test_digital_agc_state_machine(workloads):
    """Peak-memory gain while scanning, lock once MORE than 2.0 s of output samples were seen
    before the current chunk (agc.c:145-149,220)."""
    rate = 65536.0
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=rate, target_rate_hz=rate,
                      no_resample=True, agc_enable=True, agc_profile=AGC_DIGITAL)
    ch = CpuChain(cfg, "oracle")
    n_chunks = 12
    x = np.zeros(n_chunks * 16384, dtype=np.complex64)
    amps = [0.1, 0.2, 0.15, 0.4, 0.3, 0.3, 0.3, 0.3, 0.3, 0.2, 0.2, 0.2]
    for c, a in enumerate(amps):
        x[c * 16384:(c + 1) * 16384] = a
    loader.get_lib("oracle")[0].iqo_set_fake_clock(1, 0.0)
    y = ch.process(x.view(np.float32)).view(np.complex64)
    loader.get_lib("oracle")[0].iqo_set_fake_clock(0, 0.0)
    peak = 0.05
    for c, a in enumerate(amps[:9]):
        peak = max(peak, np.float32(a))
        g = np.float32(0.9) / np.float32(peak)
        assert np.allclose(y[c * 16384].real, np.float32(a) * g, rtol=1e-6), c
    info = ch.info()
    # 16384 samples per chunk at 65536 sps: elapsed > 2.0 s first holds when 9 chunks were seen (2.25 s)
    assert info.agc_locked == 1 and info.agc_samples_seen == n_chunks * 16384
    assert abs(info.agc_gain - 0.9 / 0.4) < 1e-6


def test_chunk_invariance_of_cf32_streams(workloads):
    """Stream stages (everything except the per-chunk digital AGC) do not depend on chunk cuts."""
    wl = workloads["cfg2"]
    from iq_tool_b200.synth import synth_numpy
    raw = synth_numpy(wl, 5 * 16384)
    a = CpuChain(wl.config, "oracle")
    b = CpuChain(wl.config, "oracle")
    ya = a.process(raw)
    parts = []
    cuts = [0, 100, 16384, 16385, 40000, 5 * 16384]
    for s, e in zip(cuts[:-1], cuts[1:]):
        parts.append(b.process(raw[2 * s:2 * e]))
    yb = np.concatenate(parts)
    assert np.array_equal(ya, yb)
