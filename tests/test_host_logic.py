"""CPU tests of the product's host side (no GPU): the C-ABI library loads and exports every
declared symbol, the host-side design (K0) equals the oracle's liquid objects tap for tap, and
the closed-form chunk bookkeeping equals what the reference's chunk loop produces."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from iq_tool_b200 import gpu
from iq_tool_b200.configs import (AGC_DIGITAL, FILTER_REQ_FFT, FILTER_REQ_FIR, ChainConfig, ChainConfigC, highpass,
                                  lowpass, pass_range, stopband)
from oracle.loader import CpuChain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "iqgpu.h")).read()
    names = sorted(set(re.findall(r"\b(iqgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = C.CDLL(gpu.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.iqgpu_abi_version() == 1


def test_config_struct_layout():
    assert C.sizeof(ChainConfigC) == 160
    assert ChainConfigC.freq_shift_hz.offset == 48 and ChainConfigC.filter_requests.offset == 64
    assert C.sizeof(gpu.ChainInfoC) % 8 == 0


EXTRA = {
    "interp_1p5": ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=1e6, target_rate_hz=1.5e6),
    "interp_6": ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=250e3, target_rate_hz=1.5e6),
    "hp_lp_chain": ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=2e6, target_rate_hz=1e6,
                               filters=[highpass(20e3), lowpass(200e3)]),
    "bpf_fir": ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=2e6, target_rate_hz=1e6,
                           filters=[pass_range(102e3, 215e3)], filter_type_request=FILTER_REQ_FIR),
    "notch_fft": ChainConfig(input_format="cu8", output_format="cu8", input_rate_hz=2.4e6, target_rate_hz=1.2e6,
                             filters=[stopband(-20e3, 20e3)], filter_type_request=FILTER_REQ_FFT),
    "upsample_prefilter": ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=1e6,
                                      target_rate_hz=1.25e6, filters=[lowpass(150e3)], transition_width_hz=20e3),
    "am_nrsc5": ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=2e6, target_rate_hz=46511.71875,
                            agc_enable=True, agc_profile=AGC_DIGITAL),
}


def _all_configs(workloads):
    d = {k: v.config for k, v in workloads.items()}
    d.update(EXTRA)
    return d


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"] + sorted(EXTRA))
def test_host_design_equals_oracle_bit_exact(name, workloads):
    cfg = _all_configs(workloads)[name]
    g = gpu.Chain(cfg, device=-1)
    o = CpuChain(cfg, "oracle")
    gi, oi, mi = g.info(), o.info(), o.msresamp_info()
    assert gi.ratio == oi.ratio
    assert gi.num_halfband == mi.num_halfband and gi.is_interp == mi.is_interp
    assert list(gi.halfband_m[: gi.num_halfband]) == list(mi.m_stage[: mi.num_halfband])
    assert gi.arb_step == mi.step and gi.rate_arbitrary == mi.rate_arbitrary
    assert gi.nco_dtheta == oi.nco_dtheta and gi.nco_is_post == oi.nco_is_post
    assert (gi.filter_impl, gi.filter_num_taps, gi.filter_block_size, gi.filter_post_resample) == \
           (oi.filter_impl, oi.filter_num_taps, oi.filter_block_size, oi.filter_post_resample)
    for i in range(gi.num_halfband):
        assert np.array_equal(g.halfband_taps(i).view(np.uint32), o.halfband_taps(i).view(np.uint32)), i
    assert np.array_equal(g.arb_taps().view(np.uint32), o.arb_taps().view(np.uint32))
    assert np.array_equal(g.filter_taps().view(np.uint32), o.filter_taps().view(np.uint32))


def test_nco_increment_for_a_shift_that_is_not_a_float():
    """AppResources.nco_shift_hz is a double (frequency_shift.c:32-33; input_wav.c:614-628 stores centre - target):
    the 32-bit phase increment must come from the double, not from its float rounding."""
    from oracle.loader import have_ref
    kinds = ["oracle"] + (["ref"] if have_ref() else [])
    rng = np.random.Generator(np.random.PCG64(11))
    differ = 0
    for _ in range(60):
        shift = float(rng.uniform(2.0 ** 24, 4.9e7)) * (1 if rng.random() < 0.5 else -1) + 1.0 / 3.0
        cfg = ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=10e6, target_rate_hz=5e6,
                          freq_shift_hz=shift, freq_shift_is_double=True)
        as_float = ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=10e6, target_rate_hz=5e6,
                               freq_shift_hz=shift)
        g = gpu.Chain(cfg, device=-1).info().nco_dtheta
        differ += g != gpu.Chain(as_float, device=-1).info().nco_dtheta
        for k in kinds:
            assert g == CpuChain(cfg, k).info().nco_dtheta, (k, shift)
    assert differ > 5       # the float round trip would have changed the increment in a good share of these


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5", "interp_1p5", "notch_fft"])
def test_closed_form_output_counts_equal_reference_chunk_loop(name, workloads):
    """iqgpu_chain_predict_output (pure integer arithmetic) == frames the oracle's chunk loop emits."""
    cfg = _all_configs(workloads)[name]
    g = gpu.Chain(cfg, device=-1)
    o = CpuChain(cfg, "oracle")
    n = 9 * 16384 + 4321
    rng = np.random.Generator(np.random.PCG64(3))
    if cfg.input_format == "cf32":
        raw = (rng.standard_normal(2 * n) * 0.1).astype(np.float32)
    elif cfg.input_format == "cu8":
        raw = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    else:
        raw = rng.integers(-2000, 2000, 2 * n, dtype=np.int16)
    out = o.process(raw)
    assert g.predict_output(n) == out.size // 2


def test_invalid_configurations_are_rejected(workloads):
    bad = ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=2e6, target_rate_hz=1e6,
                      filters=[lowpass(800e3)])  # beyond the output Nyquist (filter.c:80-84)
    with pytest.raises(gpu.IqGpuError):
        gpu.Chain(bad, device=-1)
    with pytest.raises(gpu.IqGpuError):
        gpu.Chain(ChainConfig(input_rate_hz=2e6, target_rate_hz=1e6, shift_after_resample=True), device=-1)
    with pytest.raises(gpu.IqGpuError):
        gpu.Chain(ChainConfig(input_rate_hz=2e6, target_rate_hz=100.0), device=-1)  # ratio < 0.001
    with pytest.raises(RuntimeError):
        CpuChain(bad, "oracle")


def test_no_cpu_fallback(workloads):
    """A plan-only chain (or a box without CUDA) must refuse to compute."""
    g = gpu.Chain(workloads["cfg1"].config, device=-1)
    with pytest.raises(gpu.IqGpuError) as e:
        g.process(np.zeros(2 * 1024, dtype=np.int16))
    assert e.value.code == -2
    if gpu.device_count() == 0:
        with pytest.raises(gpu.IqGpuError):
            gpu.Chain(workloads["cfg1"].config, device=0)
        with pytest.raises(gpu.IqGpuError):
            gpu.convert_block_to_cf32(np.zeros(8, dtype=np.int16), 11, 4, 1.0)


def test_fused_front_plans_keep_their_invariants(tmp_path):
    """The compile-time cascade plans of the warp-streaming fused front (S = 0..6): alignment of the per-warp shared-memory
    regions, whole register tiles, history sizes, run periods and producer/consumer ratios, zero padding around the
    polyphase rows, >= 16 warps per CTA.  Host-only program, compiled for sm_100a (no GPU needed)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "plan_check"
    subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "iq_tool_b200", "csrc"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "native", "plan_check.cu")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_closed_form_counts_at_maximum_sizes(name, workloads):
    """The 64-bit closed form ceil(floor(N / 2^S) * 2^24 / step) of the product's host side against Python's big
    integers, from one frame to 2^60 frames (a capture of centuries): no overflow anywhere on the way (the
    intermediate product needs more than 64 bits from N = 2^(40+S) on)."""
    g = gpu.Chain(workloads[name].config, device=-1)
    info = g.info()
    S, step = info.num_halfband, info.arb_step
    assert step > 0
    rng = np.random.Generator(np.random.PCG64(11))
    sizes = [0, 1, 16383, 16384, 10**6, 10**9, 2**36 + 12345, 2**40 + 7, 10**13, 2**44 + 1, 2**52 + 5, 2**60 + 11]
    sizes += [int(rng.integers(1, 2**62)) for _ in range(200)]
    # the hand-over from 64-bit to 128-bit arithmetic: pushes = N >> S around 2^39 and 2^40
    sizes += [((2**k + d) << S) + e for k in (39, 40) for d in (-2, -1, 0, 1) for e in (0, (1 << S) - 1)]
    for n in sizes:
        assert g.resampler_outputs_after(n) == -((-((n >> S) << 24)) // step), n
    # a train the host cannot tabulate (more than 2^30 chunks) is refused, not attempted
    assert g.predict_output(10**9 + 1) > 0
    with pytest.raises(gpu.IqGpuError, match="train too long"):
        g.predict_output(2**44)
    with pytest.raises(gpu.IqGpuError, match="train too long"):
        g.predict_output(2**63)

