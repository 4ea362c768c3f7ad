"""Time-sharding: host logic on CPU (plan-only chains, gloo world_size 2) and
shard-stitch parity on the GPU (`-m gpu`): a capture processed as N shards with halo + closed-form
seek + the digital-AGC peak exchange must equal the single-stream output byte for byte."""
import os
import socket

import numpy as np
import pytest

from iq_tool_b200.configs import AGC_DIGITAL, ChainConfig
from oracle import loader
from oracle.loader import CpuChain

CHUNK = 16384


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _agc_amplitudes():
    """Per-chunk constant amplitudes that walk the digital AGC through every branch at
    rate = 65536 sps (one chunk = 0.25 s): scanning with a rising peak memory, lock after > 2.0 s,
    ratchet down on a strong chunk, 'strong' refresh, > 4 s of weak chunks -> creep up."""
    a = [0.10, 0.20, 0.15, 0.40, 0.30, 0.30, 0.30, 0.30, 0.30, 0.20]       # scan + lock (gain 0.9/0.4)
    a += [0.50, 0.45]                                                          # 0.5*2.25 > 1 -> ratchet to 0.99/0.5
    a += [0.36, 0.05] + [0.04] * 20                                            # strong refresh, then weak > 4 s -> creep
    a += [0.30, 0.02, 0.02]
    return np.array(a, dtype=np.float32)


def test_agc_advance_equals_the_oracle_state_machine():
    """iqgpu_agc_digital_advance (host, no device) against the oracle's agc_apply
    (reference src/agc.c:105-222) driven on the sample clock."""
    from iq_tool_b200 import gpu
    rate = 65536.0
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=rate, target_rate_hz=rate,
                      no_resample=True, agc_enable=True, agc_profile=AGC_DIGITAL)
    lib = loader.get_lib("oracle")[0]
    ch = CpuChain(cfg, "oracle")
    amps = _agc_amplitudes()
    ref_gain = np.zeros(amps.size, dtype=np.float32)
    try:
        for c, a in enumerate(amps):
            lib.iqo_set_fake_clock(1, c * CHUNK / rate)
            x = np.full(CHUNK, a, dtype=np.complex64)
            y = ch.process(x.view(np.float32)).view(np.complex64)
            ref_gain[c] = np.float32(y[0].real) / a
    finally:
        lib.iqo_set_fake_clock(0, 0.0)
    st = gpu.agc_initial_state()
    gains = gpu.agc_digital_advance(st, 0.9, rate, amps, np.full(amps.size, CHUNK, np.uint32), want_gains=True)
    assert np.allclose(gains, ref_gain, rtol=2e-7, atol=0)
    info = ch.info()
    assert (st.locked, st.samples_seen) == (info.agc_locked, info.agc_samples_seen)
    assert st.gain == pytest.approx(info.agc_gain, rel=1e-7) and st.peak_memory == info.agc_peak_memory
    assert gains[11] < gains[9] and gains[-1] > gains[14]           # ratchet happened, creep happened
    # advancing in two pieces == advancing at once; empty chunks are ignored (agc.c:89)
    a2 = gpu.agc_initial_state()
    k = 13
    cnt = np.full(amps.size, CHUNK, np.uint32)
    gpu.agc_digital_advance(a2, 0.9, rate, amps[:k], cnt[:k])
    gpu.agc_digital_advance(a2, 0.9, rate, np.array([9.0], np.float32), np.zeros(1, np.uint32))
    gpu.agc_digital_advance(a2, 0.9, rate, amps[k:], cnt[k:])
    assert a2.as_tuple() == st.as_tuple()


@pytest.mark.parametrize("name,total", [("cfg1", 40 * CHUNK + 777), ("cfg2", 1000 * CHUNK), ("cfg5", 257 * CHUNK + 1)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_plan_tiles_the_capture(name, total, world, workloads):
    from iq_tool_b200 import gpu
    from iq_tool_b200.shard import plan_shards
    probe = gpu.Chain(workloads[name].config, -1)
    shards = plan_shards(probe, total, world)
    assert [s.rank for s in shards] == list(range(world))
    assert shards[0].start == 0 and shards[0].lead == 0 and shards[0].drop == 0
    assert sum(s.frames for s in shards) == total
    halo = probe.halo_frames()
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.start + a.frames == b.start and a.out_end == b.out_start
        assert b.start % CHUNK == 0 and b.lead % CHUNK == 0
        assert b.start - b.lead >= min(halo, b.start) and b.skip_chunks * CHUNK == b.start - b.lead
        assert b.out_lead <= b.out_start
    assert shards[-1].out_end == probe.predict_output(total)


def _gloo_worker(rank, world, port, total_chunks):
    import torch.distributed as dist
    from iq_tool_b200 import baseline_workloads, gpu
    from iq_tool_b200.shard import _all_gather_var, exchange_agc_state, plan_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = baseline_workloads()["cfg1"].config
        probe = gpu.Chain(cfg, -1)
        total = total_chunks * CHUNK - 1234                       # ragged last chunk
        shard = plan_shards(probe, total, world)[rank]
        # the whole capture's per-chunk peaks and frame counts (every rank can regenerate them)
        rng = np.random.default_rng(7)
        peaks_all = (0.02 + 0.5 * rng.random(total_chunks) ** 4).astype(np.float32)
        peaks_all[total_chunks // 3] = 3.0                       # forces a ratchet once locked
        edges = [probe.resampler_outputs_after(min(c * CHUNK, total)) for c in range(total_chunks + 1)]
        counts_all = np.diff(np.array(edges, dtype=np.int64)).astype(np.uint32)
        c0, c1 = shard.start // CHUNK, (shard.start + shard.frames + CHUNK - 1) // CHUNK
        rate = float(np.float32(cfg.target_rate_hz))
        state = exchange_agc_state(peaks_all[c0:c1], counts_all[c0:c1], 0.9, rate, None, "cpu")
        gains = gpu.agc_digital_advance(state, 0.9, rate, peaks_all[c0:c1], counts_all[c0:c1], want_gains=True)
        stitched = np.concatenate(_all_gather_var(gains, None, "cpu"))
        single = gpu.agc_initial_state()
        ref = gpu.agc_digital_advance(single, 0.9, rate, peaks_all, counts_all, want_gains=True)
        assert stitched.size == total_chunks and np.array_equal(stitched, ref)
        if rank == world - 1:
            assert state.as_tuple() == single.as_tuple()
        spans = _all_gather_var(np.array([shard.out_start, shard.out_end], dtype=np.uint32), None, "cpu")
        assert all(int(a[1]) == int(b[0]) for a, b in zip(spans[:-1], spans[1:]))
        assert int(spans[-1][1]) == probe.predict_output(total)
    finally:
        dist.destroy_process_group()


def test_agc_exchange_over_gloo_world2_equals_single_stream():
    """The N>1 host path on CPU: 2 gloo ranks plan their shards, exchange per-chunk peaks and end
    up with exactly the per-chunk gains of the single-stream state machine."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2, _free_port(), 64 * 30), nprocs=2, join=True)


# ------------------------------------------------------------------------------------------------
# GPU: shard-stitch parity
# ------------------------------------------------------------------------------------------------
def _run_sharded(gpu, wl, raw, world, device_exchange=False):
    """Process `raw` as `world` shards one after the other on cuda:0 (the data path has no
    collective, so ranks need not run concurrently; the AGC exchange is replayed in rank order).
    device_exchange: the peaks stay on the GPU (the all-gather is a set of device buffers here) and the lower
    shards' chunks are replayed by the scan kernel (iqgpu_chain_agc_advance_device) instead of on the host."""
    import torch
    from iq_tool_b200.shard import ShardedChain
    cfg = wl.config
    total = raw.size // 2
    dev = torch.device("cuda", 0)
    raw_d = torch.from_numpy(raw).to(dev)
    sc = ShardedChain(cfg, 0, shard_frames_hint=total)
    shards = sc.plan(total, world)
    parts, live_peaks, live_counts = [], [], []
    esz = raw.dtype.itemsize * 2
    for sh in shards:
        ch = sc.chain
        ch.seek(sh.lead)
        out = torch.zeros(ch.out_capacity_frames(sh.read_frames) * cfg.out_bytes, dtype=torch.uint8, device=dev)
        ptr = raw_d.data_ptr() + sh.lead * esz
        if sc.digital_agc and device_exchange:
            ch.process_device_begin(ptr, sh.read_frames)
            nlive = (sh.frames + CHUNK - 1) // CHUNK
            mine = torch.zeros(max(nlive, 1), dtype=torch.float32, device=dev)
            assert ch.pending_chunk_peaks_device(sh.skip_chunks, mine.data_ptr(), mine.numel()) == nlive
            for q, buf in zip(shards, live_peaks):
                if q.frames:
                    ch.agc_advance_device(buf.data_ptr(), q.start, q.frames)
            live_peaks.append(mine)
            produced = ch.process_device_finish(sh.skip_chunks, out.data_ptr(), out.numel())
        elif sc.digital_agc:
            ch.process_device_begin(ptr, sh.read_frames)
            pk, ct = ch.pending_chunk_peaks()
            st = gpu.agc_initial_state()
            for p_, c_ in zip(live_peaks, live_counts):
                gpu.agc_digital_advance(st, sc.agc_target, sc.target_rate, p_, c_)
            ch.set_agc_state(st)
            live_peaks.append(pk[sh.skip_chunks:].copy()); live_counts.append(ct[sh.skip_chunks:].copy())
            produced = ch.process_device_finish(sh.skip_chunks, out.data_ptr(), out.numel())
        else:
            produced = ch.process_device(ptr, sh.read_frames, out.data_ptr(), out.numel())
        torch.cuda.synchronize()
        assert produced - sh.drop == sh.out_frames
        parts.append(out[sh.drop * cfg.out_bytes: produced * cfg.out_bytes].cpu().numpy())
    from iq_tool_b200.configs import NUMPY_DTYPE
    return np.concatenate(parts).view(NUMPY_DTYPE[cfg.output_format])


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunks,world", [("cfg1", 96, 2), ("cfg1", 97, 4), ("cfg5", 256, 2), ("cfg5", 250, 8)])
def test_sharded_output_equals_single_stream_bit_for_bit(name, chunks, world, gpu, workloads):
    """Finite-memory chains (FIR resampler + LUT NCO + digital AGC): sharded == single, exactly."""
    from iq_tool_b200.synth import synth_numpy
    wl = workloads[name]
    raw = synth_numpy(wl, chunks * CHUNK - 321)
    # make the AGC do something: a loud burst in the second half
    h = raw.size // 2
    raw[h: h + 4 * CHUNK] = (raw[h: h + 4 * CHUNK].astype(np.int32) * 2).clip(-32768, 32767).astype(raw.dtype)
    single = gpu.Chain(wl.config, 0).process(raw)
    sharded = _run_sharded(gpu, wl, raw, world)
    assert sharded.size == single.size
    assert np.array_equal(sharded, single)
    on_device = _run_sharded(gpu, wl, raw, world, device_exchange=True)
    assert np.array_equal(on_device, single)


@pytest.mark.gpu
def test_sharded_cfg2_with_dc_block_halo_meets_the_bar(gpu, workloads):
    """cfg2 has the infinite-memory DC blocker: the shard re-computes a 16-time-constant halo, so
    the stitched output equals the single stream within +-1 LSB (cs16)."""
    from iq_tool_b200.synth import synth_numpy
    wl = workloads["cfg2"]
    probe = gpu.Chain(wl.config, -1)
    halo = probe.halo_frames()
    total = 2 * (halo + 40 * CHUNK)
    raw = synth_numpy(wl, total)
    single = gpu.Chain(wl.config, 0).process(raw)
    sharded = _run_sharded(gpu, wl, raw, 2)
    assert sharded.size == single.size
    d = np.abs(sharded.astype(np.int32) - single.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 5])
def test_sharded_interpolating_chain_equals_single_stream(world, gpu):
    """r = 6 (arbitrary stage at 1.5, two halfband interpolators) with a pre-resample shift: the halo covers the
    arbitrary stage's 14-frame window plus the interpolators' histories mapped back to input frames (it used
    to be 0 for interpolating chains), so the stitched shards equal the single stream bit for bit."""
    import types
    cfg = ChainConfig(input_format="cs16", output_format="cs16", input_rate_hz=250e3, target_rate_hz=1.5e6,
                      freq_shift_hz=-20e3)
    probe = gpu.Chain(cfg, -1)
    assert probe.info().is_interp == 1 and probe.halo_frames() >= 14
    rng = np.random.default_rng(5)
    n = 23 * CHUNK + 4321
    t = np.arange(n)
    x = 0.4 * np.exp(2j * np.pi * 0.05 * t) + 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    raw = np.empty(2 * n, dtype=np.int16)
    raw[0::2] = np.clip(np.rint(x.real * 32767), -32768, 32767)
    raw[1::2] = np.clip(np.rint(x.imag * 32767), -32768, 32767)
    wl = types.SimpleNamespace(config=cfg)
    single = gpu.Chain(cfg, 0).process(raw)
    sharded = _run_sharded(gpu, wl, raw, world)
    assert sharded.size == single.size and np.array_equal(sharded, single)


@pytest.mark.gpu
def test_shard_stitch_at_2_pow_28_frames(gpu, workloads):
    """shard-stitch parity of cfg5 on 2^28 input frames (1 GiB of cs16, generated in HBM): eight time shards
    with halo, closed-form seek and the device-side AGC peak exchange == the single stream, byte for byte."""
    import torch
    from iq_tool_b200.shard import ShardedChain
    from iq_tool_b200.synth import synth_torch
    wl = workloads["cfg5"]
    cfg = wl.config
    total = 1 << 28
    dev = torch.device("cuda", 0)
    raw = synth_torch(wl, total, dev)
    h = (total // 8) * 6 + 100 * CHUNK                # a loud stretch: the AGC ratchets inside shard 6
    raw[2 * h: 2 * (h + 8 * CHUNK)] = torch.clamp(raw[2 * h: 2 * (h + 8 * CHUNK)].to(torch.int32) * 3, -32768, 32767).to(torch.int16)
    single_chain = gpu.Chain(cfg, 0, subtrain_frames=1 << 30)
    out1 = torch.zeros(single_chain.out_capacity_frames(total) * cfg.out_bytes, dtype=torch.uint8, device=dev)
    n1 = single_chain.process_device(raw.data_ptr(), total, out1.data_ptr(), out1.numel())
    torch.cuda.synchronize()
    from iq_tool_b200.shard import lower_shard_pieces
    sc = ShardedChain(cfg, 0, shard_frames_hint=total // 8 + (1 << 20))
    shards = sc.plan(total, 8)
    m = max((sh.frames + CHUNK - 1) // CHUNK for sh in shards)
    gathered = torch.zeros(8 * m, dtype=torch.float32, device=dev)        # the all-gather's receive buffer: rank r at r * m
    stitched = torch.zeros_like(out1)
    pos, calls = 0, []
    assert shards[3].start < sc.lock_frames < shards[3].start + shards[3].frames     # the lock falls inside shard 3
    for sh in shards:
        ch = sc.chain
        ch.seek(sh.lead)
        out = torch.zeros(ch.out_capacity_frames(sh.read_frames) * cfg.out_bytes, dtype=torch.uint8, device=dev)
        ch.process_device_begin(raw.data_ptr() + sh.lead * 4, sh.read_frames)
        nlive = (sh.frames + CHUNK - 1) // CHUNK
        mine = gathered[sh.rank * m: sh.rank * m + m]
        assert ch.pending_chunk_peaks_device(sh.skip_chunks, mine.data_ptr(), m) == nlive
        # the pieces the NCCL path advances over (ShardedChain._finish_device_exchange): the head of the capture up to the
        # lock on its own, contiguous locked stretches merged into one call (settled by the grid-wide quiet test when
        # nothing happens in them: ranks 5 and 6; rank 7's stretch holds the ratchet and takes the tile walk)
        pieces = lower_shard_pieces(shards, sh.rank, m, gathered.data_ptr(), sc.lock_frames)
        calls.append(len(pieces))
        for ptr, first, frames in pieces:
            ch.agc_advance_device(ptr, first, frames)
        produced = ch.process_device_finish(sh.skip_chunks, out.data_ptr(), out.numel())
        torch.cuda.synchronize()
        assert produced - sh.drop == sh.out_frames
        k = sh.out_frames * cfg.out_bytes
        stitched[pos: pos + k] = out[sh.drop * cfg.out_bytes: produced * cfg.out_bytes]
        pos += k
    assert calls == [0, 1, 2, 3, 5, 5, 5, 5]             # three pre-lock shards, the split one, then ONE merged call
    assert pos == n1 * cfg.out_bytes
    assert torch.equal(stitched[:pos], out1[:pos])
    assert single_chain.info().agc_locked == 1


@pytest.mark.gpu
def test_state_only_advance_quiet_test_equals_the_state_machine(gpu):
    """iqgpu_chain_agc_advance_device over long chunk tables: the grid-wide quiet test (locked AGC, nothing happens -> the
    state moves on in one step) and its fall-back (any ratchet / creep / not yet locked -> the sequential scan) must both
    leave exactly the state of the host state machine (src/agc.c:105-222 on the sample clock).  Tables: quiet; one ratchet;
    a weak stretch longer than 4 s (creep) that straddles several 4096-chunk parts; a weak stretch just short of 4 s;
    weak chunks in front of the first strong one; an unlocked start."""
    import torch
    rate, chunk = 65536.0, 64                             # 4 s = 4096 chunks
    cfg = ChainConfig(input_format="cf32", output_format="cf32", input_rate_hz=rate, target_rate_hz=rate,
                      no_resample=True, agc_enable=True, agc_profile=AGC_DIGITAL)
    nch = 20000
    rng = np.random.default_rng(9)
    base = (0.40 + 0.04 * rng.random(nch)).astype(np.float32)          # strong at gain 2.0 (0.8 .. 0.88 > 0.675), no ratchet
    tables = {"quiet": base.copy()}
    t = base.copy(); t[12345] = 0.51; tables["ratchet"] = t             # 0.51 * 2.0 > 1
    t = base.copy(); t[3000:3000 + 4200] = 0.05; tables["creep"] = t    # weak for > 4 s: creeps
    t = base.copy(); t[8000:8000 + 4090] = 0.05; tables["almost"] = t   # weak for < 4 s: nothing happens
    t = base.copy(); t[:5000] = 0.05; tables["weak_first"] = t          # weak from the start, last strong 0.5 s before the table
    start = gpu.agc_initial_state()
    start.locked, start.gain, start.peak_memory, start.samples_seen, start.last_strong_s = 1, 2.0, 0.45, 10 * 65536, 9.5
    cases = [(k, v, start) for k, v in tables.items()] + [("unlocked", base.copy(), gpu.agc_initial_state())]
    counts = np.full(nch, chunk, dtype=np.uint32)
    for name, peaks, st0 in cases:
        ch = gpu.Chain(cfg, 0, chunk_frames=chunk)
        ch.seek(0)
        ch.set_agc_state(st0)
        d = torch.from_numpy(peaks).cuda()
        ch.agc_advance_device(d.data_ptr(), 0, nch * chunk)
        got = ch.get_agc_state().as_tuple()
        ref = gpu.agc_initial_state()
        ref.locked, ref.gain, ref.peak_memory, ref.samples_seen, ref.last_strong_s = st0.locked, st0.gain, st0.peak_memory, st0.samples_seen, st0.last_strong_s
        gpu.agc_digital_advance(ref, 0.9, rate, peaks, counts)
        want = ref.as_tuple()
        assert got[0] == want[0] and got[3] == want[3], (name, got, want)
        assert got[1] == pytest.approx(want[1], rel=1e-6) and got[2] == pytest.approx(want[2], rel=1e-6), (name, got, want)
        assert got[4] == pytest.approx(want[4], abs=1e-9), (name, got, want)
    # the quiet and the almost tables left the gain alone; ratchet lowered it, creep raised it
