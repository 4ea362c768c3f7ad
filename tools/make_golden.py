"""Generate the golden vectors under tests/golden/ from the REFERENCE'S OWN code
(oracle/_ref/libiqref.so = /root/reference/src/*.c compiled in place on the liquid_compat shim).

Run in the authoring container (needs /root/reference):   python tools/make_golden.py
The fixtures are small (.npz, a few hundred KB) and are committed; the GPU box and fresh clones
check the restated oracle (oracle/iq_oracle.c) and the CUDA chain against them.
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from iq_tool_b200 import baseline_workloads
from iq_tool_b200.configs import BYTES_PER_SAMPLE, FORMAT_CODES, NUMPY_DTYPE
from iq_tool_b200.synth import synth_numpy
from oracle.loader import CpuChain, convert_from_cf32, convert_to_cf32

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GOLDEN_FRAMES = {"cfg1": 3 * 16384 + 1000, "cfg2": 5 * 16384 + 777, "cfg3": 13 * 16384 + 5,
                 "cfg4": 3 * 16384 + 1000, "cfg5": 6 * 16384 + 333}


def conversion_kats():
    """All 65 536 cs16 codes / all 256 8-bit codes x gains, and cf32 -> every integer format on a
    grid that includes clamp edges and .5 ties; recorded as SHA-256 of the reference's output."""
    kats = {}
    gains = [1.0, 0.5, 1.2345]
    codes16 = np.arange(-32768, 32768, dtype=np.int32)
    i16 = np.empty(2 * codes16.size, dtype=np.int16)
    i16[0::2] = codes16.astype(np.int16)
    i16[1::2] = codes16[::-1].astype(np.int16)
    for fmt, raw in (("cs16", i16), ("sc16q11", i16), ("cu16", i16.view(np.uint16)),
                     ("cs8", np.arange(-128, 128, dtype=np.int16).astype(np.int8).repeat(2)),
                     ("cu8", np.arange(0, 256, dtype=np.int16).astype(np.uint8).repeat(2))):
        n = raw.size // 2
        for g in gains:
            y = convert_to_cf32("ref", raw, FORMAT_CODES[fmt], n, g)
            kats[f"to_cf32/{fmt}/gain={g}"] = hashlib.sha256(y.tobytes()).hexdigest()
    rng = np.random.Generator(np.random.PCG64(777))
    grid = np.concatenate([
        np.linspace(-1.25, 1.25, 20001), (np.arange(-300, 301) + 0.5) / 127.0, (np.arange(-300, 301) + 0.5) / 32767.0,
        rng.uniform(-1.1, 1.1, 50000), np.array([0.0, -0.0, 1.0, -1.0, 1e-9, -1e-9])]).astype(np.float32)
    x = (grid + 1j * grid[::-1]).astype(np.complex64)
    for fmt in ("cs8", "cu8", "cs16", "cu16", "sc16q11", "cs24", "cs32", "cu32", "cf32"):
        y = convert_from_cf32("ref", x, FORMAT_CODES[fmt], NUMPY_DTYPE[fmt], BYTES_PER_SAMPLE[fmt])
        kats[f"from_cf32/{fmt}"] = hashlib.sha256(y.tobytes()).hexdigest()
    return kats


def main():
    os.makedirs(OUT, exist_ok=True)
    w = baseline_workloads()
    meta = {}
    for name, wl in w.items():
        n = GOLDEN_FRAMES[name]
        raw = synth_numpy(wl, n)
        ch = CpuChain(wl.config, "ref")
        pre = ch.capture(0, n + 16)
        rs = ch.capture(1, n + 16)
        ch.trace(n // 16384 + 2)
        out = ch.process(raw)
        info = ch.info()
        mi = ch.msresamp_info()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), raw=raw, out=out,
                            pre_head=ch.captured(0)[:8192].copy(), rs=ch.captured(1).copy(), counts=ch.traced().copy(),
                            filter_taps=ch.filter_taps())
        meta[name] = {
            "frames_in": n, "frames_out": int(out.size // 2), "ratio": float(info.ratio),
            "num_halfband": int(mi.num_halfband), "halfband_m": [int(m) for m in mi.m_stage[: mi.num_halfband]],
            "arb_step": int(mi.step), "nco_dtheta": int(info.nco_dtheta),
            "filter_impl": int(info.filter_impl), "filter_num_taps": int(info.filter_num_taps),
            "filter_block_size": int(info.filter_block_size),
            "out_sha256": hashlib.sha256(out.tobytes()).hexdigest(),
        }
        print(name, meta[name])
    meta["conversion_kats"] = conversion_kats()
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
