"""Developer diagnostic: run each BASELINE config through the GPU chain and the CPU oracle and
print agreement metrics per tap.  (The pytest suite is the gate; this prints numbers.)"""
import sys, time, traceback
sys.path.insert(0, '.')
import numpy as np
from iq_tool_b200 import baseline_workloads
from iq_tool_b200.synth import synth_numpy
from iq_tool_b200.gpu import Chain, device_count
from oracle.loader import CpuChain, have_ref

def metrics(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    if a.size == 0:
        return "empty"
    if np.iscomplexobj(a) or a.dtype.kind == 'f':
        d = (a.astype(np.complex128) - b.astype(np.complex128)) if np.iscomplexobj(a) else (a.astype(np.float64) - b.astype(np.float64))
        rms = np.sqrt(np.mean(np.abs(d) ** 2)); ref = np.sqrt(np.mean(np.abs(b.astype(np.complex128) if np.iscomplexobj(b) else b.astype(np.float64)) ** 2))
        snr = 20 * np.log10(ref / rms) if rms > 0 else np.inf
        return f"rms_err={rms:.3e} (fullscale-rel) max={np.abs(d).max():.3e} ref_rms={ref:.3e} snr={snr:.1f} dB bitexact={np.array_equal(a, b)}"
    d = a.astype(np.int64) - b.astype(np.int64)
    return f"max_lsb={np.abs(d).max()} n_diff={np.count_nonzero(d)}/{d.size} bitexact={np.array_equal(a, b)}"

def main():
    print("devices", device_count(), "have_ref", have_ref())
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20) + 12345
    names = sys.argv[2].split(',') if len(sys.argv) > 2 else ['cfg1', 'cfg2', 'cfg4', 'cfg5', 'cfg3']
    kind = 'ref' if have_ref() else 'oracle'
    for name in names:
        wl = baseline_workloads()[name]
        try:
            raw = synth_numpy(wl, n)
            o = CpuChain(wl.config, kind)
            c0 = o.capture(0, n + 16); c1 = o.capture(1, n + 16); c2 = o.capture(2, n + 16) if kind == 'oracle' else None
            o.trace(n // 16384 + 2)
            t = time.time(); ref_out = o.process(raw); t_cpu = time.time() - t
            g = Chain(wl.config, 0, record_taps=1)
            t = time.time(); out, counts = g.process(raw, return_chunk_counts=True); t_gpu = time.time() - t
            print(f"== {name}: n={n} cpu {n/t_cpu/1e6:.1f} Msps, gpu(host path, cold) {n/t_gpu/1e6:.1f} Msps, launches {g.info().kernel_launches}")
            print("   chunk counts equal:", np.array_equal(counts, o.traced()), "frames", out.size // 2, ref_out.size // 2)
            print("   tap0 pre :", metrics(g.read_tap(0), o.captured(0)))
            print("   tap1 rs  :", metrics(g.read_tap(1), o.captured(1)))
            if c2 is not None:
                print("   tap2 post:", metrics(g.read_tap(2), o.captured(2)))
            if wl.config.output_format == 'cf32':
                print("   out      :", metrics(out.view(np.complex64), ref_out.view(np.complex64)))
            else:
                print("   out      :", metrics(out, ref_out))
            t = time.time(); out2 = g.process(raw); t_gpu = time.time() - t
            print(f"   second call (state carried): {n/t_gpu/1e6:.1f} Msps host path")
        except Exception as e:
            print(f"== {name}: FAILED {e}")
            traceback.print_exc()

main()
