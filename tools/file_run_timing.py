"""Wall-clock throughput of a whole file run: a synthetic WAV capture in tmpfs ->
iqgpu_wavfile_run (reader thread, chain, writer thread) -> WAV in tmpfs.  Not part of bench.py's contract; run it on the
GPU box:  python tools/file_run_timing.py [cfg2] [frames] [train_chunks]"""
import os, struct, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iq_tool_b200 import baseline_workloads, gpu
from iq_tool_b200.configs import BYTES_PER_SAMPLE
from iq_tool_b200.synth import synth_numpy

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
frames = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 28
train_chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
wl = baseline_workloads()[name]
cfg = wl.config
assert cfg.input_format in ("cs16", "cu8") and cfg.output_format in ("cs16", "cu8"), "WAV holds cs16 / cu8 only"
tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
src, dst = os.path.join(tmp, "iqgpu_timing_in.wav"), os.path.join(tmp, "iqgpu_timing_out.wav")
block = synth_numpy(wl, min(frames, 1 << 24)).tobytes()
payload_bytes = frames * BYTES_PER_SAMPLE[cfg.input_format]
with open(src, "wb") as f:
    bits = 8 * BYTES_PER_SAMPLE[cfg.input_format] // 2
    f.write(gpu.wav_build_header(gpu.CONTAINER_RF64, {"cs16": 11, "cu8": 8}[cfg.input_format], int(cfg.input_rate_hz), payload_bytes))
    left = payload_bytes
    while left:
        n = min(left, len(block))
        f.write(block[:n])
        left -= n
for rep in range(3):
    t0 = time.perf_counter()
    st, info = gpu.wavfile_run(cfg, src, dst, out_container=gpu.CONTAINER_RF64, train_chunks=train_chunks)
    dt = time.perf_counter() - t0
    print(f"{name}: {st.frames_in} frames -> {st.frames_out} frames in {dt * 1e3:.1f} ms = {st.frames_in / dt / 1e6:.0f} Msamples/s "
          f"({st.trains} trains of {train_chunks} chunks, {payload_bytes / dt / 1e9:.2f} GB/s read from {tmp})")
os.remove(src); os.remove(dst)
