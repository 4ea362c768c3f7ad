#!/usr/bin/env python
"""How long does the HOST spend enqueueing one step (reset + process_device) compared with the GPU time of the step?
    python tools/host_overhead.py [workload]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iq_tool_b200 import baseline_workloads, gpu
from iq_tool_b200.synth import synth_torch

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = baseline_workloads()[name]
cfg = wl.config
n = wl.throughput_samples - wl.throughput_samples % 16384
dev = torch.device("cuda", 0)
raw = synth_torch(wl, n, dev)
ch = gpu.Chain(cfg, 0, subtrain_frames=1 << 30)
out = torch.empty(ch.out_capacity_frames(n) * cfg.out_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ch.restart(); ch.process_device(raw.data_ptr(), n, out.data_ptr(), out.numel(), st)
torch.cuda.synchronize()
K = 20
host = []
t0 = time.perf_counter()
for _ in range(K):
    a = time.perf_counter()
    ch.restart()
    b = time.perf_counter()
    ch.process_device(raw.data_ptr(), n, out.data_ptr(), out.numel(), st)
    c = time.perf_counter()
    host.append((b - a, c - b))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{name}: host enqueue per step: reset {1e3*sum(h[0] for h in host)/K:.3f} ms, process_device {1e3*sum(h[1] for h in host)/K:.3f} ms; "
      f"enqueue loop {1e3*(t1-t0)/K:.3f} ms/step, with final sync {1e3*(t2-t0)/K:.3f} ms/step")
