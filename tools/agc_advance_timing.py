import sys, time; sys.path.insert(0, '/root/repo')
import torch, numpy as np
from iq_tool_b200 import baseline_workloads, gpu
wl = baseline_workloads()["cfg5"]
ch = gpu.Chain(wl.config, 0, subtrain_frames=1 << 22)
ch.seek(0)
n_frames = 1 << 30
nch = n_frames // 16384
g = torch.Generator(device="cuda").manual_seed(1)
peaks = (0.25 + 0.2 * torch.rand(nch, device="cuda", generator=g)).float()
st = gpu.agc_initial_state()
st.locked, st.gain, st.peak_memory, st.samples_seen, st.last_strong_s = 1, 2.0, 0.45, 2000000, 2.0
def run(pk):
    ch.set_agc_state(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream()
    e0.record(s)
    for _ in range(10):
        ch.agc_advance_device(pk.data_ptr(), 0, n_frames, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
run(peaks)
print("quiet shard (65536 chunks, weak and strong chunks mixed, no event): %.1f us per shard" % (1e3 * run(peaks)))
loud = peaks.clone(); loud[nch // 2] = 0.9
print("same shard with one ratchet in the middle (tile walk): %.1f us per shard" % (1e3 * run(loud)))
print(ch.get_agc_state().as_tuple())
