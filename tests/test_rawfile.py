"""Raw-file streaming: file -> reader thread -> chain -> writer thread -> file."""
import os

import numpy as np
import pytest

from iq_tool_b200.configs import NUMPY_DTYPE
from iq_tool_b200.synth import synth_numpy
from oracle.loader import CpuChain

CHUNK = 16384


def test_rawfile_fails_loudly_without_device_or_file(tmp_path, workloads):
    from iq_tool_b200 import gpu
    cfg = workloads["cfg1"].config
    src = tmp_path / "in.cs16"
    src.write_bytes(b"\0" * 4096)
    if gpu.device_count() < 1:
        with pytest.raises(gpu.IqGpuError):               # no CPU fallback: chain creation needs a device
            gpu.rawfile_run(cfg, str(src), str(tmp_path / "out.cu8"))
    with pytest.raises(gpu.IqGpuError):
        gpu.rawfile_run(cfg, str(tmp_path / "missing.cs16"), str(tmp_path / "out.cu8"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,frames,train_chunks", [("cfg1", 70 * CHUNK + 4321, 16), ("cfg2", 40 * CHUNK, 7),
                                                      ("cfg3", 64 * CHUNK + 5, 0)])
def test_rawfile_run_equals_one_call_and_the_oracle(name, frames, train_chunks, tmp_path, gpu, workloads):
    """The streamed file equals one in-memory chain call byte for byte (train invariance) and the CPU oracle
    within the integer bar; a trailing partial frame in the file is dropped like input_rawfile.c:236 does."""
    import dataclasses
    wl = workloads[name]
    cfg = wl.config
    if cfg.dc_block:      # with a DC offset the reference's own fp32 integrator noise applies (DC-blocker exception)
        wl = dataclasses.replace(wl, dc=0.0)
    raw = synth_numpy(wl, frames)
    src, dst = tmp_path / "capture.raw", tmp_path / "out.raw"
    with open(src, "wb") as f:
        f.write(raw.tobytes())
        f.write(b"\x7f")                                  # stray trailing byte: not a whole frame
    st = gpu.rawfile_run(cfg, str(src), str(dst), train_chunks=train_chunks)
    out = np.fromfile(dst, dtype=NUMPY_DTYPE[cfg.output_format])
    one = gpu.Chain(cfg, 0).process(raw)
    assert st.frames_in == frames and st.frames_out * 2 == out.size and st.bytes_written == os.path.getsize(dst)
    if cfg.dc_block:      # the DC blocker's fp32 block sums depend on the call cut in the last bit (numerics)
        assert out.size == one.size and int(np.abs(out.astype(np.int32) - one.astype(np.int32)).max()) <= 1
    else:
        assert np.array_equal(out, one)
    ref = CpuChain(cfg, "oracle").process(raw)
    assert out.size == ref.size
    assert int(np.abs(out.astype(np.int32) - ref.astype(np.int32)).max()) <= 1
