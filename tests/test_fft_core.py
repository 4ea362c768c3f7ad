"""CPU unit test of the FFT butterfly network the CUDA FFT-filter kernel runs
(iq_tool_b200/csrc/fft_core.cuh compiled with g++): forward -> digit-reversed spectrum,
multiply, inverse == circular convolution, for every transform size class the kernel uses."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fftcore") / "libfftcore_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", out,
                    os.path.join(ROOT, "tests", "native", "fft_core_host.cpp")], check=True)
    lib = C.CDLL(out)
    vp, u = C.c_void_p, C.c_uint
    lib.fftcore_twiddles.argtypes = [vp, u]
    lib.fftcore_forward.argtypes = [vp, u, vp, u]
    lib.fftcore_inverse.argtypes = [vp, u, vp, u]
    lib.fftcore_forward_split.argtypes = [vp, u, u, vp]
    lib.fftcore_inverse_split.argtypes = [vp, u, u, vp]
    return lib


def _tw(core, n):
    tw = np.zeros(n, dtype=np.complex64)
    core.fftcore_twiddles(tw.ctypes.data, n)
    return tw


@pytest.mark.parametrize("m", [2, 4, 8, 16, 32, 128, 1024, 2048, 16384])
def test_forward_is_a_permuted_dft_and_inverse_undoes_it(core, m):
    rng = np.random.Generator(np.random.PCG64(m))
    x = (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)
    tw = _tw(core, m)
    f = x.copy()
    core.fftcore_forward(f.ctypes.data, m, tw.ctypes.data, m)
    ref = np.fft.fft(x.astype(np.complex128))
    # same multiset of bins (a permutation of the DFT)
    a = np.sort_complex(np.round(f.astype(np.complex128), 2))
    b = np.sort_complex(np.round(ref, 2))
    assert np.abs(np.sort(np.abs(f)) - np.sort(np.abs(ref))).max() <= 2e-5 * np.abs(ref).max() * np.log2(max(m, 2))
    assert a.size == b.size
    g = f.copy()
    core.fftcore_inverse(g.ctypes.data, m, tw.ctypes.data, m)
    assert np.abs(g / m - x).max() <= 1e-6 * np.log2(max(m, 2)) * max(1.0, np.abs(x).max())


@pytest.mark.parametrize("n,msub", [(256, 256), (16384, 16384), (8192, 1024), (65536, 16384), (4096, 64)])
def test_block_convolution_through_the_permuted_domain(core, n, msub):
    """IFFT(FFT(x) .* FFT(h)) with both spectra in the network's own order == circular convolution."""
    rng = np.random.Generator(np.random.PCG64(n + msub))
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    h = np.zeros(n, dtype=np.complex64)
    nt = n // 2 - 1
    h[:nt] = ((rng.standard_normal(nt) + 1j * rng.standard_normal(nt)) / nt).astype(np.complex64)
    tw = _tw(core, n)
    X, H = x.copy(), h.copy()
    core.fftcore_forward_split(X.ctypes.data, n, msub, tw.ctypes.data)
    core.fftcore_forward_split(H.ctypes.data, n, msub, tw.ctypes.data)
    Y = (X * H).astype(np.complex64)
    core.fftcore_inverse_split(Y.ctypes.data, n, msub, tw.ctypes.data)
    y = Y / n
    ref = np.fft.ifft(np.fft.fft(x.astype(np.complex128)) * np.fft.fft(h.astype(np.complex128)))
    err = np.sqrt(np.mean(np.abs(y - ref) ** 2)) / np.sqrt(np.mean(np.abs(ref) ** 2))
    assert err <= 1e-6
