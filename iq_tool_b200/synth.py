"""Deterministic synthetic I/Q input:

    x[n] = sum_k A_k exp(j(2 pi f_k n / Fs + phi_k)) + dc + sigma (w_I[n] + j w_Q[n])

optionally followed by an I/Q imbalance (cfg4) and quantised to the input sample format with
round-to-nearest and saturation.  `synth_numpy` is used for parity-sized vectors (fed to the
oracle and the GPU alike); `synth_torch` generates throughput-sized captures directly in HBM.
"""
from __future__ import annotations

import numpy as np

from .configs import Workload

_FULL_SCALE = {"cs16": 32767.0, "cs8": 127.0, "cu8": 127.0, "cu16": 32767.0, "sc16q11": 2047.0}


def _phases(k: int) -> float:
    return 0.7 * k + 0.1


def synth_complex_numpy(w: Workload, n: int, start: int = 0, seed: int = 12345) -> np.ndarray:
    """complex128 baseband signal for absolute sample indices [start, start+n)."""
    fs = float(int(w.config.input_rate_hz))
    idx = np.arange(start, start + n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    for k, (a, f) in enumerate(w.tones):
        x += a * np.exp(1j * (2.0 * np.pi * (f / fs) * idx + _phases(k)))
    x += w.dc * (1.0 + 1.0j)
    rng = np.random.Generator(np.random.PCG64(seed if start == 0 else [seed, start]))
    x += w.sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if w.iq_imbalance:
        i, q = x.real, x.imag
        ph = np.deg2rad(3.0)
        x = i + 1j * (1.05 * (q * np.cos(ph) + i * np.sin(ph)))
    return x


def quantise(x: np.ndarray, fmt: str) -> np.ndarray:
    """complex -> interleaved raw samples of `fmt` (numpy array, length 2n)."""
    iq = np.empty(2 * x.shape[0], dtype=np.float64)
    iq[0::2], iq[1::2] = x.real, x.imag
    if fmt == "cs16":
        return np.clip(np.rint(iq * 32767.0), -32768, 32767).astype(np.int16)
    if fmt == "sc16q11":
        return np.clip(np.rint(iq * 2047.0), -2048, 2047).astype(np.int16)
    if fmt == "cs8":
        return np.clip(np.rint(iq * 127.0), -128, 127).astype(np.int8)
    if fmt == "cu8":
        return np.clip(np.rint(iq * 127.0 + 127.5), 0, 255).astype(np.uint8)
    if fmt == "cu16":
        return np.clip(np.rint(iq * 32767.0 + 32767.5), 0, 65535).astype(np.uint16)
    if fmt == "cs32":
        return np.clip(np.rint(iq * 2147483647.0), -2147483648, 2147483647).astype(np.int32)
    if fmt == "cu32":
        return np.clip(np.rint(iq * 2147483647.0 + 2147483647.5), 0, 4294967295).astype(np.uint32)
    if fmt == "cf32":
        return iq.astype(np.float32)
    if fmt == "cs24":
        v = np.clip(np.rint(iq * 8388607.0), -8388608, 8388607).astype(np.int32)
        out = np.empty(3 * v.shape[0], dtype=np.uint8)
        out[0::3] = v & 0xFF
        out[1::3] = (v >> 8) & 0xFF
        out[2::3] = (v >> 16) & 0xFF
        return out
    raise ValueError(fmt)


def synth_numpy(w: Workload, n: int, start: int = 0, seed: int = 12345) -> np.ndarray:
    return quantise(synth_complex_numpy(w, n, start, seed), w.config.input_format)


def synth_torch(w: Workload, n: int, device, start: int = 0, seed: int = 20261017, block: int = 1 << 24):
    """Throughput-sized raw capture generated in HBM (torch RNG keyed by (seed, block index))."""
    import torch

    fmt = w.config.input_format
    fs = float(int(w.config.input_rate_hz))
    dt = {"cs16": torch.int16, "cu8": torch.uint8, "cs8": torch.int8}[fmt]
    out = torch.empty(2 * n, dtype=dt, device=device)
    gen = torch.Generator(device=device)
    for b0 in range(0, n, block):
        m = min(block, n - b0)
        gen.manual_seed(seed + (start + b0) // block)
        idx = torch.arange(start + b0, start + b0 + m, device=device, dtype=torch.float64)
        re = torch.zeros(m, device=device, dtype=torch.float32)
        im = torch.zeros(m, device=device, dtype=torch.float32)
        for k, (a, f) in enumerate(w.tones):
            ph = torch.remainder(idx * (f / fs), 1.0) * (2.0 * np.pi) + _phases(k)
            ph = ph.to(torch.float32)
            re += a * torch.cos(ph)
            im += a * torch.sin(ph)
        re += w.dc + w.sigma * torch.randn(m, device=device, generator=gen)
        im += w.dc + w.sigma * torch.randn(m, device=device, generator=gen)
        if w.iq_imbalance:
            ph = float(np.deg2rad(3.0))
            im = 1.05 * (im * float(np.cos(ph)) + re * float(np.sin(ph)))
        if fmt == "cs16":
            q_re = torch.clamp(torch.round(re * 32767.0), -32768, 32767).to(dt)
            q_im = torch.clamp(torch.round(im * 32767.0), -32768, 32767).to(dt)
        elif fmt == "cs8":
            q_re = torch.clamp(torch.round(re * 127.0), -128, 127).to(dt)
            q_im = torch.clamp(torch.round(im * 127.0), -128, 127).to(dt)
        else:
            q_re = torch.clamp(torch.round(re * 127.0 + 127.5), 0, 255).to(dt)
            q_im = torch.clamp(torch.round(im * 127.0 + 127.5), 0, 255).to(dt)
        out[2 * b0:2 * (b0 + m):2] = q_re
        out[2 * b0 + 1:2 * (b0 + m):2] = q_im
    return out
