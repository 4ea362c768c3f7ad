"""Parity metrics shared by the tests (tolerances are stated where they are used)."""
import numpy as np


def rel_rms_fullscale(a, b):
    """RMS of (a-b) relative to full scale 1.0 (cf32 streams are normalised to +-1)."""
    d = np.asarray(a).astype(np.complex128) - np.asarray(b).astype(np.complex128)
    return float(np.sqrt(np.mean(np.abs(d) ** 2))) if d.size else 0.0


def snr_db(a, ref):
    d = np.asarray(a).astype(np.complex128) - np.asarray(ref).astype(np.complex128)
    num = float(np.mean(np.abs(np.asarray(ref).astype(np.complex128)) ** 2))
    den = float(np.mean(np.abs(d) ** 2))
    return float("inf") if den == 0 else 10.0 * np.log10(num / den)


def max_lsb(a, b):
    return int(np.abs(np.asarray(a).astype(np.int64) - np.asarray(b).astype(np.int64)).max()) if len(a) else 0


def as_complex(raw_cf32):
    return np.asarray(raw_cf32, dtype=np.float32).view(np.complex64)
