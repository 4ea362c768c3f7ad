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


def parity_metrics(cfg, out, ref):
    """Measured deviation between a chain output and the reference's: integer formats in LSB (max, rms, share of frames
    off by more than one), cf32 as RMS of full scale and SNR.  Nothing is asserted here."""
    m = {"frames": int(len(out) // 2)}
    if cfg.output_format == "cf32":
        a, b = as_complex(out), as_complex(ref)
        d = a.astype(np.complex128) - b.astype(np.complex128)
        m.update(rel_rms=rel_rms_fullscale(a, b), snr_db=snr_db(a, b), max_abs=float(np.abs(d).max()) if d.size else 0.0,
                 ref_rms=float(np.sqrt(np.mean(np.abs(b.astype(np.complex128)) ** 2))) if b.size else 0.0)
    else:
        d = np.asarray(out).astype(np.int64) - np.asarray(ref).astype(np.int64)
        full = {"cs16": 32767.0, "cu16": 32767.0, "cs8": 127.0, "cu8": 127.0, "sc16q11": 2048.0}.get(cfg.output_format, 1.0)
        m.update(max_lsb=int(np.abs(d).max()) if d.size else 0, rms_lsb=float(np.sqrt(np.mean(d.astype(np.float64) ** 2))) if d.size else 0.0,
                 share_gt1=float((np.abs(d) > 1).mean()) if d.size else 0.0,
                 hist={int(k): int(v) for k, v in zip(*np.unique(np.abs(d), return_counts=True))},
                 rel_rms=float(np.sqrt(np.mean((d / full) ** 2))) if d.size else 0.0)
    return m


def error_spectrum_db(out, ref, cfg, nfft=4096):
    """Welch-style spectrum of (out - ref) in dB relative to full scale per bin (fftshifted), for the parity record:
    a white floor means rounding noise, a line or a slope means a systematic difference."""
    full = {"cs16": 32767.0, "cu16": 32767.0, "cs8": 127.0, "cu8": 127.0, "sc16q11": 2048.0}.get(cfg.output_format, 1.0)
    if cfg.output_format == "cf32":
        d = as_complex(out).astype(np.complex128) - as_complex(ref).astype(np.complex128)
    else:
        e = (np.asarray(out).astype(np.float64) - np.asarray(ref).astype(np.float64)) / full
        d = e[0::2] + 1j * e[1::2]
    k = d.size // nfft
    if k == 0:
        return None
    seg = d[: k * nfft].reshape(k, nfft) * np.hanning(nfft)
    p = np.mean(np.abs(np.fft.fftshift(np.fft.fft(seg, axis=1), axes=1)) ** 2, axis=0) / (nfft * np.sum(np.hanning(nfft) ** 2) / nfft)
    return 10.0 * np.log10(p / nfft + 1e-30)


def record_parity(key, metrics):
    """Append one measured-parity record to gpurun_out/parity_r02.jsonl (scratch on the GPU box, merged back by gpurun;
    the summary that is judged lives under profiles/)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_r02.jsonl"), "a") as f:
            f.write(json.dumps({"key": key, **metrics}) + "\n")
    except OSError:
        pass
    print(f"[parity] {key}: {json.dumps(metrics)}")
