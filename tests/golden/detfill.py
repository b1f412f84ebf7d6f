"""Deterministic inputs / weights shared by the golden-vector generator and the tests.

Nothing here depends on torch's RNG stream or module construction order: every tensor is a
function of its *name* and shape only (numpy MT19937 seeded with crc32(name)), so the reference
module (in make_golden.py) and the module under test get bit-identical parameters by key.
"""
import zlib

import numpy as np


def det_array(name, shape, scale=1.0, kind="normal"):
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    if kind == "normal":
        a = rs.standard_normal(size=shape)
    else:
        a = rs.uniform(-1.0, 1.0, size=shape)
    return (a * scale).astype(np.float32)


def det_param(name, shape):
    """weights ~ N(0, 0.05), LayerNorm/BN weights ~ 1 + N(0, 0.1), biases ~ N(0, 0.05)."""
    leaf = name.split(".")[-1]
    is_norm = (".norm" in name or name.startswith("norm") or ".1." in name) and len(shape) == 1
    if is_norm and leaf == "weight":
        return 1.0 + det_array(name, shape, 0.1)
    if "running_var" in name:
        return np.ones(shape, np.float32)
    if "running_mean" in name:
        return np.zeros(shape, np.float32)
    return det_array(name, shape, 0.05)


def fill_state_dict(sd):
    """returns {key: np.ndarray} for every float tensor in a state dict (ints left alone)."""
    out = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            out[k] = det_param(k, tuple(v.shape))
    return out


def signal(kind, n, sr=16000):
    t = np.arange(n, dtype=np.float64) / sr
    if kind == "noise":
        return det_array("noise%d" % n, (n,), 0.1)
    if kind == "sine_silence":  # tone then digital silence: most bins hit the top_db clamp
        x = 0.5 * np.sin(2 * np.pi * 440.0 * t)
        x[n // 2:] = 0.0
        return x.astype(np.float32)
    if kind == "chirp":  # decaying chirp 100 Hz -> 6 kHz
        f = 100.0 + (6000.0 - 100.0) * t / max(t[-1], 1e-9)
        x = 0.8 * np.sin(2 * np.pi * np.cumsum(f) / sr) * np.exp(-3.0 * t / max(t[-1], 1e-9))
        return x.astype(np.float32)
    if kind == "zeros":
        return np.zeros(n, np.float32)
    if kind == "impulses":
        x = np.zeros(n, np.float32)
        x[::997] = 1.0
        return x
    raise ValueError(kind)


def summarize(a, n_head=64, n_stride=257):
    """compact, order-sensitive summary of a big array for fixtures."""
    a = np.asarray(a, np.float32).ravel()
    idx = np.arange(0, a.size, max(1, a.size // n_stride))[:n_stride]
    return {"sum": np.float64(a.astype(np.float64).sum()), "abs": np.float64(np.abs(a).astype(np.float64).sum()),
            "sq": np.float64((a.astype(np.float64) ** 2).sum()), "head": a[:n_head].copy(), "samp": a[idx].copy(),
            "idx": idx.astype(np.int64)}
