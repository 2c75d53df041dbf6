"""Shared helpers for the parity tests (oracle = checker, GPU library = thing under test)."""
import numpy as np

TOL = 1e-5  # BASELINE.json north_star: max normalised error <= 1e-5 in f32


def rand_cf32(rng, shape):
    """Complex samples ~ U(-1,1) already rounded to f32 ("identical inputs" rule)."""
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)


def f32_taps(h):
    """Taps rounded once to f32 and promoted back: both sides consume the same values."""
    return np.asarray(h, dtype=np.float32).astype(np.float64)


def nerr(got, ref):
    """max_n |got - ref| / max_n |ref| (per call; SURVEY 8c tolerance metric)."""
    got = np.asarray(got, dtype=np.complex128).ravel()
    ref = np.asarray(ref, dtype=np.complex128).ravel()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    den = np.max(np.abs(ref))
    if den == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref)) / den)
