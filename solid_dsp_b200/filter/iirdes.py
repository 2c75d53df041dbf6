"""solid::filter::iirdes::pll (iirdes/pll/mod.rs:24-99) plus the benchmark's biquad table.

The reference has no Butterworth/Chebyshev SOS designer, so the
8-section cascade of BASELINE config 5 is synthesised here from a fixed table of stable low-pass
sections: conjugate pole pairs of radius 0.50 .. 0.95, double zero at z = -1, unit DC gain."""
from __future__ import annotations

import math


class IirdesError(ValueError):
    pass


def active_lag(bandwidth: float, damping_factor: float, loop_gain: float):
    """iirdes/pll/mod.rs:24-52 -> (numerator[3], denominator[3])"""
    if bandwidth <= 0.0:
        raise IirdesError("Bandwidth")
    if damping_factor <= 0.0:
        raise IirdesError("DampingFactor")
    if loop_gain <= 0.0:
        raise IirdesError("Gain")
    t1 = loop_gain / (bandwidth * bandwidth)
    t2 = 2.0 * damping_factor / bandwidth - 1.0 / loop_gain
    num = [2.0 * loop_gain * (1.0 + t2 / 2.0), 2.0 * loop_gain * 2.0, 2.0 * loop_gain * (1.0 - t2 / 2.0)]
    den = [1.0 + t1 / 2.0, -t1, -1.0 + t1 / 2.0]
    return num, den


def active_proportional_integral(bandwidth: float, damping_factor: float, loop_gain: float):
    """iirdes/pll/mod.rs:54-99"""
    if bandwidth <= 0.0:
        raise IirdesError("Bandwidth")
    if damping_factor <= 0.0:
        raise IirdesError("DampingFactor")
    if loop_gain <= 0.0:
        raise IirdesError("Gain")
    t1 = loop_gain / (bandwidth * bandwidth)
    t2 = 2.0 * damping_factor / bandwidth - 1.0 / loop_gain
    num = [2.0 * loop_gain * (1.0 + t2 / 2.0), 2.0 * loop_gain * 2.0, 2.0 * loop_gain * (1.0 - t2 / 2.0)]
    den = [t1 / 2.0, -t1, t1 / 2.0]
    return num, den


# (pole radius, pole angle / pi) of the benchmark's sections
_SECTION_TABLE = [(0.50, 0.10), (0.60, 0.14), (0.70, 0.18), (0.78, 0.22),
                  (0.84, 0.26), (0.88, 0.30), (0.92, 0.34), (0.95, 0.38),
                  (0.55, 0.12), (0.65, 0.16), (0.74, 0.20), (0.81, 0.24),
                  (0.86, 0.28), (0.90, 0.32), (0.93, 0.36), (0.94, 0.40)]


def stable_lowpass_sections(n_sections: int = 8):
    """Flat (ff, fb) arrays of 3*n values, f32-representable, a0 = 1, poles inside |z| <= 0.95."""
    import numpy as np
    ff, fb = [], []
    for r, th in _SECTION_TABLE[:n_sections]:
        a1 = -2.0 * r * math.cos(math.pi * th)
        a2 = r * r
        g = (1.0 + a1 + a2) / 4.0  # unit gain at DC with zeros at z = -1
        ff += [g, 2.0 * g, g]
        fb += [1.0, a1, a2]
    ff = np.asarray(ff, dtype=np.float32).astype(np.float64)
    fb = np.asarray(fb, dtype=np.float32).astype(np.float64)
    return ff, fb
