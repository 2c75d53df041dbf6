"""solid::group_delay (group_delay/mod.rs:51-129) -- host-side f64 analysis."""
from __future__ import annotations

import cmath
import math

TOLERANCE = 0.00000000001  # group_delay/mod.rs:24


class DelayError(Exception):
    """DelayError(DelayErrorCode) -- group_delay/mod.rs:26-47"""

    def __init__(self, code: str):
        self.code = code
        super().__init__({"EmptyCoefficients": "Delay Error: Empty Coefficients",
                          "FrequencyOutOfBounds": "Delay Error: Frequency Out of Bounds [-0.5, 0.5]",
                          "DivideByZero": "Delay Error: Denominator Coefficents Divide Numerator by Zero"}[code])


def fir_group_delay(coefs, frequency: float) -> float:
    """group_delay/mod.rs:51-79"""
    coefs = list(coefs)
    if not coefs:
        raise DelayError("EmptyCoefficients")
    if frequency < -0.5 or frequency > 0.5:
        raise DelayError("FrequencyOutOfBounds")
    t0 = 0j
    t1 = 0j
    for i, c in enumerate(coefs):
        rot = cmath.rect(1.0, frequency * 2.0 * math.pi * i)
        t0 += c * rot * float(i)
        t1 += c * rot
    return (t0 / t1).real


def iir_group_delay(num, den, frequency: float) -> float:
    """group_delay/mod.rs:82-129"""
    num, den = list(num), list(den)
    if not num or not den:
        raise DelayError("EmptyCoefficients")
    if frequency < -0.5 or frequency > 0.5:
        raise DelayError("FrequencyOutOfBounds")
    n = len(num) + len(den) - 1
    coefs = [0.0] * n
    for i in range(len(den)):
        for j in range(len(num)):
            d = den[len(den) - i - 1]
            d = d.conjugate() if isinstance(d, complex) else d
            coefs[i + j] = coefs[i + j] + d * num[j]
    t0 = 0j
    t1 = 0j
    for i, c in enumerate(coefs):
        c0 = c * cmath.rect(1.0, frequency * 2.0 * math.pi * i)
        t0 += c0 * float(i)
        t1 += c0
    if math.hypot(t1.real, t1.imag) <= TOLERANCE:
        raise DelayError("DivideByZero")
    return (t0 / t1).real - float(len(den) - 1)
