"""solid::filter::firdes + solid::windows::kaiser + the solid::math functions they use -- host-side
f64 design helpers (firdes/mod.rs:243-305,329-364,443-526; windows/kaiser.rs:33-46;
math/mod.rs:17-27,41-100,156-183).  Design runs once on the host in the reference too; it is
restated here so product code can build its taps without touching oracle/."""
from __future__ import annotations

import math

BESSEL_ITERATIONS = 64  # math/mod.rs:8


class FirdesError(ValueError):
    pass


class WindowError(ValueError):
    pass


def sinc(x: float) -> float:  # math/mod.rs:17-27
    if abs(x) < 0.01:
        return math.cos(math.pi * x / 2.0) * math.cos(math.pi * x / 4.0) * math.cos(math.pi * x / 8.0)
    return math.sin(math.pi * x) / (math.pi * x)


def lngamma(x: float) -> float:  # math/mod.rs:171-183
    if x < 0.0:
        return 0.0
    if x < 10.0:
        return lngamma(x + 1.0) - math.log(x)
    g = 0.5 * (math.log(2.0 * math.pi) - math.log(x))
    return g + x * (math.log(x + (1.0 / (12.0 * x - 0.1 / x))) - 1.0)


def gamma(x: float) -> float:  # math/mod.rs:156-169
    if x < 0.0:
        return math.pi / (gamma(1.0 - x) * math.sin(math.pi * x))
    return math.exp(lngamma(x))


def lnbesseli(z: float, nu: float) -> float:  # math/mod.rs:66-100
    if z == 0.0:
        return 0.0 if nu == 0.0 else -1.7976931348623157e308
    if nu == 0.5:
        return 0.5 * math.log(2.0 / (math.pi * z)) + math.log(math.sinh(z))
    if z < 0.001 * math.sqrt(nu + 1.0):
        return -gamma(nu + 1.0) + nu * math.log(0.5 * z)
    t0 = nu * math.log(0.5 * z)
    y = 0.0
    for k in range(BESSEL_ITERATIONS):
        t1 = 2.0 * float(k) * math.log(0.5 * z)
        t2 = lngamma(float(k) + 1.0)
        t3 = lngamma(nu + float(k) + 1.0)
        y += math.exp(t1 - t2 - t3)
    return t0 + math.log(y)


def besseli(z: float, nu: float) -> float:  # math/mod.rs:41-64
    if z == 0.0:
        return 1.0 if nu == 0.0 else 0.0
    if nu == 0.5:
        return math.sqrt(2.0 / (math.pi * z)) * math.sinh(z)
    if z < 0.001 * math.sqrt(nu + 1.0):
        return math.pow(0.5 * z, nu) / gamma(nu + 1.0)
    return math.exp(lnbesseli(z, nu))


def kaiser(index: int, window_length: int, beta: float) -> float:  # windows/kaiser.rs:33-46
    if index > window_length:
        raise WindowError("OutOfBounds")
    if beta < 0.0:
        raise WindowError("BetaLessThanZero")
    t = float(index) - float(window_length - 1) / 2.0
    r = 2.0 * t / float(window_length - 1)
    return besseli(beta * math.sqrt(1.0 - r * r), 0.0) / besseli(beta, 0.0)


def kaiser_beta(stop_band_attenuation: float) -> float:  # firdes/mod.rs:243-253
    a = abs(stop_band_attenuation)
    if a > 50.0:
        return 0.1102 * (a - 8.7)
    if a > 21.0:
        return 0.5842 * math.pow(a - 21.0, 0.4) + 0.07886 * (a - 21.0)
    return 0.0


def firdes_kaiser(filter_length: int, cutoff_frequency: float, stop_band_attenuation: float,
                  fractional_sample_offset: float = 0.0) -> list:  # firdes/mod.rs:278-305
    if not (-0.5 <= fractional_sample_offset <= 0.5):
        raise FirdesError("Mu")
    if not (0.0 <= cutoff_frequency <= 0.5):
        raise FirdesError("Bandwidth")
    if stop_band_attenuation <= 0.0:
        raise FirdesError("StopBandLevel")
    beta = kaiser_beta(stop_band_attenuation)
    h = []
    for i in range(filter_length):
        t = float(i) - float(filter_length - 1) / 2.0 + fractional_sample_offset
        h.append(sinc(2.0 * cutoff_frequency * t) * kaiser(i, filter_length, beta))
    return h


def firdes_notch(semi_length: int, notch_frequency: float, stop_band_attenuation: float) -> list:
    """firdes/mod.rs:329-364"""
    if not (1 <= semi_length <= 1000):
        raise FirdesError("SemiLength")
    if not (0.0 <= notch_frequency <= 0.5):
        raise FirdesError("Bandwidth")
    if stop_band_attenuation <= 0.0:
        raise FirdesError("StopBandLevel")
    beta = kaiser_beta(stop_band_attenuation)
    n = 2 * semi_length + 1
    h, scale = [], 0.0
    for i in range(n):
        tone = -math.cos(2.0 * math.pi * notch_frequency * (float(i) - float(semi_length)))
        w = kaiser(i, n, beta)
        h.append(tone * w)
        scale += h[-1] * tone
    h = [c / scale for c in h]
    h[semi_length] += 1.0
    return h


def filter_autocorrelation(h, lag: int) -> float:  # firdes/mod.rs:443-456
    lag = abs(lag)
    if lag >= len(h):
        return 0.0
    r = 0.0
    for i in range(lag, len(h)):
        r += h[i] * h[i - lag]
    return r


def filter_crosscorrelation(h, g, lag: int) -> float:  # firdes/mod.rs:487-526
    if len(h) < len(g):
        return filter_crosscorrelation(g, h, lag)
    if lag <= -len(g) or lag >= len(h):
        return 0.0
    ig = -lag if lag < 0 else 0
    ih = lag if lag > 0 else 0
    if lag < 0:
        n = len(g) + lag
    elif lag < len(h) - len(g):
        n = len(g)
    else:
        n = len(h) - lag
    r = 0.0
    for i in range(n):
        r += h[ih + i] * g[ig + i]
    return r
