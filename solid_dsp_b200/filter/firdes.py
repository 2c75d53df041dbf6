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
    den = float(window_length - 1)
    r = 2.0 * t / den if den != 0.0 else float("nan")  # a 1-tap window: 0.0 / 0.0 = NaN in the reference's f64 arithmetic
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


def filter_isi(h, samples_per_symbol: int, filter_delay: int):  # firdes/mod.rs:553-573
    """(rms, max) inter-symbol interference of a filter of 2 * sps * delay + 1 taps; (0, 0) on a length mismatch."""
    if 2 * samples_per_symbol * filter_delay + 1 != len(h):
        return 0.0, 0.0
    rxx0 = filter_autocorrelation(h, 0)
    isi_rms, isi_max = 0.0, 0.0
    for i in range(1, 2 * filter_delay):
        e = abs(filter_autocorrelation(h, i * samples_per_symbol) / rxx0)
        isi_rms += e * e
        if i == 1 or e > isi_max:
            isi_max = e
    return math.sqrt(isi_rms / (2.0 * float(filter_delay))), isi_max


def _energy_checks(h, cutoff_frequency, fft_size):  # firdes/mod.rs:608-614
    if not (0.0 <= cutoff_frequency <= 0.5):
        raise FirdesError("Bandwidth")
    if len(h) == 0:
        raise FirdesError("FilterSize")
    if fft_size == 0:
        raise FirdesError("FFTSize")


def filter_energy(h, cutoff_frequency: float, fft_size: int) -> float:
    """Relative out-of-band energy, firdes/mod.rs:603-640, host f64: the crate's own caller of `DotProduct::execute`
    (FORWARD coefficients against e^{j 2 pi f k}, `sum += value * sample` sequentially, dot_product/mod.rs:159-170)."""
    _energy_checks(h, cutoff_frequency, fft_size)
    e_total, e_stop = 0.0, 0.0
    for i in range(fft_size):
        f = 0.5 * float(i) / float(fft_size)
        re, im = 0.0, 0.0
        for k, c in enumerate(h):
            th = 2.0 * math.pi * f * float(k)
            re += c * (1.0 * math.cos(th))
            im += c * (1.0 * math.sin(th))
        e2 = re * re - im * (-im)
        e_total += e2
        if f > cutoff_frequency:
            e_stop += e2
    return e_stop / e_total


def filter_energy_device(h, cutoff_frequency: float, fft_size: int) -> float:
    """filter_energy with its `DotProduct::execute` calls on the GPU: the fft_size sample vectors e^{j 2 pi f k} are ONE
    batched sgpu_dot_execute (f32 on the device, so the result carries f32 rounding: ~1e-6 relative)."""
    import numpy as np
    from ..dot_product import Direction, DotProduct
    _energy_checks(h, cutoff_frequency, fft_size)
    f = 0.5 * np.arange(fft_size, dtype=np.float64) / float(fft_size)
    ejwt = np.exp(2j * np.pi * f[:, None] * np.arange(len(h), dtype=np.float64)[None, :]).astype(np.complex64)
    v = np.asarray(DotProduct(list(h), Direction.FORWARD).execute(ejwt), dtype=np.complex128)
    e2 = (v * np.conj(v)).real
    return float(np.sum(e2[f > cutoff_frequency]) / np.sum(e2))


def firdes_kaiser_device(filter_length: int, cutoff_frequency, stop_band_attenuation, fractional_sample_offset=0.0,
                         device_out=None, stream=None):
    """firdes_kaiser (firdes/mod.rs:278-305) computed ON THE GPU (sgpu_firdes_kaiser, csrc/firdes.cu): one design, or
    one design per element when the parameters are sequences (a bank of per-channel filters in one launch).  Returns a
    list of taps (a list of lists for several designs), or fills `device_out` (a torch float64 CUDA tensor of shape
    [n_designs, filter_length]) and returns it.  Raises FirdesError with the reference's variants."""
    import ctypes as C
    from .. import _ffi
    many = hasattr(cutoff_frequency, "__len__")
    fc = [float(v) for v in (cutoff_frequency if many else [cutoff_frequency])]
    n = len(fc)

    def arr(v):
        vals = [float(t) for t in v] if hasattr(v, "__len__") else [float(v)] * n
        if len(vals) != n:
            raise ValueError("firdes_kaiser_device: parameter sequences of different lengths")
        return (C.c_double * n)(*vals)
    a_fc, a_as, a_mu = arr(fc), arr(stop_band_attenuation), arr(fractional_sample_offset)
    if device_out is not None:
        assert device_out.is_cuda and device_out.numel() == n * filter_length and device_out.is_contiguous()
        out_ptr, mem = C.c_void_p(device_out.data_ptr()), _ffi.DEVICE
    else:
        host = (C.c_double * max(n * filter_length, 1))()
        out_ptr, mem = C.cast(host, C.c_void_p), _ffi.HOST
    st = _ffi.lib.sgpu_firdes_kaiser(filter_length, a_fc, a_as, a_mu, n, out_ptr, mem, stream)
    if st in (_ffi.ERR_FIRDES_MU, _ffi.ERR_FIRDES_BANDWIDTH, _ffi.ERR_FIRDES_STOP_BAND_LEVEL):
        raise FirdesError({_ffi.ERR_FIRDES_MU: "Mu", _ffi.ERR_FIRDES_BANDWIDTH: "Bandwidth",
                           _ffi.ERR_FIRDES_STOP_BAND_LEVEL: "StopBandLevel"}[st])
    _ffi.check(st)
    if device_out is not None:
        return device_out
    rows = [list(host[d * filter_length:(d + 1) * filter_length]) for d in range(n)]
    return rows if many else rows[0]
