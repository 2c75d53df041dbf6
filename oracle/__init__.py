"""ctypes front-end for the CPU oracle (oracle/solid_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package (solid_dsp_b200) never imports this.

Complex arrays are numpy complex128 on this side and interleaved doubles on the C side.
Class and method names follow the reference (FIRFilter.execute_block, ...) so tests read like
the reference's doc-tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "solid_oracle.c"

c_size = C.c_size_t
c_dp = C.POINTER(C.c_double)


def build(native: bool = False, force: bool = False) -> Path:
    """Compile the oracle.  native=True builds a -march=native copy for timing on this host."""
    out = _HERE / "_build" / ("libsolid_oracle_native.so" if native else "libsolid_oracle.so")
    if out.exists() and not force and out.stat().st_mtime >= _SRC.stat().st_mtime and not native:
        return out
    if native and out.exists() and not force and out.stat().st_mtime >= _SRC.stat().st_mtime:
        stamp = out.with_suffix(".host")
        if stamp.exists() and stamp.read_text() == _host_id():
            return out
    env = dict(os.environ)
    cmd = ["make", "-C", str(_HERE), f"OUT={out.relative_to(_HERE)}", "-B"]
    if native:
        cmd.append("MARCH=native")
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
    if native:
        out.with_suffix(".host").write_text(_host_id())
    return out


def _host_id() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.strip()
    except OSError:
        pass
    return "unknown"


_libs: dict = {}


def lib(native: bool = False):
    if native in _libs:
        return _libs[native]
    L = C.CDLL(str(build(native)))
    vp = C.c_void_p
    ip = C.POINTER(C.c_int)

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("so_msb_index", c_size, c_size)
    sig("so_dot_execute", None, c_dp, c_size, C.c_int, C.c_int, c_dp, c_size, c_dp)
    sig("so_dot_coefficients", None, c_dp, c_size, C.c_int, C.c_int, c_dp)
    sig("so_fir_new", vp, c_dp, c_size, C.c_int, C.c_double, C.c_double, C.c_int, c_size, ip)
    sig("so_fir_free", None, vp)
    sig("so_fir_set_scale", None, vp, C.c_double, C.c_double)
    sig("so_fir_window_capacity", c_size, vp)
    sig("so_fir_current_item", c_size, vp)
    sig("so_fir_write", None, vp, c_dp, c_size)
    sig("so_fir_execute_block", c_size, vp, c_dp, c_size, c_dp, c_size)
    sig("so_fir_fast", c_size, c_dp, c_size, C.c_int, C.c_double, C.c_double, c_size, c_size,
        c_dp, c_dp, c_size, c_dp)
    sig("so_pfb_new", vp, c_dp, c_size, C.c_int, c_size, C.c_double, C.c_double, ip)
    sig("so_pfb_free", None, vp)
    sig("so_pfb_push", None, vp, C.c_double, C.c_double)
    sig("so_pfb_execute", None, vp, c_size, c_dp)
    sig("so_pfb_sub_len", c_size, vp)
    sig("so_pfb_coefficients", None, vp, c_dp)
    sig("so_interp_sub_len", c_size, c_size, c_size)
    sig("so_firinterp_new", vp, c_dp, c_size, C.c_int, c_size, ip)
    sig("so_firinterp_free", None, vp)
    sig("so_firinterp_set_scale", None, vp, C.c_double, C.c_double)
    sig("so_firinterp_coefficients", None, vp, c_dp)
    sig("so_firinterp_sub_len", c_size, vp)
    sig("so_firinterp_execute_block", c_size, vp, c_dp, c_size, c_dp, c_size)
    sig("so_firinterp_fast", c_size, c_dp, c_size, C.c_int, c_size, c_dp, c_dp, c_size, c_dp)
    sig("so_sos_new", vp, c_dp, c_size, c_dp, c_size, ip)
    sig("so_sos_free", None, vp)
    sig("so_sos_execute", None, vp, C.c_double, C.c_double, c_dp)
    sig("so_sos_numerator_coefs", None, vp, c_dp)
    sig("so_sos_denominator_coefs", None, vp, c_dp)
    sig("so_iir_new", vp, c_dp, c_size, c_dp, c_size, C.c_int, C.c_int, c_size, ip)
    sig("so_iir_free", None, vp)
    sig("so_iir_execute_block", c_size, vp, c_dp, c_size, c_dp, c_size)
    sig("so_sos_cascade_fast", None, c_dp, c_dp, c_size, c_dp, c_dp, c_size, c_dp)
    sig("so_autocorr_new", vp, c_size, c_size)
    sig("so_autocorr_free", None, vp)
    sig("so_autocorr_push", None, vp, C.c_double, C.c_double)
    sig("so_autocorr_execute", None, vp, c_dp)
    sig("so_autocorr_execute_block", c_size, vp, c_dp, c_size, c_dp)
    sig("so_autocorr_get_energy", C.c_double, vp)
    sig("so_autocorr_fast", None, c_size, c_size, c_dp, c_size, c_dp, c_size, c_dp)
    sig("so_sinc", C.c_double, C.c_double)
    sig("so_lngamma", C.c_double, C.c_double)
    sig("so_gamma", C.c_double, C.c_double)
    sig("so_besseli", C.c_double, C.c_double, C.c_double)
    sig("so_lnbesseli", C.c_double, C.c_double, C.c_double)
    sig("so_window_kaiser", C.c_double, c_size, c_size, C.c_double)
    sig("so_kaiser_beta", C.c_double, C.c_double)
    sig("so_firdes_kaiser", C.c_int, c_size, C.c_double, C.c_double, C.c_double, c_dp)
    sig("so_firdes_notch", C.c_int, c_size, C.c_double, C.c_double, c_dp)
    sig("so_filter_autocorrelation", C.c_double, c_dp, c_size, C.c_ssize_t)
    sig("so_filter_crosscorrelation", C.c_double, c_dp, c_size, c_dp, c_size, C.c_ssize_t)
    sig("so_filter_isi", None, c_dp, c_size, c_size, c_size, c_dp)
    sig("so_filter_energy", C.c_int, c_dp, c_size, C.c_double, c_size, c_dp)
    sig("so_pll_active_lag", C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_dp)
    sig("so_nco_constrain", C.c_uint32, C.c_double)
    sig("so_nco_new", vp)
    sig("so_nco_free", None, vp)
    sig("so_nco_reset", None, vp)
    for nm in ("set_frequency", "adjust_frequency", "set_phase", "adjust_phase"):
        sig("so_nco_" + nm, None, vp, C.c_double)
    sig("so_nco_step", None, vp)
    sig("so_nco_set_raw", None, vp, C.c_uint32, C.c_uint32)
    sig("so_nco_theta", C.c_uint32, vp)
    sig("so_nco_delta_theta", C.c_uint32, vp)
    sig("so_nco_sin", C.c_double, vp)
    sig("so_nco_cos", C.c_double, vp)
    sig("so_nco_mix", None, vp, C.c_int, C.c_double, C.c_double, c_dp)
    sig("so_nco_mix_block", None, vp, C.c_int, c_dp, c_size, c_dp)
    sig("so_run_units", c_size, C.c_int, c_dp, c_size, C.c_int, C.c_double, C.c_double, C.c_int,
        c_size, c_dp, c_dp, c_size, c_dp, c_size, c_size, c_size, c_dp, c_size, c_size, C.c_int)
    _libs[native] = L
    return L


# ---------------------------------------------------------------------------------------------
class OracleError(Exception):
    """Construction error; .code is the reference enum variant name."""

    NAMES = {
        -1: "FIRErrorCode::CoefficientsLengthZero",
        -2: "FIRErrorCode::DecimationLessThanOne",
        -3: "FIRErrorCode::InterpolationLessThanOne",
        -4: "FIRErrorCode::NotEnoughFilters",
        -10: "IIRErrorCode::NumeratorLengthZero",
        -11: "IIRErrorCode::DenominatorLengthZero",
        -12: "IIRErrorCode::SecondOrderSectionSizeZero",
        -13: "IIRErrorCode::SecondOrderSectionSizeMismatch",
        -14: "IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3",
        -15: "IIRErrorCode::DecimationLessThanOne",
        -16: "IIRErrorCode::InterpolationLessThanOne",
        -17: "SecondOrderErrorCode::CoefficientsNotInRange",
    }

    def __init__(self, err: int):
        self.err = err
        self.code = self.NAMES.get(err, str(err))
        super().__init__(self.code)


def _coefs(coefs):
    a = np.asarray(coefs)
    if np.iscomplexobj(a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        return a, 1, a.view(np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, 0, a


def _cx(x):
    a = np.ascontiguousarray(np.asarray(x), dtype=np.complex128)
    return a, a.view(np.float64)


def _p(a):
    if a is None:
        return None
    if a.size == 0:
        return (C.c_double * 2)()
    return a.ctypes.data_as(c_dp)


def _scale(s):
    s = complex(s)
    return s.real, s.imag


class DotProduct:
    """dot_product/mod.rs:37-171"""
    FORWARD, REVERSE = 0, 1

    def __init__(self, coefficients, direction):
        self._c, self._cc, self._cv = _coefs(coefficients)
        self._dir = direction

    def len(self):
        return len(self._c)

    def is_empty(self):
        return len(self._c) == 0

    def coefficents(self):
        out = np.zeros_like(self._c)
        lib().so_dot_coefficients(_p(self._cv), len(self._c), self._cc, self._dir,
                                  _p(out.view(np.float64)))
        return out

    def execute(self, samples):
        x, xv = _cx(samples)
        out = np.zeros(2)
        lib().so_dot_execute(_p(self._cv), len(self._c), self._cc, self._dir, _p(xv), len(x), _p(out))
        return complex(out[0], out[1])


class _FirBase:
    def __init__(self, coefs, scale, is_decim, decimation, native=False):
        self._L = lib(native)
        c, cc, cv = _coefs(coefs)
        err = C.c_int(0)
        sr, si = _scale(scale)
        self._h = self._L.so_fir_new(_p(cv), len(c), cc, sr, si, is_decim, decimation, C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        self._T = len(c)
        self._M = decimation if is_decim else 1

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_fir_free(self._h)
            self._h = None

    def set_scale(self, s):
        self._L.so_fir_set_scale(self._h, *_scale(s))

    def len(self):
        return self._T

    def window_capacity(self):
        return self._L.so_fir_window_capacity(self._h)

    def execute_block(self, samples):
        x, xv = _cx(samples)
        cap = len(x) + 1
        out = np.zeros(cap, dtype=np.complex128)
        n = self._L.so_fir_execute_block(self._h, _p(xv), len(x), _p(out.view(np.float64)), cap)
        return out[:n].copy()

    def execute(self, sample):
        return self.execute_block([sample])


class FIRFilter(_FirBase):
    """fir/mod.rs:58-241"""

    def __init__(self, coefs, scale=1.0, native=False):
        super().__init__(coefs, scale, 0, 0, native)


class DecimatingFIRFilter(_FirBase):
    """fir/decim.rs:5-256"""

    def __init__(self, coefs, scale, decimation, native=False):
        super().__init__(coefs, scale, 1, decimation, native)

    def get_decimation(self):
        return self._M

    def write(self, samples):
        x, xv = _cx(samples)
        self._L.so_fir_write(self._h, _p(xv), len(x))

    def push(self, sample):
        self.write([sample])

    def current_item(self):
        return self._L.so_fir_current_item(self._h)


class PolyPhaseFilterBank:
    """fir/pfb.rs:3-90"""

    def __init__(self, coefs, filters, scale=1.0):
        self._L = lib()
        c, cc, cv = _coefs(coefs)
        self._cc = cc
        err = C.c_int(0)
        self._h = self._L.so_pfb_new(_p(cv), len(c), cc, filters, *_scale(scale), C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        self._filters = filters

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_pfb_free(self._h)
            self._h = None

    def len(self):
        return self._filters

    def sub_len(self):
        return self._L.so_pfb_sub_len(self._h)

    def push(self, sample):
        s = complex(sample)
        self._L.so_pfb_push(self._h, s.real, s.imag)

    def execute(self, index):
        out = np.zeros(2)
        self._L.so_pfb_execute(self._h, index, _p(out))
        return complex(out[0], out[1])

    def coefficents(self):
        n = self._filters * self.sub_len()
        out = np.zeros(n, dtype=np.complex128 if self._cc else np.float64)
        self._L.so_pfb_coefficients(self._h, _p(out.view(np.float64)))
        return out.reshape(self._filters, -1)


class InterpolatingFIRFilter:
    """fir/interp.rs:6-111"""

    def __init__(self, coefs, interpolation, native=False):
        self._L = lib(native)
        c, cc, cv = _coefs(coefs)
        err = C.c_int(0)
        self._h = self._L.so_firinterp_new(_p(cv), len(c), cc, interpolation, C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        self._Lf = interpolation

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_firinterp_free(self._h)
            self._h = None

    def interpolation(self):
        return self._Lf

    def len(self):
        return self._Lf

    def sub_len(self):
        return self._L.so_firinterp_sub_len(self._h)

    def set_scale(self, s):
        self._L.so_firinterp_set_scale(self._h, *_scale(s))

    def execute_block(self, samples):
        x, xv = _cx(samples)
        cap = len(x) * self._Lf + 1
        out = np.zeros(cap, dtype=np.complex128)
        n = self._L.so_firinterp_execute_block(self._h, _p(xv), len(x), _p(out.view(np.float64)), cap)
        return out[:n].copy()

    def execute(self, sample):
        return self.execute_block([sample])


class SecondOrderFilter:
    """iir/sos.rs:34-114"""

    def __init__(self, ff, fb):
        self._L = lib()
        ff = np.ascontiguousarray(ff, dtype=np.float64)
        fb = np.ascontiguousarray(fb, dtype=np.float64)
        err = C.c_int(0)
        self._h = self._L.so_sos_new(_p(ff), len(ff), _p(fb), len(fb), C.byref(err))
        if not self._h:
            raise OracleError(err.value)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_sos_free(self._h)
            self._h = None

    def execute(self, sample):
        s = complex(sample)
        out = np.zeros(2)
        self._L.so_sos_execute(self._h, s.real, s.imag, _p(out))
        return complex(out[0], out[1])

    def numerator_coefs(self):
        out = np.zeros(2)
        self._L.so_sos_numerator_coefs(self._h, _p(out))
        return out

    def denominator_coefs(self):
        out = np.zeros(3)
        self._L.so_sos_denominator_coefs(self._h, _p(out))
        return out


NORMAL, SECOND_ORDER = 0, 1


class IIRFilter:
    """iir/mod.rs:68-316 (+ iir/decim.rs, iir/interp.rs through `wrapper`)"""

    def __init__(self, ff, fb, iirtype, _wrapper=0, _factor=0, native=False):
        self._L = lib(native)
        ff = np.ascontiguousarray(ff, dtype=np.float64)
        fb = np.ascontiguousarray(fb, dtype=np.float64)
        err = C.c_int(0)
        self._h = self._L.so_iir_new(_p(ff), len(ff), _p(fb), len(fb), iirtype, _wrapper, _factor,
                                     C.byref(err))
        if not self._h:
            raise OracleError(err.value)
        self._grow = _factor if _wrapper == 2 else 1

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_iir_free(self._h)
            self._h = None

    def execute_block(self, samples):
        x, xv = _cx(samples)
        cap = len(x) * self._grow + 1
        out = np.zeros(cap, dtype=np.complex128)
        n = self._L.so_iir_execute_block(self._h, _p(xv), len(x), _p(out.view(np.float64)), cap)
        return out[:n].copy()

    def execute(self, sample):
        return self.execute_block([sample])


class DecimatingIIRFilter(IIRFilter):
    def __init__(self, ff, fb, iirtype, decimation):
        super().__init__(ff, fb, iirtype, 1, decimation)


class InterpolatingIIRFilter(IIRFilter):
    def __init__(self, ff, fb, iirtype, interpolation):
        super().__init__(ff, fb, iirtype, 2, interpolation)


class AutoCorrelator:
    """filter/auto_correlator/mod.rs:24-216 (structural mirror, incl. the Window delay quirk)"""

    def __init__(self, window_size, delay, native=False):
        self._L = lib(native)
        self._h = self._L.so_autocorr_new(window_size, delay)
        if not self._h:
            raise OracleError(-1)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_autocorr_free(self._h)
            self._h = None

    def push(self, sample):
        z = complex(sample)
        self._L.so_autocorr_push(self._h, z.real, z.imag)

    def write(self, samples):
        for z in np.asarray(samples).ravel():
            self.push(z)

    def execute(self):
        out = np.zeros(2)
        self._L.so_autocorr_execute(self._h, _p(out))
        return complex(out[0], out[1])

    def execute_block(self, samples):
        x, xv = _cx(samples)
        out = np.zeros(max(len(x), 1), dtype=np.complex128)
        n = self._L.so_autocorr_execute_block(self._h, _p(xv), len(x), _p(out.view(np.float64)))
        return out[:n].copy()

    def get_energy(self):
        return self._L.so_autocorr_get_energy(self._h)


# ----------------------------------------------------------------------------- closed forms
class NCO:
    """solid::nco::NCO (nco/mod.rs), structural restatement.  Parity unpinned: the reference has no NCO golden."""

    def __init__(self, native=False):
        self._L = lib(native)
        self._h = self._L.so_nco_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.so_nco_free(self._h)
            self._h = None

    def reset(self):
        self._L.so_nco_reset(self._h)

    def set_frequency(self, dt):
        self._L.so_nco_set_frequency(self._h, float(dt))

    def adjust_frequency(self, dt):
        self._L.so_nco_adjust_frequency(self._h, float(dt))

    def set_phase(self, phi):
        self._L.so_nco_set_phase(self._h, float(phi))

    def adjust_phase(self, dphi):
        self._L.so_nco_adjust_phase(self._h, float(dphi))

    def set_raw(self, theta, delta_theta):
        self._L.so_nco_set_raw(self._h, theta & 0xFFFFFFFF, delta_theta & 0xFFFFFFFF)

    def raw(self):
        return int(self._L.so_nco_theta(self._h)), int(self._L.so_nco_delta_theta(self._h))

    def step(self):
        self._L.so_nco_step(self._h)

    def sin(self):
        return float(self._L.so_nco_sin(self._h))

    def cos(self):
        return float(self._L.so_nco_cos(self._h))

    def sincos(self):
        return self.sin(), self.cos()

    def complex_exponential(self):
        return complex(self.cos(), self.sin())

    def _mix(self, up, sample):
        out = np.zeros(2)
        z = complex(sample)
        self._L.so_nco_mix(self._h, up, z.real, z.imag, _p(out))
        return complex(out[0], out[1])

    def mix_up(self, sample):
        return self._mix(1, sample)

    def mix_down(self, sample):
        return self._mix(0, sample)

    def _mix_block(self, up, samples):
        x, xv = _cx(samples)
        out = np.zeros_like(x)
        self._L.so_nco_mix_block(self._h, up, _p(xv), x.size, _p(out.view(np.float64)))
        return out

    def mix_up_block(self, samples):
        """y[i] = mix_up(x[i]); step() -- the loop nco/mod.rs:153-161 was written to be."""
        return self._mix_block(1, samples)

    def mix_down_block(self, samples):
        return self._mix_block(0, samples)


def nco_constrain(theta):
    return int(lib().so_nco_constrain(float(theta)))


def nco_mix_down_block(x, frequency=0.0, phase=0.0, n_threads=1, raw=None, up=False):
    """Rows of x through independent NCOs (all set to `frequency` / `phase`, or to raw = (theta, delta) words per row),
    mix_down (or mix_up) + step per sample.  Returns complex128 of x's shape."""
    x2 = np.atleast_2d(_cx(x)[0])
    out = np.zeros_like(x2)

    def one(i):
        n = NCO()
        if raw is not None:
            th, dl = raw[i] if np.ndim(raw) == 2 else raw
            n.set_raw(int(th), int(dl))
        else:
            n.set_frequency(frequency)
            n.set_phase(phase)
        out[i] = n._mix_block(1 if up else 0, x2[i])

    if n_threads > 1 and x2.shape[0] > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(one, range(x2.shape[0])))
    else:
        for i in range(x2.shape[0]):
            one(i)
    return out.reshape(np.shape(x)) if np.ndim(x) == 1 else out


def ddc_fast(coefs, x, scale, decimation, frequency=0.0, phase=0.0, count0=0, hist=None, raw=None):
    """NCO mix-down (structural) feeding the closed-form decimator: the checker for the fused DDC kernel."""
    mixed = nco_mix_down_block(x, frequency, phase, raw=raw)
    return fir_fast(coefs, mixed, scale, decimation, count0, hist)


def autocorr_fast(window_size, delay, x, hist=None):
    """out[n] = sum_{i < W-d} x[n-i] conj(x[n-d-i]); hist = samples preceding x[0], oldest first"""
    x, xv = _cx(x)
    out = np.zeros(max(len(x), 1), dtype=np.complex128)
    hv, nh = None, 0
    if hist is not None and len(hist):
        h, hv = _cx(hist)
        nh = len(h)
    lib().so_autocorr_fast(window_size, delay, _p(hv) if hv is not None else None, nh, _p(xv), len(x),
                           _p(out.view(np.float64)))
    return out[:len(x)].copy()


def fir_fast(coefs, x, scale=1.0, decimation=0, count0=0, hist=None):
    c, cc, cv = _coefs(coefs)
    x, xv = _cx(x)
    out = np.zeros(len(x) + 1, dtype=np.complex128)
    hv = None
    if hist is not None:
        h, hv = _cx(hist)
        assert len(h) == len(c) - 1
    n = lib().so_fir_fast(_p(cv), len(c), cc, *_scale(scale), decimation, count0,
                          _p(hv) if hv is not None else None, _p(xv), len(x),
                          _p(out.view(np.float64)))
    return out[:n].copy()


def firinterp_fast(coefs, L, x, hist=None):
    c, cc, cv = _coefs(coefs)
    x, xv = _cx(x)
    out = np.zeros(len(x) * L + 1, dtype=np.complex128)
    hv = None
    if hist is not None:
        h, hv = _cx(hist)
    n = lib().so_firinterp_fast(_p(cv), len(c), cc, L, _p(hv) if hv is not None else None,
                                _p(xv), len(x), _p(out.view(np.float64)))
    return out[:n].copy()


def sos_cascade_fast(ff, fb, x, state=None):
    """Returns (y, new_state); state is [nsec, 2] complex128 (v1, v2)."""
    ff = np.ascontiguousarray(ff, dtype=np.float64)
    fb = np.ascontiguousarray(fb, dtype=np.float64)
    nsec = len(ff) // 3
    st = np.zeros((nsec, 2), dtype=np.complex128) if state is None else \
        np.array(state, dtype=np.complex128).reshape(nsec, 2).copy()
    x, xv = _cx(x)
    out = np.zeros(max(len(x), 1), dtype=np.complex128)
    lib().so_sos_cascade_fast(_p(ff), _p(fb), nsec, _p(st.view(np.float64)), _p(xv), len(x),
                              _p(out.view(np.float64)))
    return out[:len(x)].copy(), st


# ----------------------------------------------------------------------------- design side
def msb_index(x):
    return lib().so_msb_index(x)


def interp_sub_len(n, L):
    return lib().so_interp_sub_len(n, L)


def kaiser_beta(a):
    return lib().so_kaiser_beta(a)


def window_kaiser(i, n, beta):
    return lib().so_window_kaiser(i, n, beta)


def firdes_kaiser(n, fc, a, mu=0.0):
    h = np.zeros(n)
    if lib().so_firdes_kaiser(n, fc, a, mu, _p(h)) != 0:
        raise ValueError("firdes_kaiser: argument rejected (firdes/mod.rs:284-290)")
    return h


def firdes_notch(m, f0, a):
    h = np.zeros(2 * m + 1)
    if lib().so_firdes_notch(m, f0, a, _p(h)) != 0:
        raise ValueError("firdes_notch: argument rejected (firdes/mod.rs:334-340)")
    return h


def filter_autocorrelation(h, lag):
    h = np.ascontiguousarray(h, dtype=np.float64)
    return lib().so_filter_autocorrelation(_p(h), len(h), lag)


def filter_crosscorrelation(h, g, lag):
    h = np.ascontiguousarray(h, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    return lib().so_filter_crosscorrelation(_p(h), len(h), _p(g), len(g), lag)


def filter_isi(h, samples_per_symbol, filter_delay):
    """firdes/mod.rs:553-573 -> (rms, max)"""
    h = np.ascontiguousarray(h, dtype=np.float64)
    out = np.zeros(2)
    lib().so_filter_isi(_p(h), len(h), samples_per_symbol, filter_delay, _p(out))
    return float(out[0]), float(out[1])


def filter_energy(h, cutoff_frequency, fft_size):
    """firdes/mod.rs:603-640 (DotProduct FORWARD on e^{j 2 pi f k})"""
    h = np.ascontiguousarray(h, dtype=np.float64)
    out = np.zeros(1)
    st = lib().so_filter_energy(_p(h), len(h), float(cutoff_frequency), fft_size, _p(out))
    if st:
        raise ValueError({-1: "Bandwidth", -2: "FilterSize", -3: "FFTSize"}[st])
    return float(out[0])


def pll_active_lag(w, zeta, k):
    num, den = np.zeros(3), np.zeros(3)
    if lib().so_pll_active_lag(w, zeta, k, _p(num), _p(den)) != 0:
        raise ValueError("active_lag: argument rejected (iirdes/pll/mod.rs:29-35)")
    return num, den


# ----------------------------------------------------------------------------- timed drivers
def run_units(kind, x, n_in, n_units, in_stride, out, out_stride, n_threads, *, coefs=None,
              scale=1.0, decimation=0, interpolation=0, ff=None, fb=None, preroll=0,
              x_offset=0, native=True):
    """Structural objects, one per unit (channel / stream segment), on n_threads pthreads.

    kind: "fir" | "decim" | "interp" | "iir".  x is a flat complex128 array; unit u reads
    x[x_offset + u*in_stride : ... + n_in] (and `preroll` samples before it to prime history).
    """
    L = lib(native)
    xv = x.view(np.float64)
    ov = out.view(np.float64)
    k = {"fir": 0, "decim": 0, "interp": 1, "iir": 2}[kind]
    if k == 2:
        ff = np.ascontiguousarray(ff, dtype=np.float64)
        fb = np.ascontiguousarray(fb, dtype=np.float64)
        cv, ncoef, cc = None, 0, 0
        ffp, fbp, nco = _p(ff), _p(fb), len(ff)
    else:
        c, cc, cvv = _coefs(coefs)
        cv, ncoef = _p(cvv), len(c)
        ffp = fbp = None
        nco = 0
    factor = decimation if kind == "decim" else interpolation
    base = C.cast(C.c_void_p(xv.ctypes.data + 16 * x_offset), c_dp)
    return L.so_run_units(k, cv, ncoef, cc, *_scale(scale), 1 if kind == "decim" else 0, factor,
                          ffp, fbp, nco, base, in_stride, n_in, preroll, _p(ov), out_stride,
                          n_units, n_threads)
