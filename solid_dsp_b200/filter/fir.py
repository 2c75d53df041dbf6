"""solid::filter::fir -- FIRFilter, decim::DecimatingFIRFilter, interp::InterpolatingFIRFilter,
pfb::PolyPhaseFilterBank.  Same constructor arguments, method names and error variants as the
reference (filter/fir/mod.rs, decim.rs, interp.rs, pfb.rs); an extra `n_channels` keyword runs
that many identical filter objects, one per row of the input, in one launch."""
from __future__ import annotations

import cmath
import ctypes as C
import math

import numpy as np

from .. import _ffi
from .._buffers import InBuf, OutBuf, as_doubles, dptr
from .._ffi import check, lib
from . import Filter
from .group_delay import DelayError, fir_group_delay


class FIRErrorCode:
    """fir/mod.rs:39-45"""
    CoefficientsLengthZero = "CoefficientsLengthZero"
    DecimationLessThanOne = "DecimationLessThanOne"
    InterpolationLessThanOne = "InterpolationLessThanOne"
    NotEnoughFilters = "NotEnoughFilters"

    _FROM_STATUS = {
        _ffi.ERR_FIR_COEFFICIENTS_LENGTH_ZERO: "CoefficientsLengthZero",
        _ffi.ERR_FIR_DECIMATION_LESS_THAN_ONE: "DecimationLessThanOne",
        _ffi.ERR_FIR_INTERPOLATION_LESS_THAN_ONE: "InterpolationLessThanOne",
        _ffi.ERR_FIR_NOT_ENOUGH_FILTERS: "NotEnoughFilters",
    }


class FIRError(Exception):
    """FIRError(FIRErrorCode) -- fir/mod.rs:47-56; Display: "FIR Filter Error {code:?}"."""

    def __init__(self, code: str):
        self.code = code
        super().__init__(f"FIR Filter Error {code}")


def _check_ctor(status: int) -> None:
    code = FIRErrorCode._FROM_STATUS.get(status)
    if code is not None:
        raise FIRError(code)
    check(status)


def _scale_parts(scale):
    s = complex(scale)
    return s.real, s.imag


class _FirHandle(Filter):
    _destroy = "sgpu_fir_destroy"

    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None:  # `lib` is already None at interpreter shutdown
            getattr(lib, self._destroy)(h)
            h.value = None

    @property
    def n_channels(self) -> int:
        return self._C

    # Filter::execute_block through the C ABI
    def _run(self, fn, samples, n_out_of):
        ib = InBuf(samples, self._C)
        n_out = n_out_of(ib.n)
        ob = OutBuf(ib, self._C, n_out)
        got = _ffi.c_size()
        check(fn(self._h, ib.ptr, ib.n, ib.stride, ob.ptr, ob.stride, C.byref(got), ib.mem, ib.stream))
        assert got.value == n_out
        return ob.result(n_out)

    def frequency_response(self, frequency: float) -> complex:
        """fir/mod.rs:263-273 (and decim.rs / interp.rs twins): over the STORED coefficient order."""
        out = 0j
        for i, c in enumerate(self.coefficients()):
            out += c * cmath.rect(1.0, frequency * 2.0 * math.pi * i)
        return self.get_scale() * out

    def group_delay(self, frequency: float) -> float:
        """fir/mod.rs:293-303: 0.0 when fir_group_delay errors."""
        try:
            return fir_group_delay(self.coefficients(), frequency)
        except (DelayError, ZeroDivisionError):
            return 0.0


class FIRFilter(_FirHandle):
    """FIRFilter<Coef, In> -- fir/mod.rs:58-316.  y[n] = scale * sum_i h[T-1-i] x[n-i]."""

    def __init__(self, coefficents, scale=1.0, n_channels: int = 1):
        """coefficents: [T] taps shared by all channels, or [C, T]: one tap set per channel (C reference objects)."""
        super().__init__()
        cv, kind, n, self._coefs_in = as_doubles(coefficents)
        per_channel = self._coefs_in.ndim == 2
        if per_channel:
            n_channels = self._coefs_in.shape[0]
        self._C = n_channels
        self._complex = kind == _ffi.TAPS_COMPLEX
        create = lib.sgpu_fir_create_per_channel if per_channel else lib.sgpu_fir_create
        _check_ctor(create(dptr(cv), n, kind, n_channels, *_scale_parts(scale), 0, 0, C.byref(self._h)))

    def set_scale(self, scale):  # fir/mod.rs:106
        check(lib.sgpu_fir_set_scale(self._h, *_scale_parts(scale)))

    def get_scale(self):  # fir/mod.rs:124
        re, im = C.c_double(), C.c_double()
        check(lib.sgpu_fir_get_scale(self._h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value) if self._complex else re.value

    def len(self) -> int:  # fir/mod.rs:142
        return lib.sgpu_fir_len(self._h)

    def is_empty(self) -> bool:  # fir/mod.rs:158
        return self.len() == 0

    def coefficients(self, channel: int = 0):  # fir/mod.rs:176 -- stored (reversed) order
        n = self.len()
        out = np.zeros(n * (2 if self._complex else 1))
        check(lib.sgpu_fir_channel_coefficients(self._h, channel, dptr(out)))
        return out.view(np.complex128) if self._complex else out

    def execute_block(self, samples):  # fir/mod.rs:235
        return self._run(lib.sgpu_fir_execute_block, samples, lambda n: lib.sgpu_fir_out_len(self._h, n))

    @property
    def last_path(self) -> str:
        """'tensor' when the last execute_block ran on the tcgen05 kernel (long real-tap filters), else 'ffma'."""
        return "tensor" if lib.sgpu_fir_last_path(self._h) == 1 else "ffma"

    def write(self, samples):  # Window::write on the filter's history (window/mod.rs:73)
        ib = InBuf(samples, self._C)
        check(lib.sgpu_fir_write(self._h, ib.ptr, ib.n, ib.stride, ib.mem, ib.stream))

    def get_state(self):
        """(history [C, T-1] complex64 oldest first, current_item)"""
        T = self.len()
        hist = np.zeros((self._C, max(T - 1, 0)), dtype=np.complex64)
        cur = C.c_uint64()
        check(lib.sgpu_fir_get_state(self._h, hist.ctypes.data if hist.size else None, C.byref(cur)))
        return hist, cur.value

    def set_state(self, history, current_item: int = 0):
        hist = np.ascontiguousarray(history, dtype=np.complex64).reshape(self._C, -1)
        assert hist.shape[1] == self.len() - 1
        check(lib.sgpu_fir_set_state(self._h, hist.ctypes.data if hist.size else None, current_item))

    def reset(self):
        check(lib.sgpu_fir_reset(self._h))

    def clone(self):  # #[derive(Clone)]
        other = object.__new__(type(self))
        _FirHandle.__init__(other)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        other._h = C.c_void_p()
        check(lib.sgpu_fir_clone(self._h, C.byref(other._h)))
        return other

    def __str__(self):  # fir/mod.rs:306-316
        return f"FIR<f32> [Scale={self.get_scale():.5}] [Coefficients=DotProduct<f32> [Size={self.len()}]]"


class DecimatingFIRFilter(FIRFilter):
    """DecimatingFIRFilter<Coef, In> -- fir/decim.rs:5-295: emits when (count+1) % M == 0."""

    def __init__(self, coefficents, scale, decimation: int, n_channels: int = 1):
        _FirHandle.__init__(self)
        cv, kind, n, self._coefs_in = as_doubles(coefficents)
        per_channel = self._coefs_in.ndim == 2
        if per_channel:
            n_channels = self._coefs_in.shape[0]
        self._C = n_channels
        self._complex = kind == _ffi.TAPS_COMPLEX
        if n > 0 and decimation < 1:  # decim.rs:30 (usize cannot be negative; mirror the check)
            raise FIRError(FIRErrorCode.DecimationLessThanOne)
        create = lib.sgpu_fir_create_per_channel if per_channel else lib.sgpu_fir_create
        _check_ctor(create(dptr(cv), n, kind, n_channels, *_scale_parts(scale), 1, max(decimation, 0), C.byref(self._h)))

    def get_decimation(self) -> int:  # decim.rs:96
        return lib.sgpu_fir_decimation(self._h)

    def push(self, sample):  # decim.rs:115
        self.write([sample] if self._C == 1 else [[s] for s in sample])

    def __str__(self):  # decim.rs:281-295
        _, cur = self.get_state()
        return (f"FIR<f32> [Scale={self.get_scale():.5}] [Coefficients=DotProduct<f32> [Size={self.len()}]] "
                f"[Decimation={cur}/{self.get_decimation()}]")


class _InterpHandle(_FirHandle):
    _destroy = "sgpu_interp_destroy"

    def set_scale(self, scale):  # interp.rs:57 / pfb.rs:52 -- stored, never applied (pfb.rs:85-90)
        check(lib.sgpu_interp_set_scale(self._h, *_scale_parts(scale)))

    def get_scale(self):  # interp.rs:62 / pfb.rs:57
        re, im = C.c_double(), C.c_double()
        check(lib.sgpu_interp_get_scale(self._h, C.byref(re), C.byref(im)))
        return re.value

    def len(self) -> int:  # interp.rs:67 / pfb.rs:62: the number of sub-filters
        return lib.sgpu_interp_interpolation(self._h)

    def is_empty(self) -> bool:
        return self.len() == 0

    def sub_len(self) -> int:
        return lib.sgpu_interp_sub_len(self._h)

    def _phase_coefs(self):
        cx = getattr(self, "_complex", False)
        out = np.zeros(self.len() * self.sub_len() * (2 if cx else 1))
        check(lib.sgpu_interp_coefficients(self._h, dptr(out)))
        if cx:
            out = out.view(np.complex128)
        return out.reshape(self.len(), self.sub_len())

    def get_state(self):
        hist = np.zeros((self._C, max(self.sub_len() - 1, 0)), dtype=np.complex64)
        if hist.size:
            check(lib.sgpu_interp_get_state(self._h, hist.ctypes.data))
        return hist

    def set_state(self, history):
        hist = np.ascontiguousarray(history, dtype=np.complex64).reshape(self._C, -1)
        assert hist.shape[1] == self.sub_len() - 1
        if hist.size:
            check(lib.sgpu_interp_set_state(self._h, hist.ctypes.data))

    def reset(self):  # pfb.rs:76
        check(lib.sgpu_interp_reset(self._h))

    def clone(self):
        other = object.__new__(type(self))
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        other._h = C.c_void_p()
        check(lib.sgpu_interp_clone(self._h, C.byref(other._h)))
        return other


class PolyPhaseFilterBank(_InterpHandle):
    """PolyPhaseFilterBank<Coef, In> -- fir/pfb.rs:3-90."""

    def __init__(self, coefficients, filters: int, scale=1.0, n_channels: int = 1):
        _FirHandle.__init__(self)
        cv, kind, n, _ = as_doubles(coefficients)
        self._C = n_channels
        self._complex = kind == _ffi.TAPS_COMPLEX
        _check_ctor(lib.sgpu_pfb_create(dptr(cv), n, kind, n_channels, max(filters, 0),
                                        *_scale_parts(scale), C.byref(self._h)))

    def coefficents(self):  # pfb.rs:71 -> Vec<Vec<Coef>>
        return self._phase_coefs()

    def push(self, sample):  # pfb.rs:81
        ib = InBuf([sample] if self._C == 1 else [[s] for s in sample], self._C)
        check(lib.sgpu_interp_push(self._h, ib.ptr, ib.n, ib.stride, ib.mem, ib.stream))

    def execute(self, index: int):  # pfb.rs:85
        out = np.zeros(self._C, dtype=np.complex64)
        check(lib.sgpu_interp_execute_phase(self._h, index, out.ctypes.data, _ffi.HOST, None))
        return out[0] if self._C == 1 else out


class InterpolatingFIRFilter(_InterpHandle):
    """InterpolatingFIRFilter<Coef, In> -- fir/interp.rs:6-137: y[nL+p], p = 0..L-1, no scale."""

    def __init__(self, coefficents, interpolation: int, n_channels: int = 1):
        _FirHandle.__init__(self)
        cv, kind, n, orig = as_doubles(coefficents)
        per_channel = orig.ndim == 2
        if per_channel:
            n_channels = orig.shape[0]
        self._C = n_channels
        self._complex = kind == _ffi.TAPS_COMPLEX
        create = lib.sgpu_interp_create_per_channel if per_channel else lib.sgpu_interp_create
        _check_ctor(create(dptr(cv), n, kind, n_channels, max(interpolation, 0), C.byref(self._h)))

    def interpolation(self) -> int:  # interp.rs:82
        return lib.sgpu_interp_interpolation(self._h)

    def coefficents(self):  # interp.rs:77 -- flattened
        return self._phase_coefs().reshape(-1)

    def coefficients(self):
        return self.coefficents()

    def execute_block(self, samples):  # interp.rs:102
        L = self.interpolation()
        return self._run(lib.sgpu_interp_execute_block, samples, lambda n: n * L)

    @property
    def last_path(self) -> str:
        """'tensor' when the last execute_block ran on the tcgen05 kernel (real taps, L = 2 / 4, sub-filters of more than 32 taps)."""
        return "tensor" if lib.sgpu_interp_last_path(self._h) == 1 else "ffma"

    def __str__(self):
        return f"InterpolatingFIR<f32> [Interpolation={self.interpolation()}] [SubLen={self.sub_len()}]"
