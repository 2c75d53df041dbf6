"""solid::filter::iir -- IIRFilter, sos::SecondOrderFilter, decim::DecimatingIIRFilter,
interp::InterpolatingIIRFilter with the reference's constructor arguments, accessors and error
variants (filter/iir/mod.rs, sos.rs, decim.rs, interp.rs)."""
from __future__ import annotations

import cmath
import ctypes as C
import math

import numpy as np

from .. import _ffi
from .._buffers import InBuf, OutBuf, dptr
from .._ffi import check, lib
from . import Filter
from .group_delay import DelayError, iir_group_delay


class IIRFilterType:
    """iir/mod.rs:62-66"""
    Normal = _ffi.IIR_NORMAL
    SecondOrder = _ffi.IIR_SECOND_ORDER


class IIRErrorCode:
    """iir/mod.rs:40-49"""
    _FROM_STATUS = {
        _ffi.ERR_IIR_NUMERATOR_LENGTH_ZERO: "NumeratorLengthZero",
        _ffi.ERR_IIR_DENOMINATOR_LENGTH_ZERO: "DenominatorLengthZero",
        _ffi.ERR_IIR_SOS_SIZE_ZERO: "SecondOrderSectionSizeZero",
        _ffi.ERR_IIR_SOS_SIZE_MISMATCH: "SecondOrderSectionSizeMismatch",
        _ffi.ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3: "SecondOrderSectionSizeNotMultpleOf3",
        _ffi.ERR_IIR_DECIMATION_LESS_THAN_ONE: "DecimationLessThanOne",
        _ffi.ERR_IIR_INTERPOLATION_LESS_THAN_ONE: "InterpolationLessThanOne",
    }


class IIRError(Exception):
    """IIRError(IIRErrorCode) -- iir/mod.rs:51-60"""

    def __init__(self, code: str):
        self.code = code
        super().__init__(f"IIR Filter Error {code}")


class SecondOrderError(Exception):
    """SecondOrderError(SecondOrderErrorCode) -- sos.rs:18-32"""

    def __init__(self, code: str = "CoefficientsNotInRange"):
        self.code = code
        super().__init__(f"Second Order Error {code}")


def _check_ctor(status: int) -> None:
    code = IIRErrorCode._FROM_STATUS.get(status)
    if code is not None:
        raise IIRError(code)
    if status == _ffi.ERR_SOS_COEFFICIENTS_NOT_IN_RANGE:
        raise SecondOrderError()
    check(status)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class IIRFilter(Filter):
    """IIRFilter<Coef, In> -- iir/mod.rs:68-419."""

    def __init__(self, feed_forward, feed_back, iirtype, n_channels: int = 1, _wrap=_ffi.IIR_PLAIN, _factor=0):
        self._h = C.c_void_p()
        ff, fb = _f64(feed_forward), _f64(feed_back)
        self._C = n_channels
        self._type = iirtype
        _check_ctor(lib.sgpu_iir_create(iirtype, dptr(ff), len(ff), dptr(fb), len(fb), n_channels, _wrap,
                                        max(_factor, 0), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            lib.sgpu_iir_destroy(h)
            h.value = None

    @property
    def n_channels(self) -> int:
        return self._C

    def _coefs(self, fn):
        n = _ffi.c_size()
        check(fn(self._h, None, C.byref(n)))
        out = np.zeros(n.value)
        check(fn(self._h, dptr(out), C.byref(n)))
        return out

    def numerator_coefs(self):  # iir/mod.rs:182
        return self._coefs(lib.sgpu_iir_numerator_coefs)

    def denominator_coefs(self):  # iir/mod.rs:202
        return self._coefs(lib.sgpu_iir_denominator_coefs)

    def second_order_filters(self):  # iir/mod.rs:222 -> one SecondOrderFilter view per section
        ff, fb = self.numerator_coefs(), self.denominator_coefs()
        if self._type != IIRFilterType.SecondOrder:
            return []
        return [SecondOrderFilter(ff[3 * i:3 * i + 3], fb[3 * i:3 * i + 3]) for i in range(len(ff) // 3)]

    def iir_type(self):  # iir/mod.rs:239
        return lib.sgpu_iir_type(self._h)

    def set_mode(self, mode: int):
        """-1 auto, 0 one-channel-per-thread batch, 1 chunked scan (fused when the filter decays), 2 three-pass scan"""
        check(lib.sgpu_iir_set_mode(self._h, mode))

    def decay_length(self) -> int:
        """samples after which an older state no longer matters (1e-10); 0 = does not decay.
        Segments of one stream on different GPUs warm up over that many preceding samples."""
        n = _ffi.c_size()
        check(lib.sgpu_iir_decay_length(self._h, C.byref(n)))
        return n.value

    def transition(self, n: int):
        """A^n [D, D] float64: evolution of the cascade's state (get_state order and scaling) over n samples of zero
        input.  state_after = A^n @ state_before + (end state of the same samples run from zero state)."""
        d = _ffi.c_size()
        check(lib.sgpu_iir_transition(self._h, n, None, C.byref(d)))
        A = np.zeros((d.value, d.value))
        check(lib.sgpu_iir_transition(self._h, n, A.ctypes.data_as(_ffi.c_dp), C.byref(d)))
        return A

    def execute_block(self, samples):  # iir/mod.rs:310
        ib = InBuf(samples, self._C)
        n_out = lib.sgpu_iir_out_len(self._h, ib.n)
        ob = OutBuf(ib, self._C, n_out)
        got = _ffi.c_size()
        check(lib.sgpu_iir_execute_block(self._h, ib.ptr, ib.n, ib.stride, ob.ptr, ob.stride, C.byref(got),
                                         ib.mem, ib.stream))
        assert got.value == n_out
        return ob.result(n_out)

    def get_state(self):
        n = lib.sgpu_iir_state_len(self._h)
        st = np.zeros((self._C, n), dtype=np.complex64)
        idx = C.c_uint64()
        check(lib.sgpu_iir_get_state(self._h, st.ctypes.data, C.byref(idx)))
        return st, idx.value

    def set_state(self, state, index: int = 0):
        st = np.ascontiguousarray(state, dtype=np.complex64).reshape(self._C, -1)
        assert st.shape[1] == lib.sgpu_iir_state_len(self._h)
        check(lib.sgpu_iir_set_state(self._h, st.ctypes.data, index))

    def reset(self):
        check(lib.sgpu_iir_reset(self._h))

    def clone(self):
        other = object.__new__(type(self))
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        other._h = C.c_void_p()
        check(lib.sgpu_iir_clone(self._h, C.byref(other._h)))
        return other

    def frequency_response(self, frequency: float) -> complex:
        """iir/mod.rs:336-373.  SecondOrder mode multiplies into a zero-initialised product, so the
        reference always returns 0 there (asserted by its doc-test, iir/mod.rs:328-334)."""
        if self._type == IIRFilterType.Normal:
            b = sum(c * cmath.rect(1.0, frequency * 2.0 * math.pi * i)
                    for i, c in enumerate(self.numerator_coefs()))
            a = sum(c * cmath.rect(1.0, frequency * 2.0 * math.pi * i)
                    for i, c in enumerate(self.denominator_coefs()))
            return b / a
        return 0j

    def group_delay(self, frequency: float) -> float:  # iir/mod.rs:374-395
        if self._type == IIRFilterType.SecondOrder:
            delay = 0.0
            for f in self.second_order_filters():
                delay = delay + f.group_delay(frequency) + 2.0
            return delay
        try:
            return iir_group_delay(self.numerator_coefs(), self.denominator_coefs(), frequency)
        except (DelayError, ZeroDivisionError):
            return 0.0


class DecimatingIIRFilter(IIRFilter):
    """iir/decim.rs:5-285: runs every sample, keeps outputs where (index+1) % M == 0."""

    def __init__(self, feed_forward, feed_back, iirtype, decimation: int, n_channels: int = 1):
        super().__init__(feed_forward, feed_back, iirtype, n_channels, _ffi.IIR_DECIMATING, decimation)
        self._M = decimation

    def get_decimation(self) -> int:  # decim.rs:64
        return self._M


class InterpolatingIIRFilter(IIRFilter):
    """iir/interp.rs:5-273: each input then L-1 zeros, all L outputs kept."""

    def __init__(self, feed_forward, feed_back, iirtype, interpolation: int, n_channels: int = 1):
        super().__init__(feed_forward, feed_back, iirtype, n_channels, _ffi.IIR_INTERPOLATING, interpolation)
        self._L = interpolation

    def get_interpolation(self) -> int:  # interp.rs:62
        return self._L


class SecondOrderFilter:
    """SecondOrderFilter<C, T> -- sos.rs:34-231: one biquad; a one-section IIR handle underneath."""

    def __init__(self, feed_forward, feed_back, n_channels: int = 1):
        ff, fb = _f64(feed_forward), _f64(feed_back)
        if len(ff) < 3 or len(fb) < 3:  # sos.rs:56-60
            raise SecondOrderError()
        self._ff, self._fb = ff[:3].copy(), fb[:3].copy()
        self._C = n_channels
        self._iir = None

    def _handle(self):
        if self._iir is None:
            self._iir = IIRFilter(self._ff, self._fb, IIRFilterType.SecondOrder, self._C)
        return self._iir

    def execute(self, sample):  # sos.rs:92 (Either<T, Out> collapses to one complex sample)
        out = self._handle().execute_block([sample] if self._C == 1 else [[s] for s in sample])
        return out[0] if self._C == 1 else out[:, 0]

    def numerator_coefs(self):  # sos.rs:116 -- holds a1, a2 (field names swapped in the reference)
        return (self._fb / self._fb[0])[1:]

    def denominator_coefs(self):  # sos.rs:136 -- holds b0, b1, b2
        return self._ff / self._fb[0]

    def frequency_response(self, frequency: float) -> complex:  # sos.rs:151-172
        b = sum(c * cmath.rect(1.0, frequency * 2.0 * math.pi * i) for i, c in enumerate(self.numerator_coefs()))
        a = sum(c * cmath.rect(1.0, frequency * 2.0 * math.pi * i) for i, c in enumerate(self.denominator_coefs()))
        return b / a

    def group_delay(self, frequency: float) -> float:  # sos.rs:208-230
        try:
            return iir_group_delay(list(self.numerator_coefs()), list(self.denominator_coefs()), frequency) + 2.0
        except (DelayError, ZeroDivisionError):
            return 0.0
