"""solid::filter::auto_correlator::AutoCorrelator (filter/auto_correlator/mod.rs:24-216) on the GPU.

    r[n] = sum_{i < W-d} x[n-i] * conj(x[n-d-i])     (W = window_size, d = delay; 0 for d >= W --
                                                      the reference's Window(capacity, delay) never
                                                      writes the delayed tail, window/mod.rs:17-71)
    get_energy() = sum_{i < W} |x[n-i]|^2

One object = `n_channels` independent correlators (one reference object per channel)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _ffi
from .._buffers import InBuf, OutBuf, dptr
from .._ffi import check, lib


class AutoCorrelator:
    def __init__(self, window_size: int, delay: int, n_channels: int = 1):  # auto_correlator/mod.rs:51
        self._h = C.c_void_p()
        self._C = n_channels
        check(lib.sgpu_autocorr_create(max(window_size, 0), max(delay, 0), n_channels, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None:
            lib.sgpu_autocorr_destroy(h)
            h.value = None

    @property
    def n_channels(self) -> int:
        return self._C

    def window_size(self) -> int:
        return lib.sgpu_autocorr_window_size(self._h)

    def delay(self) -> int:
        return lib.sgpu_autocorr_delay(self._h)

    def reset(self):  # auto_correlator/mod.rs:76
        check(lib.sgpu_autocorr_reset(self._h))

    def write(self, samples):  # auto_correlator/mod.rs:130: push every sample, no output
        ib = InBuf(samples, self._C)
        if ib.n:
            check(lib.sgpu_autocorr_write(self._h, ib.ptr, ib.n, ib.stride, ib.mem, ib.stream))

    def push(self, sample):  # auto_correlator/mod.rs:99
        self.write([sample] if self._C == 1 else [[s] for s in sample])

    def execute(self):  # auto_correlator/mod.rs:165: output of the current window
        out = np.zeros(self._C, dtype=np.complex128)
        check(lib.sgpu_autocorr_execute(self._h, dptr(out.view(np.float64))))
        return out[0] if self._C == 1 else out

    def execute_block(self, samples):  # auto_correlator/mod.rs:184: one output per input
        ib = InBuf(samples, self._C)
        ob = OutBuf(ib, self._C, ib.n)
        got = _ffi.c_size()
        check(lib.sgpu_autocorr_execute_block(self._h, ib.ptr, ib.n, ib.stride, ob.ptr, ob.stride, C.byref(got),
                                              ib.mem, ib.stream))
        assert got.value == ib.n
        return ob.result(ib.n)

    def get_energy(self):  # auto_correlator/mod.rs:214
        out = np.zeros(self._C)
        check(lib.sgpu_autocorr_get_energy(self._h, dptr(out)))
        return float(out[0]) if self._C == 1 else out

    def get_state(self):
        st = np.zeros((self._C, self.window_size()), dtype=np.complex64)
        check(lib.sgpu_autocorr_get_state(self._h, st.ctypes.data))
        return st

    def set_state(self, state):
        st = np.ascontiguousarray(state, dtype=np.complex64).reshape(self._C, self.window_size())
        check(lib.sgpu_autocorr_set_state(self._h, st.ctypes.data))

    def clone(self):
        other = object.__new__(type(self))
        other._C = self._C
        other._h = C.c_void_p()
        check(lib.sgpu_autocorr_clone(self._h, C.byref(other._h)))
        return other

    def __str__(self):  # auto_correlator/mod.rs:219-228
        e = self.get_energy()
        return f"AutoCorrelator<f64> [Size={self.window_size()}] [Delay={self.delay()}] [Energy={e}]"
