"""solid::dot_product -- DotProduct<T>, Direction, trait Execute (dot_product/mod.rs, execute.rs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._buffers import InBuf, as_doubles, dptr
from ._ffi import check, lib


class Direction:
    """dot_product/mod.rs:31-34"""
    FORWARD = _ffi.FORWARD
    REVERSE = _ffi.REVERSE


class DotProduct:
    """DotProduct<T> -- dot_product/mod.rs:37-196; execute() is trait Execute (execute.rs:1-18)."""

    def __init__(self, coefficients, direction):
        self._h = C.c_void_p()
        cv, kind, n, _ = as_doubles(coefficients)
        self._complex = kind == _ffi.TAPS_COMPLEX
        check(lib.sgpu_dot_create(dptr(cv), n, kind, direction, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            lib.sgpu_dot_destroy(h)
            h.value = None

    def len(self) -> int:  # dot_product/mod.rs:124
        return lib.sgpu_dot_len(self._h)

    def is_empty(self) -> bool:  # dot_product/mod.rs:141
        return self.len() == 0

    def coefficents(self):  # [sic] dot_product/mod.rs:102 -- the STORED order
        out = np.zeros(self.len() * (2 if self._complex else 1))
        if out.size:
            check(lib.sgpu_dot_coefficients(self._h, dptr(out)))
        return out.view(np.complex128) if self._complex else out

    def execute(self, samples):
        """Execute::execute: sum over min(len_c, len_x) terms.  [n] -> scalar, [V, n] -> [V]."""
        a = np.asarray(samples) if not type(samples).__module__.startswith("torch") else samples
        n_vec = 1 if a.ndim <= 1 else a.shape[0]
        ib = InBuf(samples, n_vec)
        if ib.torch:
            import torch
            out = torch.empty(n_vec, dtype=torch.complex64, device=ib.device)
            optr = out.data_ptr()
        else:
            out = np.zeros(n_vec, dtype=np.complex64)
            optr = out.ctypes.data
        check(lib.sgpu_dot_execute(self._h, ib.ptr, ib.n, ib.stride, n_vec, optr, ib.mem, ib.stream))
        return out[0] if ib.squeeze else out

    def __str__(self):  # dot_product/mod.rs:146-151
        return f"DotProduct<{'Complex<f32>' if self._complex else 'f32'}> [Size={self.len()}]"
