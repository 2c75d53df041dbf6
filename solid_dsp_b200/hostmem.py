"""Pinned host buffers on the NUMA node of a GPU (sgpu_host_alloc): the staging memory of the SGPU_HOST calls."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import check, lib


class PinnedArray:
    """A complex64 array [rows, cols] in pinned host memory local to `device`; `.array` is the numpy view."""

    def __init__(self, rows: int, cols: int, device: int = -1):
        self._p = C.c_void_p()
        self.nbytes = max(rows * cols, 1) * 8
        check(lib.sgpu_host_alloc(self.nbytes, device, C.byref(self._p)))
        buf = (C.c_char * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.complex64, count=rows * cols).reshape(rows, cols)

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p and self._p.value and lib is not None:
            self.array = None
            lib.sgpu_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        self.free()
