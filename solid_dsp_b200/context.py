"""Multi-GPU context: several GPUs of the box behind one filter object, host (numpy) buffers in and out
(sgpu_ctx_* / sgpu_sharded_*).  Channels are split into contiguous ranges; one FIR / decimating-FIR stream is
split into time segments with the T-1 sample halo sliced from the caller's buffer.  Results and streaming state are those of
the single-GPU types in solid_dsp_b200.filter."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._buffers import as_doubles, dptr
from ._ffi import check, lib
from .filter.fir import _check_ctor, _scale_parts


class ShardedFilter:
    def __init__(self, handle, n_channels: int):
        self._h = handle
        self._C = n_channels

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None:
            lib.sgpu_sharded_destroy(h)
            h.value = None

    @property
    def n_channels(self) -> int:
        return self._C

    @property
    def shards(self):
        """[(device, first_channel, n_channels), ...]"""
        out = []
        for i in range(lib.sgpu_sharded_shards(self._h)):
            d, f, n = C.c_int(), _ffi.c_size(), _ffi.c_size()
            check(lib.sgpu_sharded_shard_info(self._h, i, C.byref(d), C.byref(f), C.byref(n)))
            out.append((d.value, f.value, n.value))
        return out

    @property
    def last_segments(self) -> int:
        return lib.sgpu_sharded_last_segments(self._h)

    def out_len(self, n_in: int) -> int:
        return lib.sgpu_sharded_out_len(self._h, n_in)

    def reset(self):
        check(lib.sgpu_sharded_reset(self._h))

    def execute_block(self, samples, out=None):
        """Filter::execute_block (filter/mod.rs:14): numpy complex64 [n] or [C, n] in host memory.  `out`: an optional
        complex64 array [C, >= n_out] to receive the outputs (e.g. a PinnedArray's: the copies then run at the PCIe rate;
        a fresh pageable array, like the reference's `Vec`, costs the driver's staged copies)."""
        a = np.asarray(samples)
        squeeze = a.ndim <= 1
        a = np.ascontiguousarray(np.atleast_2d(a), dtype=np.complex64)
        if a.shape[0] != self._C:
            raise ValueError(f"expected {self._C} channels, got {a.shape[0]}")
        n = a.shape[1]
        n_out = self.out_len(n)
        if out is None:
            out = np.zeros((self._C, max(n_out, 1)), dtype=np.complex64)
        else:
            out = np.atleast_2d(out)
            if out.dtype != np.complex64 or out.shape[0] != self._C or out.shape[1] < max(n_out, 1) or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous complex64 array [channels, >= n_out]")
        got = _ffi.c_size()
        check(lib.sgpu_sharded_execute_block(self._h, a.ctypes.data, n, max(n, 1), out.ctypes.data, out.shape[1], C.byref(got)))
        assert got.value == n_out
        r = out[:, :n_out]
        return r[0] if squeeze else r


class Context:
    """sgpu_ctx: devices=None takes every visible GPU, an int the first n, a list names them (repeats allowed)."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None or isinstance(devices, int):
            check(lib.sgpu_ctx_create(int(devices or 0), C.byref(self._h)))
        else:
            arr = (C.c_int * len(devices))(*devices)
            check(lib.sgpu_ctx_create_devices(arr, len(devices), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None:
            lib.sgpu_ctx_destroy(h)
            h.value = None

    @property
    def n_devices(self) -> int:
        return lib.sgpu_ctx_devices(self._h)

    def fir(self, coefficents, scale=1.0, n_channels: int = 1, decimation: int = 0) -> ShardedFilter:
        """FIRFilter::new (fir/mod.rs:79) or, with decimation >= 1, DecimatingFIRFilter::new (decim.rs:27)."""
        cv, kind, n, _ = as_doubles(coefficents)
        h = C.c_void_p()
        _check_ctor(lib.sgpu_ctx_fir_create(self._h, dptr(cv), n, kind, n_channels, *_scale_parts(scale),
                                            1 if decimation else 0, decimation, C.byref(h)))
        return ShardedFilter(h, n_channels)

    def interpolator(self, coefficents, interpolation: int, n_channels: int = 1) -> ShardedFilter:
        """InterpolatingFIRFilter::new (interp.rs:27)."""
        cv, kind, n, _ = as_doubles(coefficents)
        h = C.c_void_p()
        _check_ctor(lib.sgpu_ctx_interp_create(self._h, dptr(cv), n, kind, n_channels, interpolation, C.byref(h)))
        return ShardedFilter(h, n_channels)

    def iir(self, ff, fb, iirtype, n_channels: int = 1, wrap: int = _ffi.IIR_PLAIN, factor: int = 1) -> ShardedFilter:
        """IIRFilter::new (iir/mod.rs:92) / Decimating- / InterpolatingIIRFilter::new."""
        ffv = np.ascontiguousarray(ff, dtype=np.float64)
        fbv = np.ascontiguousarray(fb, dtype=np.float64)
        h = C.c_void_p()
        check(lib.sgpu_ctx_iir_create(self._h, int(iirtype), dptr(ffv), ffv.size, dptr(fbv), fbv.size, n_channels, wrap, factor,
                                      C.byref(h)))
        return ShardedFilter(h, n_channels)
