"""Marshalling between numpy / torch arrays and the C ABI's (pointer, n, stride, mem) tuples."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class InBuf:
    """Input samples as cf32: numpy (HOST) or torch CUDA tensor (DEVICE), shape [n] or [C, n]."""

    def __init__(self, samples, n_channels: int):
        self.torch = _is_torch(samples)
        if self.torch:
            import torch
            t = samples
            if not t.is_cuda:
                t = t.cpu().numpy()
                self.torch = False
                samples = t
            else:
                if t.dtype != torch.complex64:
                    t = t.to(torch.complex64)
                if t.dim() == 1:
                    t = t.unsqueeze(0)
                if t.stride(-1) != 1:
                    t = t.contiguous()
                if t.shape[0] != n_channels:
                    raise ValueError(f"expected {n_channels} channels, got {t.shape[0]}")
                self.keep = t
                self.ptr = t.data_ptr()
                self.n = t.shape[1]
                self.stride = t.stride(0) if t.shape[0] > 1 else max(self.n, 1)
                self.mem = _ffi.DEVICE
                self.device = t.device
                self.stream = torch.cuda.current_stream(t.device).cuda_stream
                self.squeeze = samples.dim() == 1
                return
        a = np.asarray(samples)
        self.squeeze = a.ndim <= 1
        a = np.atleast_1d(a)
        if a.ndim == 1:
            a = a[None, :]
        a = np.ascontiguousarray(a, dtype=np.complex64)
        if a.shape[0] != n_channels:
            raise ValueError(f"expected {n_channels} channels, got {a.shape[0]}")
        self.keep = a
        self.ptr = a.ctypes.data
        self.n = a.shape[1]
        self.stride = max(self.n, 1)
        self.mem = _ffi.HOST
        self.stream = None
        self.device = None


class OutBuf:
    def __init__(self, like: InBuf, n_channels: int, n_out: int):
        self.n_out = n_out
        cap = max(n_out, 1)
        if like.torch:
            import torch
            self.arr = torch.empty((n_channels, cap), dtype=torch.complex64, device=like.device)
            self.ptr = self.arr.data_ptr()
        else:
            self.arr = np.zeros((n_channels, cap), dtype=np.complex64)
            self.ptr = self.arr.ctypes.data
        self.stride = cap
        self.squeeze = like.squeeze

    def result(self, n_out: int):
        r = self.arr[:, :n_out]
        return r[0] if self.squeeze else r


def as_doubles(coefs):
    """-> (contiguous float64 view for the ABI, tap kind, n_taps, original array).  A 2-D array [C, T] means one tap
    set per channel (n_taps = T)."""
    a = np.asarray(coefs)
    if np.iscomplexobj(a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        return a.view(np.float64), _ffi.TAPS_COMPLEX, a.shape[-1], a
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, _ffi.TAPS_REAL, a.shape[-1], a


def dptr(a):
    if a is None or a.size == 0:
        return None
    return a.ctypes.data_as(_ffi.c_dp)
