"""solid::nco::NCO (nco/mod.rs) -- numerically controlled oscillator: a 32-bit phase accumulator, a 32-bit phase step and
a 1024-entry sine table.  `n_channels` objects advance in lock-step on the GPU (one per row of the input); the scalar
accessors default to channel 0 and the setters to every channel."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _ffi
from ._buffers import InBuf, OutBuf
from ._ffi import check, lib


def constrain(theta: float) -> int:
    """nco/mod.rs:176-188: the fraction of a turn as a u32."""
    return int(lib.sgpu_nco_constrain(float(theta)))


class NCO:
    _TABLE = [math.sin(2.0 * math.pi * i / 1024.0) for i in range(1024)]  # nco/mod.rs:36-41 (f64)

    def __init__(self, n_channels: int = 1, _handle=None, _owner=None):
        self._h = C.c_void_p()
        self._owner = _owner  # a DigitalDownConverter owns the handle it hands out
        if _handle is not None:
            self._h.value = _handle
        else:
            check(lib.sgpu_nco_create(n_channels, C.byref(self._h)))
        self._C = int(lib.sgpu_nco_channels(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and lib is not None and self._owner is None:
            lib.sgpu_nco_destroy(h)
            h.value = None

    @property
    def n_channels(self) -> int:
        return self._C

    def reset(self):  # nco/mod.rs:53
        check(lib.sgpu_nco_reset(self._h))

    def set_frequency(self, delta_theta: float, channel: int = _ffi.ALL_CHANNELS):  # nco/mod.rs:59
        check(lib.sgpu_nco_set_frequency(self._h, channel, float(delta_theta)))

    def adjust_frequency(self, dt: float, channel: int = _ffi.ALL_CHANNELS):  # nco/mod.rs:64
        check(lib.sgpu_nco_adjust_frequency(self._h, channel, float(dt)))

    def set_phase(self, phi: float, channel: int = _ffi.ALL_CHANNELS):  # nco/mod.rs:79
        check(lib.sgpu_nco_set_phase(self._h, channel, float(phi)))

    def adjust_phase(self, delta_phi: float, channel: int = _ffi.ALL_CHANNELS):  # nco/mod.rs:84
        check(lib.sgpu_nco_adjust_phase(self._h, channel, float(delta_phi)))

    def raw(self, channel: int = 0):
        """(theta, delta_theta) as the u32 words of the accumulator."""
        t, d = C.c_uint32(), C.c_uint32()
        check(lib.sgpu_nco_get(self._h, channel, C.byref(t), C.byref(d)))
        return t.value, d.value

    def set_raw(self, theta: int, delta_theta: int, channel: int = _ffi.ALL_CHANNELS):
        check(lib.sgpu_nco_set(self._h, channel, theta & 0xFFFFFFFF, delta_theta & 0xFFFFFFFF))

    def get_frequency(self, channel: int = 0) -> float:
        """nco/mod.rs:69-76: the integer division by 2^32 makes this 0.0 for every value (kept)."""
        dt = float(self.raw(channel)[1] // (1 << 32)) * 2.0 * math.pi
        return dt - 2.0 * math.pi if dt > math.pi else dt

    def get_phase(self, channel: int = 0) -> float:  # nco/mod.rs:89-91, same quirk
        return float(self.raw(channel)[0] // (1 << 32)) * 2.0 * math.pi

    def step(self, count: int = 1):  # nco/mod.rs:93
        check(lib.sgpu_nco_step(self._h, count))

    def _index(self, channel: int = 0) -> int:  # nco/mod.rs:98-101
        return (((self.raw(channel)[0] + (1 << 21)) & 0xFFFFFFFF) >> 22) & 0x3FF

    def sin(self, channel: int = 0) -> float:  # nco/mod.rs:103
        return self._TABLE[self._index(channel)]

    def cos(self, channel: int = 0) -> float:  # nco/mod.rs:108
        return self._TABLE[(self._index(channel) + 256) & 0x3FF]

    def sincos(self, channel: int = 0):  # nco/mod.rs:114
        return self.sin(channel), self.cos(channel)

    def complex_exponential(self, channel: int = 0) -> complex:  # nco/mod.rs:119
        return complex(self.cos(channel), self.sin(channel))

    def mix_up(self, sample, channel: int = 0) -> complex:  # nco/mod.rs:141 (host arithmetic, one sample)
        return self.complex_exponential(channel) * complex(sample)

    def mix_down(self, sample, channel: int = 0) -> complex:  # nco/mod.rs:147
        return self.complex_exponential(channel).conjugate() * complex(sample)

    def _mix_block(self, up: bool, samples):
        ib = InBuf(samples, self._C)
        ob = OutBuf(ib, self._C, ib.n)
        check(lib.sgpu_nco_mix_block(self._h, 1 if up else 0, ib.ptr, ib.n, ib.stride, ob.ptr, ob.stride, ib.mem, ib.stream))
        return ob.result(ib.n)

    def mix_up_block(self, samples):
        """nco/mod.rs:153-161 as it was meant (the reference indexes an empty Vec): y[i] = mix_up(x[i]); step()."""
        return self._mix_block(True, samples)

    def mix_down_block(self, samples):  # nco/mod.rs:164-172
        return self._mix_block(False, samples)

    def clone(self):
        h = C.c_void_p()
        check(lib.sgpu_nco_clone(self._h, C.byref(h)))
        return NCO(_handle=h.value)

    def __str__(self):  # nco/mod.rs:196-203
        t, d = self.raw(0)
        return f"NCO [Theta={t}] [ΔTheta={d}]"


__all__ = ["NCO", "constrain", "np"]
