"""solid::window::Window<T> (window/mod.rs:9-126) -- host-side history type.

Fixed-capacity shift register, newest element at index 0, zero initialised.  On the GPU path the
same state lives in the filter handle as a history tail (sgpu_fir_get_state / _set_state); this
host type keeps the reference's API for callers that build their own pipelines, and converts to
and from the handle's layout (oldest first) with `to_history` / `from_history`."""
from __future__ import annotations

import numpy as np


class Window:
    def __init__(self, capacity: int, delay: int = 0, dtype=np.complex64):
        assert capacity > 0  # window/mod.rs:18
        self._capacity = capacity
        self._delay = delay
        self._buf = np.zeros(capacity + delay, dtype=dtype)

    def capacity(self) -> int:  # window/mod.rs:59
        return self._capacity

    def to_vec(self):  # window/mod.rs:44-51: `capacity` elements starting at `delay`
        return self._buf[self._delay:self._delay + self._capacity].copy()

    def as_ptr(self):  # window/mod.rs:36-42 returns a fresh copy (the reference leaks it)
        return self.to_vec()

    def push(self, element) -> None:  # window/mod.rs:63-71: moves capacity-1 elements, writes index 0
        self._buf[1:self._capacity] = self._buf[0:self._capacity - 1].copy()
        self._buf[0] = element

    def write(self, other) -> None:  # window/mod.rs:73-77
        for e in other:
            self.push(e)

    def reset(self) -> None:  # window/mod.rs:54-56 (without the reference's leak)
        self._buf[:] = 0

    def clone(self) -> "Window":  # window/mod.rs:103-120
        w = Window(self._capacity, self._delay, self._buf.dtype)
        w._buf[:self._capacity] = self._buf[:self._capacity]
        return w

    # ---- bridge to the GPU handles' state layout
    def to_history(self, n: int):
        """The n most recent samples, oldest first -- what sgpu_*_set_state expects."""
        return self._buf[:n][::-1].copy()

    @classmethod
    def from_history(cls, history, capacity: int | None = None) -> "Window":
        history = np.asarray(history)
        w = cls(capacity or max(len(history), 1), 0, history.dtype)
        w.write(history)
        return w

    def __str__(self):  # window/mod.rs:90-100
        return f"Window<{self._buf.dtype}> [Capacity={self._capacity}] [Delay={self._delay}]"
