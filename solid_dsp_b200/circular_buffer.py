"""solid::circular_buffer::CircularBuffer<T> (circular_buffer/mod.rs:55-627) -- host-side ring FIFO.

No filter in the reference uses this type; it is kept as API with the
same method names, index behaviour and error codes."""
from __future__ import annotations

import numpy as np


class BufferErrorCode:
    """circular_buffer/mod.rs:27-33"""
    EmptyBuffer = "EmptyBuffer"
    FullBuffer = "FullBuffer"
    NotEnoughBuffer = "NotEnoughBuffer"
    NegativeBuffer = "NegativeBuffer"
    NonExistantBuffer = "NonExistantBuffer"


class BufferError(Exception):
    """BufferError(BufferErrorCode) -- circular_buffer/mod.rs:36-48"""

    def __init__(self, code: str):
        self.code = code
        super().__init__(f"Buffer Error {code}")


class CircularBuffer:
    def __init__(self, capacity: int, dtype=np.complex64):
        assert capacity > 0  # :80
        self._capacity = int(capacity)
        self._buf = np.zeros(self._capacity, dtype=dtype)
        self._read = 0
        self._write = 0
        self._n = 0

    @classmethod
    def from_vec(cls, vec, dtype=None):  # :114
        vec = np.asarray(vec)
        cb = cls(len(vec), dtype or vec.dtype)
        cb.append(vec)
        return cb

    from_slice = from_vec  # :136

    def as_ptr(self):  # :164 -- the raw storage
        return self._buf

    def as_mut_ptr(self):  # :191 -- linearises first
        self.linearize()
        return self._buf

    def linearize(self) -> None:  # :220-238
        self._buf = np.concatenate([self._buf[self._read:], self._buf[:self._read]])
        # Rust's % keeps the dividend's sign, so the reference can go negative here (:235)
        d = self._write - self._read
        self._write = int(np.fmod(d, self._capacity))
        self._read = 0

    def to_vec(self):  # :261-269: all `capacity` slots starting at read_index
        return np.concatenate([self._buf[self._read:], self._buf[:self._read]])

    def reset(self) -> None:  # :289
        self._read = self._write = self._n = 0

    def len(self) -> int:  # :313
        return self._n

    def __len__(self):
        return self._n

    def capacity(self) -> int:  # :326
        return self._capacity

    def reserved(self) -> int:  # :343
        return self._capacity - self._n

    def is_empty(self) -> bool:  # :357
        return self._n == 0

    def is_full(self) -> bool:  # :375
        return self._n == self._capacity

    def read_index(self) -> int:  # :395
        return self._read

    def write_index(self) -> int:  # :414
        return self._write

    def push(self, element) -> None:  # :433-447
        if self.is_full():
            raise BufferError(BufferErrorCode.FullBuffer)
        self._buf[self._write % self._capacity] = element
        self._write = (self._write + 1) % self._capacity
        self._n += 1

    def append(self, other) -> None:  # :469-494
        other = np.asarray(other)
        k = len(other)
        if self._n + k > self._capacity:
            raise BufferError(BufferErrorCode.NotEnoughBuffer)
        room = self._capacity - self._write
        if k <= room:
            self._buf[self._write:self._write + k] = other
        else:
            self._buf[self._write:] = other[:room]
            # reference quirk kept (:486-490): the wrapped part is copied from offset k-room, not
            # from offset room (identical only when k == 2*room); reads past the slice are clipped
            src = other[k - room:k - room + (k - room)]
            self._buf[:len(src)] = src
        self._write = (self._write + k) % self._capacity
        self._n += k

    def pop(self):  # :512-524
        if self.is_empty():
            raise BufferError(BufferErrorCode.EmptyBuffer)
        v = self._buf[self._read]
        self._read = (self._read + 1) % self._capacity
        self._n -= 1
        return v

    def release(self, n: int) -> None:  # :548-557
        if n < 0:
            raise BufferError(BufferErrorCode.NegativeBuffer)
        if n > self._n:
            raise BufferError(BufferErrorCode.NotEnoughBuffer)
        self._read = (self._read + n) % self._capacity
        self._n -= n

    def deref(self):  # Deref<[T]> :603-610: the first len() storage slots
        return self._buf[:self._n]

    def clone(self):  # :577-601
        cb = CircularBuffer(self._capacity, self._buf.dtype)
        cb._buf[:] = self._buf
        cb._read, cb._write, cb._n = self._read, self._write, self._n
        return cb

    def __str__(self):  # :619-627
        vals = ", ".join(str(self._buf[(self._read + i) % self._capacity]) for i in range(self._n))
        return f"CircularBuffer<{self._buf.dtype}> [{vals}]"
