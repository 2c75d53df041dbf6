"""Digital down-converter: NCO::mix_down + step per sample (nco/mod.rs:93-96,147-151) in front of a
DecimatingFIRFilter (filter/fir/decim.rs) -- the chain a user of the reference writes as

    for x in samples: out += decim.execute(nco.mix_down(x)); nco.step()

run as one kernel per call on the shapes the decimator's warp kernel serves (M in {2, 4, 8}, real taps): the mixed
stream never reaches HBM.  `n_channels` independent (NCO, decimator) pairs, one per row of the input."""
from __future__ import annotations

import ctypes as C

from .. import _ffi
from .._buffers import InBuf, as_doubles, dptr
from .._ffi import check, lib
from ..nco import NCO
from .fir import DecimatingFIRFilter, FIRError, FIRErrorCode, _check_ctor, _FirHandle, _scale_parts


class DigitalDownConverter(_FirHandle):
    _destroy = "sgpu_ddc_destroy"

    def __init__(self, coefficents, scale, decimation: int, frequency: float = 0.0, n_channels: int = 1):
        super().__init__()
        cv, kind, n, _ = as_doubles(coefficents)
        self._C = n_channels
        self._complex = kind == _ffi.TAPS_COMPLEX
        if n > 0 and decimation < 1:  # decim.rs:30
            raise FIRError(FIRErrorCode.DecimationLessThanOne)
        _check_ctor(lib.sgpu_ddc_create(dptr(cv), n, kind, n_channels, *_scale_parts(scale), max(decimation, 0),
                                        C.byref(self._h)))
        # views on the two halves; the DDC handle owns them
        self.nco = NCO(_handle=lib.sgpu_ddc_nco(self._h), _owner=self)
        self._fir_h = C.c_void_p(lib.sgpu_ddc_filter(self._h))
        if frequency:
            self.nco.set_frequency(frequency)

    @property
    def filter(self) -> DecimatingFIRFilter:
        """The decimator as a DecimatingFIRFilter view (scale, taps, state); valid while this object lives."""
        v = object.__new__(_BorrowedDecimator)
        v.__dict__.update(_h=self._fir_h, _C=self._C, _complex=self._complex, _owner=self)
        return v

    def get_decimation(self) -> int:
        return lib.sgpu_fir_decimation(self._fir_h)

    def len(self) -> int:
        return lib.sgpu_fir_len(self._fir_h)

    def execute_block(self, samples):
        return self._run(lib.sgpu_ddc_execute_block, samples, lambda n: lib.sgpu_ddc_out_len(self._h, n))

    def write(self, samples):
        ib = InBuf(samples, self._C)
        check(lib.sgpu_ddc_write(self._h, ib.ptr, ib.n, ib.stride, ib.mem, ib.stream))

    def reset(self):
        check(lib.sgpu_ddc_reset(self._h))

    @property
    def last_fused(self) -> bool:
        """True when the last call mixed inside the decimator kernel (no mixed stream in HBM)."""
        return lib.sgpu_ddc_last_fused(self._h) == 1

    def clone(self):
        other = object.__new__(type(self))
        _FirHandle.__init__(other)
        other._C, other._complex = self._C, self._complex
        check(lib.sgpu_ddc_clone(self._h, C.byref(other._h)))
        other.nco = NCO(_handle=lib.sgpu_ddc_nco(other._h), _owner=other)
        other._fir_h = C.c_void_p(lib.sgpu_ddc_filter(other._h))
        return other


class _BorrowedDecimator(DecimatingFIRFilter):
    """A DecimatingFIRFilter whose handle belongs to a DigitalDownConverter."""

    def __del__(self):
        pass
