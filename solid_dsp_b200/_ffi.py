"""ctypes binding of libsolid_gpu.so -- one Python declaration per prototype in include/solid_gpu.h.

There is no fallback: if the library is missing or fails to load, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
ROOT = _PKG.parent
LIB_PATH = _PKG / "lib" / "libsolid_gpu.so"
PEAK_LIB_PATH = _PKG / "lib" / "libsgpu_peakbench.so"
HEADER = ROOT / "include" / "solid_gpu.h"

c_size = C.c_size_t
c_fp = C.POINTER(C.c_float)
c_dp = C.POINTER(C.c_double)
c_u64p = C.POINTER(C.c_uint64)
c_sizep = C.POINTER(c_size)
vp = C.c_void_p
vpp = C.POINTER(vp)

# status codes (include/solid_gpu.h: sgpu_status)
OK = 0
ERR_FIR_COEFFICIENTS_LENGTH_ZERO = -1
ERR_FIR_DECIMATION_LESS_THAN_ONE = -2
ERR_FIR_INTERPOLATION_LESS_THAN_ONE = -3
ERR_FIR_NOT_ENOUGH_FILTERS = -4
ERR_IIR_NUMERATOR_LENGTH_ZERO = -10
ERR_IIR_DENOMINATOR_LENGTH_ZERO = -11
ERR_IIR_SOS_SIZE_ZERO = -12
ERR_IIR_SOS_SIZE_MISMATCH = -13
ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3 = -14
ERR_IIR_DECIMATION_LESS_THAN_ONE = -15
ERR_IIR_INTERPOLATION_LESS_THAN_ONE = -16
ERR_SOS_COEFFICIENTS_NOT_IN_RANGE = -17
ERR_FIRDES_BANDWIDTH = -20
ERR_FIRDES_STOP_BAND_LEVEL = -21
ERR_FIRDES_MU = -22
ERR_INVALID_ARGUMENT = -30
ERR_CAPACITY = -31
ERR_CUDA = -32
ERR_UNSUPPORTED = -33
ERR_NO_DEVICE = -34
ERR_ALLOC = -35

HOST, DEVICE = 0, 1
TAPS_REAL, TAPS_COMPLEX = 0, 1
FORWARD, REVERSE = 0, 1
IIR_NORMAL, IIR_SECOND_ORDER = 0, 1
IIR_PLAIN, IIR_DECIMATING, IIR_INTERPOLATING = 0, 1, 2
ALL_CHANNELS = (1 << (8 * C.sizeof(C.c_size_t))) - 1

# name -> (restype, argtypes); must cover every prototype in the header (tests check this)
PROTOTYPES = {
    "sgpu_abi_version": (C.c_int, []),
    "sgpu_last_error": (C.c_char_p, []),
    "sgpu_status_name": (C.c_char_p, [C.c_int]),
    "sgpu_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4 + [c_sizep]),
    "sgpu_launch_count": (C.c_uint64, []),
    "sgpu_host_alloc": (C.c_int, [c_size, C.c_int, vpp]),
    "sgpu_host_free": (C.c_int, [vp]),
    "sgpu_fir_create": (C.c_int, [c_dp, c_size, C.c_int, c_size, C.c_double, C.c_double, C.c_int, c_size, vpp]),
    "sgpu_fir_create_per_channel": (C.c_int, [c_dp, c_size, C.c_int, c_size, C.c_double, C.c_double, C.c_int, c_size, vpp]),
    "sgpu_fir_taps_per_channel": (C.c_int, [vp]),
    "sgpu_fir_channel_coefficients": (C.c_int, [vp, c_size, c_dp]),
    "sgpu_fir_destroy": (C.c_int, [vp]),
    "sgpu_fir_clone": (C.c_int, [vp, vpp]),
    "sgpu_fir_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep, C.c_int, vp]),
    "sgpu_fir_write": (C.c_int, [vp, vp, c_size, c_size, C.c_int, vp]),
    "sgpu_fir_out_len": (c_size, [vp, c_size]),
    "sgpu_fir_set_scale": (C.c_int, [vp, C.c_double, C.c_double]),
    "sgpu_fir_get_scale": (C.c_int, [vp, c_dp, c_dp]),
    "sgpu_fir_len": (c_size, [vp]),
    "sgpu_fir_decimation": (c_size, [vp]),
    "sgpu_fir_channels": (c_size, [vp]),
    "sgpu_fir_last_path": (C.c_int, [vp]),
    "sgpu_fir_coefficients": (C.c_int, [vp, c_dp]),
    "sgpu_fir_get_state": (C.c_int, [vp, vp, c_u64p]),
    "sgpu_fir_set_state": (C.c_int, [vp, vp, C.c_uint64]),
    "sgpu_fir_reset": (C.c_int, [vp]),
    "sgpu_interp_create": (C.c_int, [c_dp, c_size, C.c_int, c_size, c_size, vpp]),
    "sgpu_pfb_create": (C.c_int, [c_dp, c_size, C.c_int, c_size, c_size, C.c_double, C.c_double, vpp]),
    "sgpu_interp_create_per_channel": (C.c_int, [c_dp, c_size, C.c_int, c_size, c_size, vpp]),
    "sgpu_interp_destroy": (C.c_int, [vp]),
    "sgpu_interp_clone": (C.c_int, [vp, vpp]),
    "sgpu_interp_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep, C.c_int, vp]),
    "sgpu_interp_push": (C.c_int, [vp, vp, c_size, c_size, C.c_int, vp]),
    "sgpu_interp_execute_phase": (C.c_int, [vp, c_size, vp, C.c_int, vp]),
    "sgpu_interp_set_scale": (C.c_int, [vp, C.c_double, C.c_double]),
    "sgpu_interp_get_scale": (C.c_int, [vp, c_dp, c_dp]),
    "sgpu_interp_interpolation": (c_size, [vp]),
    "sgpu_interp_sub_len": (c_size, [vp]),
    "sgpu_interp_channels": (c_size, [vp]),
    "sgpu_interp_last_path": (C.c_int, [vp]),
    "sgpu_interp_coefficients": (C.c_int, [vp, c_dp]),
    "sgpu_interp_get_state": (C.c_int, [vp, vp]),
    "sgpu_interp_set_state": (C.c_int, [vp, vp]),
    "sgpu_interp_reset": (C.c_int, [vp]),
    "sgpu_iir_create": (C.c_int, [C.c_int, c_dp, c_size, c_dp, c_size, c_size, C.c_int, c_size, vpp]),
    "sgpu_iir_destroy": (C.c_int, [vp]),
    "sgpu_iir_clone": (C.c_int, [vp, vpp]),
    "sgpu_iir_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep, C.c_int, vp]),
    "sgpu_iir_out_len": (c_size, [vp, c_size]),
    "sgpu_iir_sections": (c_size, [vp]),
    "sgpu_iir_channels": (c_size, [vp]),
    "sgpu_iir_type": (C.c_int, [vp]),
    "sgpu_iir_numerator_coefs": (C.c_int, [vp, c_dp, c_sizep]),
    "sgpu_iir_denominator_coefs": (C.c_int, [vp, c_dp, c_sizep]),
    "sgpu_iir_get_state": (C.c_int, [vp, vp, c_u64p]),
    "sgpu_iir_set_state": (C.c_int, [vp, vp, C.c_uint64]),
    "sgpu_iir_reset": (C.c_int, [vp]),
    "sgpu_iir_state_len": (c_size, [vp]),
    "sgpu_iir_set_mode": (C.c_int, [vp, C.c_int]),
    "sgpu_iir_decay_length": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "sgpu_iir_transition": (C.c_int, [vp, C.c_uint64, c_dp, c_sizep]),
    "sgpu_autocorr_create": (C.c_int, [c_size, c_size, c_size, vpp]),
    "sgpu_autocorr_destroy": (C.c_int, [vp]),
    "sgpu_autocorr_clone": (C.c_int, [vp, vpp]),
    "sgpu_autocorr_window_size": (c_size, [vp]),
    "sgpu_autocorr_delay": (c_size, [vp]),
    "sgpu_autocorr_channels": (c_size, [vp]),
    "sgpu_autocorr_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep, C.c_int, vp]),
    "sgpu_autocorr_write": (C.c_int, [vp, vp, c_size, c_size, C.c_int, vp]),
    "sgpu_autocorr_execute": (C.c_int, [vp, c_dp]),
    "sgpu_autocorr_get_energy": (C.c_int, [vp, c_dp]),
    "sgpu_autocorr_reset": (C.c_int, [vp]),
    "sgpu_autocorr_get_state": (C.c_int, [vp, vp]),
    "sgpu_autocorr_set_state": (C.c_int, [vp, vp]),
    "sgpu_dot_create": (C.c_int, [c_dp, c_size, C.c_int, C.c_int, vpp]),
    "sgpu_dot_destroy": (C.c_int, [vp]),
    "sgpu_dot_len": (c_size, [vp]),
    "sgpu_dot_coefficients": (C.c_int, [vp, c_dp]),
    "sgpu_dot_execute": (C.c_int, [vp, vp, c_size, c_size, c_size, vp, C.c_int, vp]),
    "sgpu_firdes_kaiser": (C.c_int, [c_size, c_dp, c_dp, c_dp, c_size, vp, C.c_int, vp]),
    "sgpu_nco_create": (C.c_int, [c_size, vpp]),
    "sgpu_nco_destroy": (C.c_int, [vp]),
    "sgpu_nco_clone": (C.c_int, [vp, vpp]),
    "sgpu_nco_channels": (c_size, [vp]),
    "sgpu_nco_reset": (C.c_int, [vp]),
    "sgpu_nco_set_frequency": (C.c_int, [vp, c_size, C.c_double]),
    "sgpu_nco_adjust_frequency": (C.c_int, [vp, c_size, C.c_double]),
    "sgpu_nco_set_phase": (C.c_int, [vp, c_size, C.c_double]),
    "sgpu_nco_adjust_phase": (C.c_int, [vp, c_size, C.c_double]),
    "sgpu_nco_step": (C.c_int, [vp, C.c_uint64]),
    "sgpu_nco_get": (C.c_int, [vp, c_size, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "sgpu_nco_set": (C.c_int, [vp, c_size, C.c_uint32, C.c_uint32]),
    "sgpu_nco_constrain": (C.c_uint32, [C.c_double]),
    "sgpu_nco_mix_block": (C.c_int, [vp, C.c_int, vp, c_size, c_size, vp, c_size, C.c_int, vp]),
    "sgpu_ddc_create": (C.c_int, [c_dp, c_size, C.c_int, c_size, C.c_double, C.c_double, c_size, vpp]),
    "sgpu_ddc_destroy": (C.c_int, [vp]),
    "sgpu_ddc_clone": (C.c_int, [vp, vpp]),
    "sgpu_ddc_filter": (vp, [vp]),
    "sgpu_ddc_nco": (vp, [vp]),
    "sgpu_ddc_out_len": (c_size, [vp, c_size]),
    "sgpu_ddc_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep, C.c_int, vp]),
    "sgpu_ddc_write": (C.c_int, [vp, vp, c_size, c_size, C.c_int, vp]),
    "sgpu_ddc_reset": (C.c_int, [vp]),
    "sgpu_ddc_last_fused": (C.c_int, [vp]),
    "sgpu_ctx_create": (C.c_int, [C.c_int, vpp]),
    "sgpu_ctx_create_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int, vpp]),
    "sgpu_ctx_destroy": (C.c_int, [vp]),
    "sgpu_ctx_devices": (C.c_int, [vp]),
    "sgpu_ctx_fir_create": (C.c_int, [vp, c_dp, c_size, C.c_int, c_size, C.c_double, C.c_double, C.c_int, c_size, vpp]),
    "sgpu_ctx_interp_create": (C.c_int, [vp, c_dp, c_size, C.c_int, c_size, c_size, vpp]),
    "sgpu_ctx_iir_create": (C.c_int, [vp, C.c_int, c_dp, c_size, c_dp, c_size, c_size, C.c_int, c_size, vpp]),
    "sgpu_sharded_destroy": (C.c_int, [vp]),
    "sgpu_sharded_shards": (C.c_int, [vp]),
    "sgpu_sharded_shard_info": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), c_sizep, c_sizep]),
    "sgpu_sharded_last_segments": (C.c_int, [vp]),
    "sgpu_sharded_out_len": (c_size, [vp, c_size]),
    "sgpu_sharded_reset": (C.c_int, [vp]),
    "sgpu_sharded_execute_block": (C.c_int, [vp, vp, c_size, c_size, vp, c_size, c_sizep]),
    "sgpu_shard_channels": (C.c_int, [c_size, C.c_int, C.c_int, c_sizep, c_sizep]),
    "sgpu_shard_stream": (C.c_int, [c_size, c_size, C.c_int, C.c_int, c_sizep, c_sizep]),
}


def header_symbols() -> list[str]:
    """Every function name declared in include/solid_gpu.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z0-9_]+)\s*\(", text)))


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C solid_dsp_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)  # AttributeError here = the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


lib = _load()


class SolidGpuError(RuntimeError):
    """A non-zero sgpu_status.  .status is the code, .name the reference enum variant."""

    def __init__(self, status: int):
        self.status = status
        self.name = lib.sgpu_status_name(status).decode()
        msg = lib.sgpu_last_error().decode()
        super().__init__(f"{self.name} ({status}): {msg}")


def check(status: int) -> None:
    if status != OK:
        raise SolidGpuError(status)


def launch_count() -> int:
    return int(lib.sgpu_launch_count())


def device_info() -> dict:
    dev, sms, maj, mnr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    mem = c_size()
    check(lib.sgpu_device_info(C.byref(dev), C.byref(sms), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return {"device": dev.value, "sm_count": sms.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}


_peak = None


def peak_lib():
    """libsgpu_peakbench.so: FP32-FMA / copy peaks for the roofline denominators."""
    global _peak
    if _peak is None:
        if not PEAK_LIB_PATH.exists():
            raise ImportError(f"{PEAK_LIB_PATH} is missing; run __graft_entry__.build()")
        P = C.CDLL(str(PEAK_LIB_PATH))
        P.sgpu_peak_fma.restype = C.c_int
        P.sgpu_peak_fma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        P.sgpu_peak_fma_ex.restype = C.c_int
        P.sgpu_peak_fma_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        P.sgpu_peak_copy.restype = C.c_int
        P.sgpu_peak_copy.argtypes = [c_size, C.c_int, c_dp, c_dp]
        _peak = P
    return _peak
