"""solid_dsp_b200 -- B200 (sm_100a) implementation of juliantos/solid-dsp's filtering hot path.

Layout mirrors the reference crate's modules for that path only:
    solid_dsp_b200.filter.fir   FIRFilter, DecimatingFIRFilter, InterpolatingFIRFilter, PolyPhaseFilterBank
    solid_dsp_b200.filter.iir   IIRFilter, SecondOrderFilter, DecimatingIIRFilter, InterpolatingIIRFilter
    solid_dsp_b200.dot_product  DotProduct, Direction
    solid_dsp_b200.window       Window           (host-side history type, window/mod.rs)
    solid_dsp_b200.circular_buffer CircularBuffer (host-side ring FIFO, circular_buffer/mod.rs)
    solid_dsp_b200.filter.firdes / iirdes  host f64 design helpers that feed the filters their taps (+ firdes_kaiser_device,
                                filter_energy_device: the same on the GPU)
    solid_dsp_b200.nco          NCO              (nco/mod.rs: 32-bit phase accumulators, 1024-entry table)
    solid_dsp_b200.filter.ddc   DigitalDownConverter (NCO mix-down fused into the decimating FIR)
    solid_dsp_b200.filter.auto_correlator  AutoCorrelator
    solid_dsp_b200.context      Context, ShardedFilter (every GPU of the box behind one filter object)
    solid_dsp_b200.hostmem      PinnedArray      (sgpu_host_alloc: pinned host buffers for the SGPU_HOST calls)
    solid_dsp_b200.sharding     one-process-per-GPU helpers (channel ranges, stream segments, halo exchange)

All execute paths call the CUDA library through the C ABI in include/solid_gpu.h; importing this
package fails loudly when libsolid_gpu.so is absent (no CPU fallback).
"""
from . import _ffi  # noqa: F401  (loads libsolid_gpu.so or raises)
from ._ffi import SolidGpuError, device_info, launch_count  # noqa: F401

__all__ = ["SolidGpuError", "device_info", "launch_count"]
