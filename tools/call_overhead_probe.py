"""Host-side cost of one sgpu_fir_execute_block call (device pointers, tiny input): wall time per call over many
back-to-back calls, through ctypes directly (no Python wrapper objects)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from solid_dsp_b200 import _ffi  # noqa: E402
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps  # noqa: E402

for T, n in ((64, 4096), (64, 1 << 20), (512, 4096)):
    h = f32_taps(O.firdes_kaiser(T, 0.25, 60.0, 0.0))
    f = FIRFilter(h, 1.0)
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    got = _ffi.c_size()
    stream = torch.cuda.current_stream().cuda_stream
    fn = _ffi.lib.sgpu_fir_execute_block

    def call():
        _ffi.check(fn(f._h, x.data_ptr(), n, n, y.data_ptr(), n, C.byref(got), _ffi.DEVICE, stream))

    for _ in range(20):
        call()
    torch.cuda.synchronize()
    reps = 2000
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"T={T} n={n}: host {1e6 * (t1 - t0) / reps:.1f} us per call issued, {1e6 * (t2 - t0) / reps:.1f} us per call completed",
          flush=True)
