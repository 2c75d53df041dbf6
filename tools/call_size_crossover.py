"""Per-call completion time of sgpu_fir_execute_block (device pointers, back-to-back calls on one handle) on the tensor
kernel and on the FFMA2 kernel, over call sizes: the measurement behind SGPU_FIR_TC_MIN_SAMPLES (csrc/fir.cu).
usage: python tools/call_size_crossover.py [taps ...]"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (tap design only)
from solid_dsp_b200 import _ffi  # noqa: E402
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps  # noqa: E402

os.environ["SGPU_FIR_TC_MIN_SAMPLES"] = "1"
taps = [int(t) for t in sys.argv[1:]] or [128, 512, 2048]
fn = _ffi.lib.sgpu_fir_execute_block
stream = torch.cuda.current_stream().cuda_stream
for T in taps:
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    for lg in range(13, 24):
        n = 1 << lg
        x = torch.zeros(n, dtype=torch.complex64, device="cuda")
        torch.view_as_real(x).uniform_(-1, 1)
        y = torch.empty(n, dtype=torch.complex64, device="cuda")
        got = _ffi.c_size()
        res = {}
        for path in ("1", "0"):
            os.environ["SGPU_FIR_TC"] = path
            f = FIRFilter(h, 1.0)
            for _ in range(5):
                _ffi.check(fn(f._h, x.data_ptr(), n, n, y.data_ptr(), n, C.byref(got), _ffi.DEVICE, stream))
            torch.cuda.synchronize()
            reps = 200 if lg <= 18 else 40
            t0 = time.perf_counter()
            for _ in range(reps):
                fn(f._h, x.data_ptr(), n, n, y.data_ptr(), n, C.byref(got), _ffi.DEVICE, stream)
            torch.cuda.synchronize()
            res[path] = (1e6 * (time.perf_counter() - t0) / reps, f.last_path)
            del f
        print(f"T={T} n=2^{lg}: tensor {res['1'][0]:8.1f} us ({res['1'][1]})   ffma2 {res['0'][0]:8.1f} us ({res['0'][1]})"
              f"   ratio {res['0'][0] / res['1'][0]:.2f}", flush=True)
