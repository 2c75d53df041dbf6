"""Per-call completion time of the tensor FIR path (main kernel + fir_tc_post_kernel) with and without programmatic
dependent launch of the post kernel (SGPU_FIR_TC_PDL).  usage: python tools/tc_pdl_probe.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import oracle as O
from solid_dsp_b200 import _ffi
from solid_dsp_b200.filter.fir import FIRFilter
from tests._util import f32_taps
os.environ["SGPU_FIR_TC_MIN_SAMPLES"] = "1"
fn = _ffi.lib.sgpu_fir_execute_block
stream = torch.cuda.current_stream().cuda_stream
h = f32_taps(O.firdes_kaiser(512, 0.1, 80.0, 0.0))
for lg in (16, 19, 20, 21, 23):
    n = 1 << lg
    x = torch.zeros(n, dtype=torch.complex64, device="cuda"); torch.view_as_real(x).uniform_(-1, 1)
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    got = _ffi.c_size()
    res = {}
    for pdl in ("1", "0"):
        os.environ["SGPU_FIR_TC_PDL"] = pdl
        f = FIRFilter(h, 1.0)
        for _ in range(5): _ffi.check(fn(f._h, x.data_ptr(), n, n, y.data_ptr(), n, C.byref(got), _ffi.DEVICE, stream))
        torch.cuda.synchronize()
        reps = 100
        t0 = time.perf_counter()
        for _ in range(reps): fn(f._h, x.data_ptr(), n, n, y.data_ptr(), n, C.byref(got), _ffi.DEVICE, stream)
        torch.cuda.synchronize()
        res[pdl] = 1e6 * (time.perf_counter() - t0) / reps
    print(f"T=512 n=2^{lg}: tensor per call with PDL {res['1']:.1f} us, without {res['0']:.1f} us", flush=True)
