"""End-to-end rate of an SGPU_HOST call on ordinary (pageable) numpy memory -- a fresh output array per call, like the
reference's `Vec` -- against the same call on pinned buffers.  usage: python tools/pageable_probe.py [log2_samples]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (tap design, checker)
from solid_dsp_b200 import _ffi  # noqa: E402
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from solid_dsp_b200.hostmem import PinnedArray  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n, T = 1 << lg, 512
h = np.asarray(O.firdes_kaiser(T, 0.1, 80.0, 0.0), dtype=np.float32).astype(np.float64)
rng = np.random.default_rng(5)
x = np.empty(n, dtype=np.complex64)
for a in range(0, n, 1 << 22):
    x[a:a + (1 << 22)] = (rng.uniform(-1, 1, 1 << 22) + 1j * rng.uniform(-1, 1, 1 << 22)).astype(np.complex64)
for label, staging, env in (("driver-staged copies", "0", None), ("library staging, host threads", "1", None),
                            ("library staging, 1 thread", "1", "1"), ("library staging, 16 threads", "1", "16"),
                            ("library staging, 4 threads", "1", "4")):
    os.environ["SGPU_HOST_STAGING"] = staging
    if env:
        os.environ["SGPU_HOST_COPY_THREADS"] = env
    else:
        os.environ.pop("SGPU_HOST_COPY_THREADS", None)
    f = FIRFilter(h, 1.0)
    f.execute_block(x[:1 << 22])
    ts = []
    for _ in range(3):
        f.reset()
        t0 = time.perf_counter()
        y = f.execute_block(x)          # fresh pageable output array inside
        ts.append(time.perf_counter() - t0)
    err = float(np.max(np.abs(y[-4096:] - O.fir_fast(h, x[-4096 - 511:].astype(np.complex128))[511:])))
    print(f"pageable, {label}: {min(ts) * 1e3:.1f} ms = {n / min(ts) / 1e9:.2f} Gsamp/s  (abs err {err:.1e})", flush=True)
xp, yp = PinnedArray(1, n), PinnedArray(1, n)
xp.array[0] = x
f = FIRFilter(h, 1.0)
got = _ffi.c_size()
ts = []
for _ in range(4):
    f.reset()
    t0 = time.perf_counter()
    _ffi.check(_ffi.lib.sgpu_fir_execute_block(f._h, xp.ptr, n, n, yp.ptr, n, C.byref(got), _ffi.HOST, None))
    ts.append(time.perf_counter() - t0)
print(f"pinned in / out: {min(ts) * 1e3:.1f} ms = {n / min(ts) / 1e9:.2f} Gsamp/s", flush=True)
