"""In-library multi-GPU context on the GPUs of this box: one 512-tap FIR stream from (pinned) host memory through
sgpu_sharded_execute_block -- time segments per GPU, halo sliced from the caller's buffer, outputs gathered in the
caller's host buffer -- against the single-GPU handle and the oracle.  usage: python tools/ctx_probe.py [log2_samples]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker)
from solid_dsp_b200.context import Context  # noqa: E402
from solid_dsp_b200.hostmem import PinnedArray  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n, T = 1 << lg, 512
h = np.asarray(O.firdes_kaiser(T, 0.1, 80.0, 0.0), dtype=np.float32).astype(np.float64)
xin = PinnedArray(1, n)
yout = PinnedArray(1, n)
rng = np.random.default_rng(3)
blk = 1 << 22
for a in range(0, n, blk):
    xin.array[0, a:a + blk] = (rng.uniform(-1, 1, blk) + 1j * rng.uniform(-1, 1, blk)).astype(np.complex64)
for ndev in sorted({1, torch.cuda.device_count()}):
    ctx = Context(ndev)
    f = ctx.fir(h, 1.0)
    y = f.execute_block(xin.array[0], out=yout.array)  # warm-up (allocations)
    ts = []
    for _ in range(3):
        f.reset()
        t0 = time.perf_counter()
        y = f.execute_block(xin.array[0], out=yout.array)
        ts.append(time.perf_counter() - t0)
    dt = min(ts)
    errs = []
    for s0 in (0, n // ndev - 100, n // 2 + 777, n - 4096):
        lo = max(0, s0 - (T - 1))
        ref = O.fir_fast(h, xin.array[0, lo:s0 + 4096].astype(np.complex128))[s0 - lo:]
        errs.append(float(np.max(np.abs(y[s0:s0 + 4096] - ref)) / np.max(np.abs(ref))))
    print(f"{ndev} GPU(s): 2^{lg} samples in {dt * 1e3:.1f} ms = {n / dt / 1e9:.2f} Gsamp/s end to end "
          f"({2 * n * 8 / dt / 1e9:.1f} GB/s moved), segments used {f.last_segments}, nerr {max(errs):.2e}", flush=True)
    del f, ctx
