"""Interpolating IIR wrapper (iir/interp.rs:184-221) on the tile path: G out-samp/s and oracle error per factor L.
usage: python tools/iir_interp_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from solid_dsp_b200.filter.iir import IIRFilterType, InterpolatingIIRFilter  # noqa: E402
from solid_dsp_b200.filter.iirdes import stable_lowpass_sections  # noqa: E402

ff, fb = stable_lowpass_sections(8)
ff, fb = np.asarray(ff, np.float32).astype(np.float64), np.asarray(fb, np.float32).astype(np.float64)
C, n = 32768, 4099
x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
for L in (2, 3, 4, 5, 6, 7, 8, 12, 31, 33):
    f = InterpolatingIIRFilter(ff, fb, IIRFilterType.SecondOrder, L, n_channels=C)
    y = f.execute_block(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        f.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = f.execute_block(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    errs = []
    for c in (0, C // 2 + 1, C - 1):
        up = np.zeros(n * L, dtype=np.complex128)
        up[::L] = x[c].cpu().numpy()
        ref = O.sos_cascade_fast(ff, fb, up)[0]
        errs.append(float(np.max(np.abs(y[c].cpu().numpy() - ref)) / np.max(np.abs(ref))))
    print(f"L={L:2d}: {ms:.3f} ms  {C * n * L / ms / 1e6:.0f} G out-samp/s  nerr={max(errs):.2e}", flush=True)
    del y, f
