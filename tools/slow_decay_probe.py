"""Fused warm-up scan vs three-pass scan for IIR cascades with a long memory (pole radius 0.99 ... 0.9999):
the measurements behind the strategy rule in iir_run.  usage: python tools/slow_decay_probe.py"""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
from tools.explore import ev_time
for r in (0.99, 0.999, 0.9999):
    fb = np.float32([1.0, -2 * r * np.cos(0.05 * np.pi), r * r]).astype(np.float64)
    ff = np.float32([1e-3, 2e-3, 1e-3]).astype(np.float64)
    x = torch.randn(1 << 28, dtype=torch.complex64, device="cuda")
    for mode in (-1, 2):
        f = IIRFilter(np.tile(ff, 8), np.tile(fb, 8), IIRFilterType.SecondOrder)
        f.set_mode(mode)
        best, med = ev_time(lambda: f.execute_block(x))
        print("r", r, "decay", f.decay_length(), "mode", mode, "Gsamp/s", round((1 << 28) / best / 1e6, 1))
        del f
