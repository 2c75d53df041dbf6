"""IIR cascade (8 sections) on channel batches: batch kernel (mode 0), fused scan (mode 1) and the automatic choice (-1) at
a few shapes -- the 1.15-wave tail at 65536 channels costs 2-3 % (0.836 vs 0.858 of HBM for exactly one wave).
usage: python tools/iir_mode_probe.py"""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
ff, fb = stable_lowpass_sections(8)
for C, n in ((65536, 1 << 14), (56832, 1 << 14), (8192, 1 << 14), (16384, 1 << 16)):
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1)
    for mode in (-1, 0, 1):
        f = IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=C)
        f.set_mode(mode)
        y = f.execute_block(x); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y = f.execute_block(x); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"C={C} n={n} mode={mode}: {ms:.3f} ms  {C * n / ms / 1e6:.0f} Gsamp/s  {C * n * 16 / ms / 1e6 / 6555.5:.3f} of HBM", flush=True)
        del f, y
    del x
