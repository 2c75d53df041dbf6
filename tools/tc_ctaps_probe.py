"""Complex-tap FIR: tensor kernel vs the FP32 kernel (speed), needs a B200."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 27
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1, generator=g)
for T in (128, 256, 512, 2048):
    hr = O.firdes_kaiser(T, 0.1, 80.0, 0.0) * np.exp(2j * np.pi * 0.05 * np.arange(T))
    h = f32_taps(hr.real) + 1j * f32_taps(hr.imag)
    for tc in ("0", "1"):
        os.environ["SGPU_FIR_TC"] = tc
        f = FIRFilter(h, 1.0)
        f.execute_block(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f.execute_block(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"complex taps T={T} path={f.last_path}: {min(ts):.3f} ms  {n / min(ts) / 1e6:.1f} Gsamp/s", flush=True)
