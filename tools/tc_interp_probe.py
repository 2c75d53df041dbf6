"""Interpolator on the tensor-core kernel vs the FP32 walking / tile kernels: crossover in sub-filter length.
Usage: python tools/tc_interp_probe.py   (needs a B200)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
from solid_dsp_b200.filter.fir import InterpolatingFIRFilter  # noqa: E402
from tests._util import f32_taps, nerr  # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(4)
    C, n = 256, 1 << 20
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    for L in (4, 2):
        for S in (24, 32, 48, 64, 96, 128, 256):
            T = S * L
            h = f32_taps(O.firdes_kaiser(T, 0.5 / L * 0.9, 80.0, 0.0))
            for tc in ("0", "1"):
                os.environ["SGPU_FIR_TC"] = tc
                os.environ["SGPU_INTERP_TC_MIN_SUB"] = "1"
                f = InterpolatingFIRFilter(h, L, n_channels=C)
                y = f.execute_block(x)
                torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    f.execute_block(x)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = min(ts)
                s0 = n - 3000
                ref = O.firinterp_fast(h, L, x[7, s0 - S:].cpu().numpy())[S * L:]
                err = nerr(y[7, s0 * L:].cpu().numpy(), ref)
                print(f"L={L} S={S:4d} path={f.last_path:6s}: {ms:.3f} ms  {C * n * L / ms / 1e6:.0f} G out-samp/s  nerr={err:.2e}",
                      flush=True)


if __name__ == "__main__":
    main()
