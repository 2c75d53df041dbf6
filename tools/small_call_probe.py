"""Per-call time of short FIR calls (BASELINE config 1: 64 taps x 2^20 samples) with one / eight tiles per warp."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps, nerr  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
for T, lg in ((64, 20), (64, 17), (64, 23), (128, 20), (16, 20)):
    n = 1 << lg
    h = f32_taps(O.firdes_kaiser(T, 0.25, 60.0, 0.0))
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    for small in ("0", "1"):
        os.environ["SGPU_FIR_SMALL"] = small
        f = FIRFilter(h, 1.0)
        y = f.execute_block(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f.execute_block(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ref = O.fir_fast(h, x[:8192 + T].cpu().numpy())
        print(f"T={T} n=2^{lg} one_tile_per_warp={small}: {min(ts) * 1e3:.1f} us  {n / min(ts) / 1e6:.1f} Gsamp/s  "
              f"nerr={nerr(y[:8192 + T].cpu().numpy(), ref):.2e}", flush=True)
