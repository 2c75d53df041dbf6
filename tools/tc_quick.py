"""Device time and oracle error of the tensor-core FIR for the environment it is started in (one line per tap count).
usage: [SGPU_FIR_TC_...=..] python tools/tc_quick.py [log2_samples] [taps ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps, nerr  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
taps = [int(t) for t in sys.argv[2:]] or [512, 256, 2048]
n = 1 << lg
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1, generator=g)
tag = " ".join(f"{k[9:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SGPU_FIR_"))
for T in taps:
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    f = FIRFilter(h, 1.0)
    y = f.execute_block(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f.execute_block(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    errs = []
    for start in (0, T - 1, n // 2 + 12345, n - 4096):
        lo = max(0, start - (T - 1))
        ref = O.fir_fast(h, x[lo:start + 4096].cpu().numpy())[start - lo:]
        errs.append(nerr(y[start:start + 4096].cpu().numpy(), ref))
    ms = sorted(ts)[len(ts) // 2]
    print(f"[{tag}] T={T} n=2^{lg} path={f.last_path}: median {ms:.3f} ms  {n / ms / 1e6:.1f} Gsamp/s  nerr={max(errs):.2e}", flush=True)
