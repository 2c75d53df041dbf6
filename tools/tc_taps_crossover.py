import os, sys, torch
sys.path.insert(0, "/root/repo")
import oracle as O
from solid_dsp_b200.filter.fir import FIRFilter
from tests._util import f32_taps
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 27
x = torch.empty(n, dtype=torch.complex64, device="cuda"); torch.view_as_real(x).uniform_(-1, 1, generator=g)
os.environ["SGPU_FIR_TC_MIN_TAPS"] = "1"
import numpy as np
cases = [(T, False) for T in (64, 96, 128, 160, 192)] + [(T, True) for T in (32, 64, 96, 128)]
if len(sys.argv) > 1 and sys.argv[1] == "complex":
    cases = [c for c in cases if c[1]]
for T, cplx in cases:
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    if cplx:
        hc = h * np.exp(2j * np.pi * 0.05 * np.arange(T))
        h = f32_taps(hc.real) + 1j * f32_taps(hc.imag)
    for tc in ("0", "1"):
        os.environ["SGPU_FIR_TC"] = tc
        f = FIRFilter(h, 1.0); f.execute_block(x); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f.execute_block(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(f"{'complex' if cplx else 'real'} taps T={T} path={f.last_path}: {n / min(ts) / 1e6:.1f} Gsamp/s", flush=True)
