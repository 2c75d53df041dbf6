"""Tensor-core FIR (fir_tc.cu) against the FFMA2 kernel and the oracle: error and device time.
Usage: python tools/tc_probe.py [log2_samples ...]   (needs a B200; prints one line per case)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402
from tests._util import f32_taps, nerr  # noqa: E402


MODES = {"ffma": ("0", "2", "f16"), "tc-f16x2": ("1", "2", "f16"), "tc-bf16x3": ("1", "2", "bf16")}


def run(h, x, tc, reps=3):
    os.environ["SGPU_FIR_TC"], os.environ["SGPU_FIR_TC_CHAIN"], os.environ["SGPU_FIR_TC_FMT"] = MODES[tc]
    f = FIRFilter(h, 1.0)
    y = f.execute_block(x)  # warm-up (allocates scratch), from zero history
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        f2 = FIRFilter(h, 1.0) if False else f
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f2.execute_block(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return y, min(ts)


def main():
    logs = [int(a) for a in sys.argv[1:]] or [22, 26]
    g = torch.Generator(device="cuda").manual_seed(2)
    for T in (512, 256, 2048):
        h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
        for lg in logs:
            n = 1 << lg
            x = torch.empty(n, dtype=torch.complex64, device="cuda")
            torch.view_as_real(x).uniform_(-1, 1, generator=g)
            res = {}
            for tc in MODES:
                try:
                    y, ms = run(h, x, tc)
                except Exception as e:  # noqa: BLE001
                    print(f"T={T} n=2^{lg} tc={tc} FAILED: {e}", flush=True)
                    continue
                errs = []
                for start in (0, T - 1, n // 2 + 12345, n - 4096):
                    lo = max(0, start - (T - 1))
                    ref = O.fir_fast(h, x[lo:start + 4096].cpu().numpy())[start - lo:]
                    errs.append(nerr(y[start:start + 4096].cpu().numpy(), ref))
                res[tc] = y
                print(f"T={T} n=2^{lg} tc={tc}: {ms:.3f} ms  {n / ms / 1e6:.1f} Gsamp/s  nerr(windows)={max(errs):.3e}",
                      flush=True)
            for k in res:
                if k != "ffma" and "ffma" in res:
                    d = (res["ffma"] - res[k]).abs().max().item() / res["ffma"].abs().max().item()
                    print(f"T={T} n=2^{lg} max|{k} - ffma| / max|ffma| over the whole stream = {d:.3e}", flush=True)
            del x, res


if __name__ == "__main__":
    main()
