"""How does the tcgen05 tf32 MMA accumulate?  Feeds the tensor-core FIR inputs that are exact in TF32 (so the
lo planes are zero and the only error left is the accumulation inside the tensor core), random and DC, and
prints size and sign of the error against an f64 reference."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
from solid_dsp_b200.filter.fir import FIRFilter  # noqa: E402


def tf32(a):
    a = np.asarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def case(name, h, x_np, exact):
    n = x_np.shape[0]
    T = len(h)
    os.environ.setdefault("SGPU_FIR_TC", "1")
    x = torch.from_numpy(x_np).cuda()
    y = FIRFilter(h.astype(np.float64), 1.0).execute_block(x)
    lo = n // 2
    ref = O.fir_fast(h.astype(np.float64), x_np[lo - (T - 1):lo + 8192])[T - 1:]
    got = y[lo:lo + 8192].cpu().numpy()
    d = (got.astype(np.complex128) - ref)
    den = np.max(np.abs(ref))
    print(f"{name:34s} T={T:5d} exact_inputs={exact}: max|err|/max|ref| = {np.max(np.abs(d)) / den:.3e}  "
          f"mean(err.re)/max|ref| = {np.mean(d.real) / den:+.3e}  mean(ref.re)/max|ref| = {np.mean(ref.real) / den:+.3e}",
          flush=True)


def main():
    n = 1 << 21
    rng = np.random.default_rng(5)
    for T in (256, 512, 1024, 2048):
        hk = np.asarray(O.firdes_kaiser(T, 0.1, 80.0, 0.0), dtype=np.float32)
        hw = (np.hanning(T + 2)[1:-1] / T).astype(np.float32)  # all positive
        xr = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        xdc = np.full(n, 0.7 + 0.3j, dtype=np.complex64)
        xpos = (rng.uniform(0, 1, n) + 1j * rng.uniform(0, 1, n)).astype(np.complex64)

        def ex(z):
            return (tf32(z.real.copy()) + 1j * tf32(z.imag.copy())).astype(np.complex64)

        case("kaiser, random", hk, xr, False)
        case("kaiser, random", tf32(hk), ex(xr), True)
        case("hann(+), DC +", hw, xdc, False)
        case("hann(+), DC +", tf32(hw), ex(xdc), True)
        case("hann(+), DC -", tf32(hw), ex(-xdc), True)
        case("hann(+), random positive", tf32(hw), ex(xpos), True)
        case("hann(+), random positive", hw, xpos, False)


if __name__ == "__main__":
    main()
