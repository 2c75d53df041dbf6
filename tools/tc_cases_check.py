import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import oracle as O
from solid_dsp_b200.filter.fir import FIRFilter, InterpolatingFIRFilter
from tests._util import f32_taps, nerr
g = torch.Generator(device="cuda").manual_seed(3)
def rnd(*shape):
    x = torch.empty(shape, dtype=torch.complex64, device="cuda"); torch.view_as_real(x).uniform_(-1, 1, generator=g); return x
h = f32_taps(O.firdes_kaiser(512, 0.1, 80.0, 0.0))
x = rnd(16384 * 150 + 777)
f = FIRFilter(h, 1.0); y = f.execute_block(x); torch.cuda.synchronize()
print("fir", f.last_path, nerr(y[-3000:].cpu().numpy(), O.fir_fast(h, x[-3511:].cpu().numpy())[511:]))
hc = h * np.exp(2j*np.pi*0.05*np.arange(512)); hc = f32_taps(hc.real) + 1j*f32_taps(hc.imag)
f = FIRFilter(hc, 0.5-0.25j); y = f.execute_block(x[1:]); torch.cuda.synchronize()
print("cfir", f.last_path, nerr(y[-3000:].cpu().numpy(), O.fir_fast(hc, x[-3511:].cpu().numpy(), 0.5-0.25j)[511:]))
hi = f32_taps(O.firdes_kaiser(256, 0.1, 80.0, 0.0))
xi = rnd(3, 4096 * 200 + 5)
fi = InterpolatingFIRFilter(hi, 4, n_channels=3); yi = fi.execute_block(xi); torch.cuda.synchronize()
print("interp", fi.last_path, nerr(yi[2, -8000:].cpu().numpy(), O.firinterp_fast(hi, 4, xi[2, -2064:].cpu().numpy())[64*4:]))
os.environ["SGPU_INTERP_TC_MIN_SUB"] = "1"
hi = f32_taps(O.firdes_kaiser(128, 0.1, 80.0, 0.0))
fi = InterpolatingFIRFilter(hi, 4, n_channels=3); yi = fi.execute_block(xi); torch.cuda.synchronize()
print("interp-one", fi.last_path, nerr(yi[2, -8000:].cpu().numpy(), O.firinterp_fast(hi, 4, xi[2, -2032:].cpu().numpy())[32*4:]))
