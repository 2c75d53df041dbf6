"""DDC (NCO mix-down fused into the decimator) against the plain decimator on the same buffers: device time per call.
usage: python tools/ddc_probe.py [channels] [log2_samples]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (tap design only)
from solid_dsp_b200.filter.ddc import DigitalDownConverter  # noqa: E402
from solid_dsp_b200.filter.fir import DecimatingFIRFilter  # noqa: E402
from tests._util import f32_taps  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(os.environ.get("REPS", "5"))
h = f32_taps(O.firdes_kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
x = torch.empty((C, 1 << lg), dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
for name, f in (("decim", DecimatingFIRFilter(h, 1.0, 8, n_channels=C)), ("ddc", DigitalDownConverter(h, 1.0, 8, 0.1234, n_channels=C))):
    f.execute_block(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f.execute_block(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name}: {C} ch x 2^{lg}: {ms:.3f} ms  {C * (1 << lg) / ms / 1e6:.1f} G in-samp/s", flush=True)
