"""Randomised parity sweep (tools/fuzz_parity.py): every filter type through the public mirror against the oracle with
random tap counts, factors, channel counts, lengths, call splits, padded row strides, 8-byte-aligned bases, real / complex
taps and scales, per-channel taps, host and device memory.  The seeds are fixed: a failure prints the trial's parameters."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("seed", [7, 8])
def test_randomised_parity_sweep(seed):
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "fuzz_parity.py"), "150", str(seed)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


def test_randomised_parity_sweep_tensor_kernels():
    """The same sweep aimed at the tcgen05 kernels: long real- and complex-tap FIR filters (112 ... 3001 taps; strip, chain
    and complex-tap variants), interpolators with long sub-filters, several channels, split calls, padded strides and odd
    bases; every call must have taken the tensor path."""
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "fuzz_parity.py"), "40", "5", "tensor"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
