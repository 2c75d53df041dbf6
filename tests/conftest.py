import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
# The library sends a FIR call to the tensor kernel only when it is long enough to pay (2^19 samples and 2^29
# sample-taps, csrc/fir.cu: tc_call_is_long_enough).  The GPU tests want the tensor kernel exercised on short streams
# too (edge tiles, odd lengths, split calls), so they pin the plain per-channel threshold the kernel itself supports;
# tests/test_fir_tc_gpu.py::test_default_dispatch_rule removes the variable and checks the rule.
os.environ.setdefault("SGPU_FIR_TC_MIN_SAMPLES", "32768")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import json
    g = {}
    for name in ("reference_doctests", "derived_vectors"):
        with open(ROOT / "tests" / "golden" / f"{name}.json") as f:
            g[name] = json.load(f)
    return g
