"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/solid_gpu.h declares, and the ctypes table covers the header exactly.  No compute calls."""
import ctypes
import subprocess

import pytest


@pytest.fixture(scope="module")
def ffi():
    from solid_dsp_b200 import _ffi
    return _ffi


def test_header_and_binding_agree(ffi):
    declared = set(ffi.header_symbols())
    bound = set(ffi.PROTOTYPES)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_library_exports_every_declared_symbol(ffi):
    out = subprocess.run(["nm", "-D", "--defined-only", str(ffi.LIB_PATH)], capture_output=True, text=True,
                         check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = set(ffi.header_symbols()) - exported
    assert not missing, missing
    L = ctypes.CDLL(str(ffi.LIB_PATH))
    for name in ffi.header_symbols():
        assert hasattr(L, name)


def test_host_only_entry_points(ffi):
    assert ffi.lib.sgpu_abi_version() == 1
    assert ffi.lib.sgpu_status_name(-1) == b"FIRErrorCode::CoefficientsLengthZero"
    assert ffi.lib.sgpu_status_name(-14) == b"IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3"
    first, count = ffi.c_size(), ffi.c_size()
    covered = 0
    for r in range(8):
        assert ffi.lib.sgpu_shard_channels(4099, 8, r, ctypes.byref(first), ctypes.byref(count)) == 0
        assert first.value == covered
        covered += count.value
    assert covered == 4099
    covered = 0
    for r in range(3):
        assert ffi.lib.sgpu_shard_stream(1000, 8, 3, r, ctypes.byref(first), ctypes.byref(count)) == 0
        assert first.value == covered and first.value % 8 == 0
        covered += count.value
    assert covered == 1000
    assert ffi.lib.sgpu_shard_stream(10, 0, 1, 0, ctypes.byref(first), ctypes.byref(count)) == ffi.ERR_INVALID_ARGUMENT


def test_construction_argument_errors_need_no_gpu(ffi):
    """Argument validation happens before any CUDA call, with the reference's error variants."""
    from solid_dsp_b200.filter.fir import (DecimatingFIRFilter, FIRError, FIRFilter,
                                           InterpolatingFIRFilter, PolyPhaseFilterBank)
    for ctor, code in [
        (lambda: FIRFilter([], 1.0), "CoefficientsLengthZero"),
        (lambda: DecimatingFIRFilter([], 1.0, 2), "CoefficientsLengthZero"),
        (lambda: DecimatingFIRFilter([1.0], 1.0, 0), "DecimationLessThanOne"),
        (lambda: InterpolatingFIRFilter([], 2), "CoefficientsLengthZero"),
        (lambda: InterpolatingFIRFilter([1.0], 0), "InterpolationLessThanOne"),
        (lambda: PolyPhaseFilterBank([1.0], 0), "NotEnoughFilters"),
        (lambda: PolyPhaseFilterBank([], 2), "CoefficientsLengthZero"),
    ]:
        with pytest.raises(FIRError) as e:
            ctor()
        assert e.value.code == code


def test_no_cpu_fallback_without_gpu(ffi):
    """On a box without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from solid_dsp_b200 import SolidGpuError
    from solid_dsp_b200.filter.fir import FIRFilter
    with pytest.raises(SolidGpuError) as e:
        FIRFilter([1.0, 2.0, 3.0], 1.0)
    assert e.value.status == ffi.ERR_NO_DEVICE


def test_rust_sys_build_compiles_the_same_sources_as_the_makefile():
    """rust/solid-gpu-sys/build.rs (the -sys crate the north star asks for; no Rust toolchain here to run it) must list
    exactly the CUDA sources libsolid_gpu.so is built from."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    mk = (root / "solid_dsp_b200" / "csrc" / "Makefile").read_text()
    srcs = re.search(r"^SRCS\s*:=\s*(.+)$", mk, re.M).group(1).split()
    rs = (root / "rust" / "solid-gpu-sys" / "build.rs").read_text()
    listed = re.findall(r'"([a-z_]+\.cu)"', re.search(r"let sources = \[(.+?)\];", rs, re.S).group(1))
    assert sorted(listed) == sorted(srcs)


def test_rust_sys_crate_declares_every_prototype_of_the_header():
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    header = set(re.findall(r"\b(sgpu_[a-z0-9_]+)\s*\(", (root / "include" / "solid_gpu.h").read_text()))
    rust = set(re.findall(r"pub fn (sgpu_[a-z0-9_]+)\s*\(", (root / "rust" / "solid-gpu-sys" / "src" / "lib.rs").read_text()))
    assert header == rust, (sorted(header - rust), sorted(rust - header))
