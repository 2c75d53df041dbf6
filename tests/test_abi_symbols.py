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


# the public surface of the reference crate's filtering path, per module of rust/solid/src.
# (type name, methods); `mod`s nested in a file are found by name, the check is textual (no Rust toolchain here).
_RUST_SURFACE = {
    "dot_product.rs": [("enum Direction", []), ("struct DotProduct", ["new", "coefficents", "len", "is_empty"]),
                       ("trait Execute", ["execute"])],
    "window.rs": [("struct Window", ["new", "as_ptr", "to_vec", "reset", "capacity", "push", "write"])],
    "circular_buffer.rs": [("enum BufferErrorCode", []), ("struct BufferError", []),
                           ("struct CircularBuffer", ["new", "from_vec", "from_slice", "as_ptr", "as_mut_ptr", "linearize",
                                                      "to_vec", "reset", "len", "capacity", "reserved", "is_empty", "is_full",
                                                      "read_index", "write_index", "push", "append", "pop", "release"])],
    "filter/mod.rs": [("trait Filter", ["execute", "execute_block", "frequency_response", "group_delay"])],
    "filter/fir.rs": [("enum FIRErrorCode", []), ("struct FIRError", []),
                      ("struct FIRFilter", ["new", "set_scale", "get_scale", "len", "is_empty", "coefficients"]),
                      ("struct DecimatingFIRFilter", ["new", "set_scale", "get_scale", "get_decimation", "push", "write", "len",
                                                      "is_empty", "coefficients"]),
                      ("struct InterpolatingFIRFilter", ["new", "set_scale", "get_scale", "len", "is_empty", "coefficents",
                                                         "interpolation"]),
                      ("struct PolyPhaseFilterBank", ["new", "set_scale", "get_scale", "len", "is_empty", "coefficents", "reset",
                                                      "push", "execute"])],
    "filter/iir.rs": [("enum IIRErrorCode", []), ("struct IIRError", []), ("enum IIRFilterType", []),
                      ("struct IIRFilter", ["new", "numerator_coefs", "denominator_coefs", "second_order_filters", "iir_type"]),
                      ("struct SecondOrderFilter", ["new", "execute", "numerator_coefs", "denominator_coefs", "frequency_response",
                                                    "group_delay"]),
                      ("struct DecimatingIIRFilter", ["new", "get_decimation", "numerator_coefs", "denominator_coefs", "iir_type"]),
                      ("struct InterpolatingIIRFilter", ["new", "get_interpolation", "numerator_coefs", "denominator_coefs",
                                                         "iir_type"])],
}


def _rust_impl_bodies(src, type_name):
    """Concatenated bodies of every `impl … type_name… { … }` block (inherent and trait impls) plus, for a trait, its own
    body: brace matching on the source text."""
    import re
    out = []
    for m in re.finditer(r"\b(?:impl\b[^{;]*?\b%s\b[^{;]*|trait\s+%s\b[^{;]*)\{" % (type_name, type_name), src):
        depth, i = 1, m.end()
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        out.append(src[m.end():i])
    return "\n".join(out)


def test_rust_safe_crate_has_the_reference_surface():
    """Every type and `pub fn` of SURVEY.md Appendix C exists in rust/solid/src with the reference's name (including the
    reference's own spelling `coefficents`), no body is `unimplemented!()` / `todo!()`, and the filters are generic over
    (Coef, In) like the reference's (filter/fir/mod.rs:58-63; main.rs:39 instantiates `IIRFilter::<f64, Complex<f64>>`)."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parents[1] / "rust" / "solid" / "src"
    for rel, items in _RUST_SURFACE.items():
        src = (root / rel).read_text()
        assert "unimplemented!" not in src and "todo!" not in src, rel
        for decl, methods in items:
            kind, name = decl.split()
            assert re.search(r"\bpub\s+%s\s+%s\b" % (kind, name), src), f"{rel}: missing `pub {decl}`"
            body = _rust_impl_bodies(src, name)
            for fn in methods:
                pat = r"\bfn\s+%s\s*[<(]" % fn if kind == "trait" else r"\bpub\s+fn\s+%s\s*[<(]" % fn
                # trait methods implemented for the type (Filter::execute …) count as well
                assert re.search(pat, body) or re.search(r"\bfn\s+%s\s*[<(]" % fn, body), f"{rel}: {name}::{fn} missing"
    fir = (root / "filter" / "fir.rs").read_text()
    iir = (root / "filter" / "iir.rs").read_text()
    for name in ("FIRFilter", "DecimatingFIRFilter", "InterpolatingFIRFilter", "PolyPhaseFilterBank"):
        assert re.search(r"pub struct %s<Coef: Coefficient, In: Sample>" % name, fir), name
    assert re.search(r"pub struct IIRFilter<Coef, In: Sample>", iir)
    lib = (root / "lib.rs").read_text()
    for mod in ("dot_product", "window", "circular_buffer", "filter"):
        assert re.search(r"pub mod %s;" % mod, lib), mod
    scalar = (root / "scalar.rs").read_text()
    assert "impl Sample for Complex<f64>" in scalar and "impl Coefficient for Complex<f64>" in scalar
