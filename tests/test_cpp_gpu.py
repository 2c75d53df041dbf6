"""Runs the compiled C++ host mirror's parity program (tests/cpp/test_parity.cpp: include/solid.hpp
over the C ABI versus the oracle) on the GPU box; on CPU only checks that it builds."""
import subprocess
from pathlib import Path

import pytest

CPP = Path(__file__).resolve().parent / "cpp"


def _build():
    subprocess.run(["make", "-C", str(CPP)], check=True, capture_output=True)
    return CPP / "_build" / "test_parity"


def test_cpp_mirror_builds():
    assert _build().exists()


def test_cpp_host_types():
    """Window<T> / CircularBuffer<T> of the C++ mirror (include/solid_host.hpp, no GPU): the reference's behaviour and error
    codes, compiled and run on the CPU tier."""
    _build()
    r = subprocess.run([str(CPP / "_build" / "test_host_types")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_parity():
    r = subprocess.run([str(_build())], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASSED" in r.stdout


def _build_c_example():
    root = CPP.parent.parent
    out = CPP / "_build" / "fir_stream"
    (CPP / "_build").mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", f"-I{root / 'include'}", str(root / "examples" / "c" / "fir_stream.c"),
                    f"-L{root / 'solid_dsp_b200' / 'lib'}", "-lsolid_gpu", "-lm",
                    f"-Wl,-rpath,{root / 'solid_dsp_b200' / 'lib'}", "-o", str(out)], check=True, capture_output=True)
    return out


def test_c_example_builds_and_fails_loudly_without_a_gpu():
    """examples/c/fir_stream.c: a plain C caller of include/solid_gpu.h (the header is C, not only C++).  Without a GPU
    the first call returns SGPU_ERR_NO_DEVICE -- there is no CPU fallback."""
    import torch
    exe = _build_c_example()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_c_example_runs")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "SGPU_ERR_NO_DEVICE" in r.stderr


@pytest.mark.gpu
def test_c_example_runs():
    r = subprocess.run([str(_build_c_example())], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
