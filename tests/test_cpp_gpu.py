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
