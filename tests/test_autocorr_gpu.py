"""GPU parity: AutoCorrelator (filter/auto_correlator/mod.rs) versus the CPU oracle."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ac():
    from solid_dsp_b200.filter import auto_correlator
    return auto_correlator


def test_reference_doctest(ac):
    """auto_correlator/mod.rs:199-211: (energy * 10000).round() == 125 for window 5, delay 10."""
    x = np.array([complex(np.cos(float(k)) * 0.05, np.sin(float(k)) * 0.05) for k in range(-250, 250)])
    a = ac.AutoCorrelator(5, 10)
    out = a.execute_block(x)
    assert len(out) == 500 and np.all(out == 0)              # delay >= window: the quirk
    assert round(a.get_energy() * 10000.0) == 125.0
    b = ac.AutoCorrelator(10, 5)
    got = b.execute_block(x)
    assert nerr(got, O.AutoCorrelator(10, 5).execute_block(x)) <= TOL
    assert str(b).startswith("AutoCorrelator<f64> [Size=10] [Delay=5]")


@pytest.mark.parametrize("W,d", [(1, 0), (2, 1), (7, 0), (10, 5), (16, 15), (16, 16), (64, 16), (300, 100), (1000, 1)])
@pytest.mark.parametrize("n", [1, 9, 2047, 2048, 2049, 10001])
def test_random(ac, W, d, n):
    rng = np.random.default_rng(W * 131 + d * 7 + n)
    x = rand_cf32(rng, n)
    f = ac.AutoCorrelator(W, d)
    got = f.execute_block(x)
    ref = O.autocorr_fast(W, d, x)
    assert len(got) == n
    if W > d:
        assert nerr(got, ref) <= TOL
    else:
        assert np.all(got == 0)
    o = O.AutoCorrelator(W, d)
    o.write(x[-min(n, 3 * W):])  # enough to fill both windows and the energy ring
    if n >= 3 * W:
        assert abs(f.get_energy() - o.get_energy()) <= 1e-5 * max(o.get_energy(), 1e-30)
        assert abs(f.execute() - o.execute()) <= 1e-5 * max(abs(o.execute()), np.max(np.abs(ref)), 1e-30)


def test_streaming_channels_state(ac):
    rng = np.random.default_rng(5)
    W, d, Cn, n = 48, 16, 5, 9000
    x = rand_cf32(rng, (Cn, n))
    whole = ac.AutoCorrelator(W, d, n_channels=Cn).execute_block(x)
    f = ac.AutoCorrelator(W, d, n_channels=Cn)
    cuts = [0, 1, 40, 47, 48, 2048, 2049, 6000, n]
    parts = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(whole, parts)
    for c in range(Cn):
        assert nerr(whole[c], O.autocorr_fast(W, d, x[c])) <= TOL
    # write() pushes without output; push() one sample; state round trip; clone; reset
    g = ac.AutoCorrelator(W, d, n_channels=Cn)
    g.write(x[:, :100])
    g.push(x[:, 100])
    st = g.get_state()
    assert st.shape == (Cn, W) and np.array_equal(st, x[:, 101 - W:101])
    h = g.clone()
    k = ac.AutoCorrelator(W, d, n_channels=Cn)
    k.set_state(st)
    a, b, c_ = (q.execute_block(x[:, 101:]) for q in (g, h, k))
    assert np.array_equal(a, b) and np.array_equal(a, whole[:, 101:])
    # a fresh handle primed with set_state sits at another absolute position: same values to rounding
    assert nerr(c_, a.astype(np.complex128)) <= 1e-6
    e = g.get_energy()
    for c in range(Cn):
        assert abs(e[c] - np.sum(np.abs(x[c, -W:].astype(np.complex128)) ** 2)) <= 1e-5 * e[c]
    g.reset()
    assert np.all(g.get_energy() == 0) and np.array_equal(g.execute_block(x), whole)


def test_device_pointers_and_errors(ac):
    import torch
    from solid_dsp_b200 import _ffi
    rng = np.random.default_rng(6)
    x = rand_cf32(rng, (3, 5000))
    f = ac.AutoCorrelator(20, 4, n_channels=3)
    got = f.execute_block(torch.from_numpy(x).cuda())
    assert got.is_cuda
    for c in range(3):
        assert nerr(got[c].cpu().numpy(), O.autocorr_fast(20, 4, x[c])) <= TOL
    with pytest.raises(_ffi.SolidGpuError):
        ac.AutoCorrelator(0, 0)                              # Window::new asserts capacity > 0


def test_more_than_65535_channels(ac):
    rng = np.random.default_rng(70000)
    C, n = 66000, 200
    x = rand_cf32(rng, (C, n))
    f = ac.AutoCorrelator(12, 4, n_channels=C)
    y = np.concatenate([f.execute_block(x[:, :77]), f.execute_block(x[:, 77:])], axis=1)
    for c in (0, 65534, 65535, 65536, C - 1):
        assert nerr(y[c], O.autocorr_fast(12, 4, x[c])) <= TOL
    e = f.get_energy()
    assert abs(e[C - 1] - np.sum(np.abs(x[C - 1, -12:].astype(np.complex128)) ** 2)) <= 1e-5 * e[C - 1]
