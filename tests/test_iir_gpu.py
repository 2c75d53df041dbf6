"""GPU parity: IIRFilter (SecondOrder batch + long-stream scan, Normal), the decimating and
interpolating wrappers, SecondOrderFilter and DotProduct, versus the CPU oracle."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def iir():
    from solid_dsp_b200.filter import iir
    return iir


def _sections(n):
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    return stable_lowpass_sections(n)


def _cx(pairs):
    return np.array([complex(a, b) for a, b in pairs])


def test_reference_goldens(iir, golden):
    """The reference's IIR goldens use pll::active_lag (double pole at z ~ 1): fine over 5 samples."""
    ref = golden["reference_doctests"]
    ff, fb = O.pll_active_lag(*ref["iir_sos_execute_block"]["design"]["active_lag"])
    g = ref["iir_sos_execute_block"]
    out = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder).execute_block(g["input"])
    assert nerr(out, g["expect"]) <= 1e-4  # a1 ~ -2, a2 ~ 1 rounded to f32: conditioning, see DESIGN.md
    out = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder).execute(1.0)
    assert abs(out[0] - 0.05816769596076701) <= 1e-5 * 0.0582
    s = iir.SecondOrderFilter(ff, fb)
    assert abs(s.execute(1.0) - 0.05816769596076701) <= 1e-5 * 0.0582
    g = ref["iir_decim_execute_block"]
    out = iir.DecimatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, 2).execute_block(g["input"])
    assert len(out) == 2 and nerr(out, g["expect"]) <= 1e-4
    f = iir.DecimatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, 2)
    assert len(f.execute(0.0)) == 0 and len(f.execute(1.0)) == 1
    g = ref["iir_interp_execute"]
    out = iir.InterpolatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, 2).execute(1.0)
    assert len(out) == 2 and nerr(out, g["expect"]) <= 1e-4
    g = ref["iir_interp_execute_block_len"]
    out = iir.InterpolatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, 5).execute_block(g["input"])
    assert len(out) == g["expect_len"]


def test_derived_vectors(iir, golden):
    der = golden["derived_vectors"]
    x = _cx(der["x"])
    g = der["iir_sos_2sections"]
    assert nerr(iir.IIRFilter(g["ff"], g["fb"], iir.IIRFilterType.SecondOrder).execute_block(x),
                _cx(g["expect"])) <= TOL
    g = der["iir_normal_vs_sos"]
    assert nerr(iir.IIRFilter(g["b"], g["a"], iir.IIRFilterType.Normal).execute_block(x), _cx(g["expect"])) <= TOL
    assert nerr(iir.IIRFilter(g["b"], g["a"], iir.IIRFilterType.SecondOrder).execute_block(x), _cx(g["expect"])) <= TOL


@pytest.mark.parametrize("nsec", [1, 2, 3, 5, 8, 11, 16])
@pytest.mark.parametrize("C,n", [(1, 1), (1, 100), (3, 1000), (32, 16), (33, 47), (70, 1025), (256, 512)])
def test_sos_batch_random(iir, nsec, C, n):
    rng = np.random.default_rng(100 * nsec + C + n)
    ff, fb = _sections(nsec)
    x = rand_cf32(rng, (C, n))
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=C)
    f.set_mode(0)
    got = f.execute_block(x)
    assert got.shape == (C, n)
    for c in sorted({0, C // 2, C - 1}):
        ref, _ = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(got[c], ref) <= TOL


def test_sos_unnormalised_a0(iir):
    """fb[3i] != 1: every section is normalised by its own a0 (sos.rs:62-68)."""
    rng = np.random.default_rng(2)
    ff = f32_taps([0.2, 0.4, 0.2, 0.5, 0, -0.5])
    fb = f32_taps([1, -0.5, 0.25, 2, 0.6, 0.2])
    x = rand_cf32(rng, (4, 777))
    got = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=4).execute_block(x)
    for c in range(4):
        assert nerr(got[c], O.IIRFilter(ff, fb, O.SECOND_ORDER).execute_block(x[c])) <= TOL


def test_sos_streaming_state_clone(iir):
    rng = np.random.default_rng(4)
    ff, fb = _sections(8)
    x = rand_cf32(rng, (40, 3000))
    whole = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=40).execute_block(x)
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=40)
    cuts = [0, 1, 17, 1000, 1016, 3000]
    parts = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(whole, parts)
    g = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=40)
    g.execute_block(x[:, :1000])
    st, idx = g.get_state()
    assert st.shape == (40, 16) and idx == 0
    _, ost = O.sos_cascade_fast(ff, fb, x[5, :1000])
    assert nerr(st[5], ost.ravel()) <= 1e-4
    h = g.clone()
    a = g.execute_block(x[:, 1000:])
    b = h.execute_block(x[:, 1000:])
    assert np.array_equal(a, b) and np.array_equal(a, whole[:, 1000:])
    k = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=40)
    k.set_state(st)  # states cross the ABI in the reference's scaling (f32): exact to rounding, not bit-exact
    assert nerr(k.execute_block(x[:, 1000:]), a) <= 1e-6
    k.reset()
    assert np.array_equal(k.execute_block(x), whole)
    assert np.array_equal(f.numerator_coefs(), ff) and np.array_equal(f.denominator_coefs(), fb)
    assert len(f.second_order_filters()) == 8 and f.iir_type() == iir.IIRFilterType.SecondOrder


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("C,n", [(1, 1 << 16), (1, 100003), (2, 70000), (5, 40000), (40, 20000)])
def test_sos_long_stream_scan(iir, C, n, mode):
    """Chunked scans == sequential recurrence, incl. ragged last chunk.  mode 1: fused warm-up scan
    (chunks of one channel side by side for C < 32, channels side by side for C >= 32); mode 2:
    three-pass scan (zero-state pass / f64 carry / output pass)."""
    rng = np.random.default_rng(C * 7 + n)
    ff, fb = _sections(8)
    x = rand_cf32(rng, (C, n))
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=C)
    f.set_mode(mode)
    got = f.execute_block(x)
    b = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=C)
    b.set_mode(0)
    batch = b.execute_block(x)
    for c in range(C):
        ref, ost = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(got[c], ref) <= TOL
        assert nerr(batch[c], ref) <= TOL
    # streaming across scan calls: state carried in and out
    f2 = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=C)
    f2.set_mode(mode)
    h = n // 2 + 3
    two = np.concatenate([f2.execute_block(x[:, :h]), f2.execute_block(x[:, h:])], axis=1)
    for c in range(C):
        ref, _ = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(two[c], ref) <= TOL
    st, _ = f.get_state()
    st2, _ = f2.get_state()
    assert nerr(st, st2) <= 1e-4


@pytest.mark.parametrize("n", [1 << 16, 100003, 65536 + 255, 1 << 21])
def test_scan_fast_decay(iir, n):
    """Pole radius <= 0.6: the filter's memory is ~64 samples, so the fused scan runs many short
    chunks, each warming up over the 64 samples in front of it.  Ragged last chunk included."""
    rng = np.random.default_rng(n)
    ff, fb = _sections(2)
    x = rand_cf32(rng, (3, n))
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=3)
    f.set_mode(1)
    got = np.concatenate([f.execute_block(x[:, :n // 3]), f.execute_block(x[:, n // 3:])], axis=1)
    for c in range(3):
        ref, ost = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(got[c], ref) <= TOL
    st, _ = f.get_state()
    assert nerr(st[2], ost.ravel()) <= 1e-4


@pytest.mark.parametrize("mode", [1, 2])
def test_scan_marginal_poles(iir, mode):
    """Pole radius 0.999: A^Lc is far from zero, so the carry propagation really matters (mode 2);
    the fused scan (mode 1) needs a 23 k-sample warm-up here."""
    rng = np.random.default_rng(12)
    r, th = 0.999, 0.05
    fb = f32_taps([1.0, -2 * r * np.cos(np.pi * th), r * r])
    ff = f32_taps([1e-3, 2e-3, 1e-3])
    x = rand_cf32(rng, 1 << 16)
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder)
    f.set_mode(mode)
    ref, _ = O.sos_cascade_fast(ff, fb, x)
    assert nerr(f.execute_block(x), ref) <= 5e-5  # conditioning of the f32 recurrence itself


@pytest.mark.parametrize("M", [1, 2, 3, 7])
def test_iir_decimating(iir, M):
    rng = np.random.default_rng(M)
    ff, fb = _sections(4)
    x = rand_cf32(rng, (35, 1000))
    f = iir.DecimatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, M, n_channels=35)
    cuts = [0, 5, 6, 500, 1000]
    parts = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert parts.shape == (35, 1000 // M) and f.get_decimation() == M
    for c in (0, 34):
        ref = O.DecimatingIIRFilter(ff, fb, O.SECOND_ORDER, M).execute_block(x[c])
        assert nerr(parts[c], ref) <= TOL


@pytest.mark.parametrize("L", [1, 2, 3, 5, 6, 7, 12, 31, 32, 33, 40])
def test_iir_interpolating(iir, L):
    rng = np.random.default_rng(L)
    ff, fb = _sections(3)
    x = rand_cf32(rng, (33, 300))
    f = iir.InterpolatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, L, n_channels=33)
    got = np.concatenate([f.execute_block(x[:, :7]), f.execute_block(x[:, 7:])], axis=1)
    assert got.shape == (33, 300 * L) and f.get_interpolation() == L
    for c in (0, 32):
        ref = O.InterpolatingIIRFilter(ff, fb, O.SECOND_ORDER, L).execute_block(x[c])
        assert nerr(got[c], ref) <= TOL
    # exact positions: an impulse at input 3 comes out of the first section chain at output 3 L and nowhere before
    # (iir/interp.rs:184-190: the input, then L - 1 zeros); every factor, tile path (L <= 32) and guarded path
    imp = np.zeros((33, 50), dtype=np.complex64)
    imp[:, 3] = 1.0
    y = iir.InterpolatingIIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, L, n_channels=33).execute_block(imp)
    assert not np.any(y[:, :3 * L]) and np.all(y[:, 3 * L] != 0)
    ref = O.InterpolatingIIRFilter(ff, fb, O.SECOND_ORDER, L).execute_block(imp[0])
    assert nerr(y[17], ref) <= TOL


def test_iir_normal_mode(iir):
    rng = np.random.default_rng(9)
    b = f32_taps([0.1, 0.3, 0.2, 0.05, 0.01])
    a = f32_taps([1.0, -0.6, 0.3, -0.05])
    x = rand_cf32(rng, (6, 500))
    f = iir.IIRFilter(b, a, iir.IIRFilterType.Normal, n_channels=6)
    got = np.concatenate([f.execute_block(x[:, :123]), f.execute_block(x[:, 123:])], axis=1)
    for c in range(6):
        assert nerr(got[c], O.IIRFilter(b, a, O.NORMAL).execute_block(x[c])) <= TOL
    d = iir.DecimatingIIRFilter(b, a, iir.IIRFilterType.Normal, 3)
    assert nerr(d.execute_block(x[0]), O.DecimatingIIRFilter(b, a, O.NORMAL, 3).execute_block(x[0])) <= TOL
    i = iir.InterpolatingIIRFilter(b, a, iir.IIRFilterType.Normal, 2)
    assert nerr(i.execute_block(x[0]), O.InterpolatingIIRFilter(b, a, O.NORMAL, 2).execute_block(x[0])) <= TOL
    assert np.allclose(f.numerator_coefs(), b) and np.allclose(f.denominator_coefs(), a[1:])


def test_iir_errors(iir):
    T = iir.IIRFilterType
    for ctor, code in [
        (lambda: iir.IIRFilter([1.0] * 3, [1.0] * 6, T.SecondOrder), "SecondOrderSectionSizeMismatch"),
        (lambda: iir.IIRFilter([], [], T.SecondOrder), "SecondOrderSectionSizeZero"),
        (lambda: iir.IIRFilter([1.0] * 4, [1.0] * 4, T.SecondOrder), "SecondOrderSectionSizeNotMultpleOf3"),
        (lambda: iir.IIRFilter([], [1.0], T.Normal), "NumeratorLengthZero"),
        (lambda: iir.IIRFilter([1.0], [], T.Normal), "DenominatorLengthZero"),
        (lambda: iir.DecimatingIIRFilter([1.0] * 3, [1.0] * 3, T.SecondOrder, 0), "DecimationLessThanOne"),
        (lambda: iir.InterpolatingIIRFilter([1.0] * 3, [1.0] * 3, T.SecondOrder, 0), "InterpolationLessThanOne"),
    ]:
        with pytest.raises(iir.IIRError) as e:
            ctor()
        assert e.value.code == code
    with pytest.raises(iir.SecondOrderError):
        iir.SecondOrderFilter([1.0, 2.0], [1.0, 2.0, 3.0])


def test_dot_product(golden):
    from solid_dsp_b200.dot_product import Direction, DotProduct
    g = golden["reference_doctests"]["dot_product_reverse"]
    dp = DotProduct(g["coefs"], Direction.REVERSE)
    assert dp.execute(g["samples"]) == 15.0 and dp.len() == 5 and not dp.is_empty()
    assert list(dp.coefficents()) == g["coefs"][::-1]
    assert list(DotProduct(g["coefs"], Direction.FORWARD).coefficents()) == g["coefs"]
    rng = np.random.default_rng(3)
    for n_c, n_x in [(1, 1), (7, 5), (5, 7), (512, 512), (3000, 4000)]:
        c = f32_taps(rng.uniform(-1, 1, n_c))
        x = rand_cf32(rng, n_x)
        for d, od in [(Direction.FORWARD, O.DotProduct.FORWARD), (Direction.REVERSE, O.DotProduct.REVERSE)]:
            ref = O.DotProduct(c, od).execute(x)
            got = DotProduct(c, d).execute(x)
            assert abs(got - ref) <= 2e-5 * max(1.0, np.sum(np.abs(c[:min(n_c, n_x)])))
    cc = (f32_taps(rng.uniform(-1, 1, 33)) + 1j * f32_taps(rng.uniform(-1, 1, 33)))
    xv = rand_cf32(rng, (4, 33))
    got = DotProduct(cc, Direction.FORWARD).execute(xv)
    for v in range(4):
        assert abs(got[v] - O.DotProduct(cc, O.DotProduct.FORWARD).execute(xv[v])) <= 1e-4


# ------------------------------------------------------------------ gain folding / strategy selection
def test_sos_sections_that_cannot_be_folded(iir):
    """A section with b0 == 0 (pure delay numerator) and one with a tiny b0 (running product leaves
    2^-40): the kernel must fall back to the 5-operation biquad; states still in reference scaling."""
    rng = np.random.default_rng(77)
    x = rand_cf32(rng, (33, 2000))
    for ff, fb in (([0.0, 1.0, 0.5, 0.3, 0.2, 0.1], [1.0, -0.5, 0.25, 1.0, 0.3, 0.1]),
                   ([1e-7, 2e-7, 1e-7, 1e-7, 2e-7, 1e-7, 0.2, 0.1, 0.3], [1.0, -1.2, 0.5, 1.0, -0.9, 0.4, 1.0, 0.1, 0.2])):
        ff, fb = f32_taps(ff), f32_taps(fb)
        f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=33)
        got = f.execute_block(x)
        for c in (0, 32):
            ref, ost = O.sos_cascade_fast(ff, fb, x[c])
            assert nerr(got[c], ref) <= TOL
        st, _ = f.get_state()
        assert nerr(st[32], ost.ravel()) <= 1e-4


def test_sos_folded_states_cross_the_abi_in_reference_scaling(iir):
    """get_state after a folded run == the oracle's (v1, v2); set_state of those values resumes."""
    rng = np.random.default_rng(78)
    ff, fb = _sections(8)
    x = rand_cf32(rng, (4, 1500))
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=4)
    f.execute_block(x[:, :700])
    st, _ = f.get_state()
    for c in range(4):
        _, ost = O.sos_cascade_fast(ff, fb, x[c, :700])
        assert nerr(st[c], ost.ravel()) <= 1e-4
    g = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=4)
    g.set_state(st)
    tail = g.execute_block(x[:, 700:])
    for c in range(4):
        ref, _ = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(tail[c], ref[700:]) <= TOL


@pytest.mark.parametrize("C,n", [(100, 30000), (8, 20000), (33, 4096), (1, 300000)])
def test_sos_auto_strategy(iir, C, n):
    """Default mode: few channels x long streams are cut into chunks (fused scan, both row layouts);
    results, split calls and final states match the sequential recurrence."""
    rng = np.random.default_rng(C + n)
    ff, fb = _sections(8)
    x = rand_cf32(rng, (C, n))
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder, n_channels=C)
    h = n // 3 + 5
    got = np.concatenate([f.execute_block(x[:, :h]), f.execute_block(x[:, h:])], axis=1)
    for c in sorted({0, C // 2, C - 1}):
        ref, ost = O.sos_cascade_fast(ff, fb, x[c])
        assert nerr(got[c], ref) <= TOL
    st, _ = f.get_state()
    assert nerr(st[C - 1], ost.ravel()) <= 1e-4


@pytest.mark.parametrize("W", [1, 2, 3, 5, 8, 9, 12])
def test_iir_normal_mode_orders(iir, W):
    """Normal mode over a range of windows: W <= 8 runs in the tile kernel (states in registers,
    coalesced rows), larger windows in the one-thread-per-channel kernel.  70 channels so that three
    warps (one of them partial) are busy; split calls; state in the reference's order (newest first)."""
    rng = np.random.default_rng(W)
    nb, na = W, max(1, W - 1)
    b = f32_taps(rng.uniform(-0.5, 0.5, nb))
    # stable denominator: small feedback taps
    a = f32_taps(np.concatenate([[1.0], rng.uniform(-0.9, 0.9, na - 1) / max(1, na - 1)]))
    x = rand_cf32(rng, (70, 3001))
    f = iir.IIRFilter(b, a, iir.IIRFilterType.Normal, n_channels=70)
    got = np.concatenate([f.execute_block(x[:, :1000]), f.execute_block(x[:, 1000:1033]),
                          f.execute_block(x[:, 1033:])], axis=1)
    for c in (0, 31, 32, 69):
        assert nerr(got[c], O.IIRFilter(b, a, O.NORMAL).execute_block(x[c])) <= TOL
    g = iir.IIRFilter(b, a, iir.IIRFilterType.Normal, n_channels=70)
    g.execute_block(x[:, :1000])
    h = g.clone()
    st, _ = g.get_state()
    k = iir.IIRFilter(b, a, iir.IIRFilterType.Normal, n_channels=70)
    k.set_state(st)
    y1, y2, y3 = (q.execute_block(x[:, 1000:]) for q in (g, h, k))
    assert np.array_equal(y1, y2) and np.array_equal(y1, y3) and np.array_equal(y1, got[:, 1000:])


def test_decay_length_and_stream_segments(iir):
    """sgpu_iir_decay_length + the segment recipe of solid_dsp_b200.sharding (multi-GPU long stream):
    segments warmed up over the previous decay_length samples reproduce the unbroken recurrence."""
    from solid_dsp_b200 import sharding
    rng = np.random.default_rng(99)
    ff, fb = _sections(8)
    f = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder)
    warm = f.decay_length()
    assert warm % 32 == 0 and 256 <= warm <= 1024           # pole radii <= 0.95: a few hundred samples
    r = 0.9999                                               # memory of ~2.3e5 samples: "does not decay"
    g = iir.IIRFilter(f32_taps([1e-3, 2e-3, 1e-3]), f32_taps([1.0, -2 * r * np.cos(0.1), r * r]),
                      iir.IIRFilterType.SecondOrder)
    assert g.decay_length() == 0
    assert iir.IIRFilter([0.5, 0.2], [1.0, -0.3], iir.IIRFilterType.Normal).decay_length() == 0
    n = 90_000
    x = rand_cf32(rng, n)
    ref, _ = O.sos_cascade_fast(ff, fb, x)
    world = 3
    for rank in range(world):
        first, count = sharding.shard_stream(n, 32, world, rank)
        halo = x[first - warm:first] if rank else np.zeros(0, dtype=np.complex64)
        seg = iir.IIRFilter(ff, fb, iir.IIRFilterType.SecondOrder)
        y = sharding.iir_segment(seg, x[first:first + count], halo, rank)
        assert nerr(y, ref[first:first + count]) <= TOL
