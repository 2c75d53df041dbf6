"""Pins the CPU oracle to every golden vector the reference's own doc-tests hold for the
filtering hot path (bit-exact f64, as the reference's assert_eq! demands), then to the
structurally derived vectors of SURVEY.md Appendix B."""
import numpy as np
import pytest

import oracle as O


def _cx(pairs):
    return np.array([complex(a, b) for a, b in pairs])


def _design(spec):
    w, z, k = spec["active_lag"]
    return O.pll_active_lag(w, z, k)


@pytest.fixture(scope="module")
def ref(golden):
    return golden["reference_doctests"]


@pytest.fixture(scope="module")
def der(golden):
    return golden["derived_vectors"]


def test_msb_index(ref):
    assert O.msb_index(ref["msb_index"]["value"]) == ref["msb_index"]["expect"]
    assert O.msb_index(1) == 1
    assert O.msb_index(64) == 7 and O.msb_index(512) == 10  # capacity 128 / 1024


def test_dot_product(ref):
    g = ref["dot_product_reverse"]
    dp = O.DotProduct(g["coefs"], O.DotProduct.REVERSE)
    assert dp.execute(g["samples"]) == complex(g["expect"], 0.0)
    assert dp.len() == 5 and not dp.is_empty()
    assert list(dp.coefficents()) == g["coefs"][::-1]  # stored (reversed) order
    g = ref["dot_product_forward_coefficents"]
    assert list(O.DotProduct(g["coefs"], O.DotProduct.FORWARD).coefficents()) == g["expect"]
    # min(len_c, len_x) terms -- dot_product/mod.rs:160
    assert O.DotProduct([1.0, 2.0, 3.0], O.DotProduct.FORWARD).execute([1.0, 1.0]) == 3.0


def test_fir_execute(ref):
    g = ref["fir_execute"]
    f = O.FIRFilter(g["coefs"], g["scale"])
    assert f.execute(g["input"][0])[0] == complex(g["expect_first"], 0.0)
    assert f.window_capacity() == 8  # 1 << msb_index(5)


def test_fir_execute_block(ref):
    g = ref["fir_execute_block"]
    out = O.FIRFilter(g["coefs"], g["scale"]).execute_block(g["input"])
    assert len(out) == len(g["input"])
    assert out[g["expect_index"]] == complex(g["expect_value"], 0.0)
    # closed form is bit-identical to the structural mirror
    assert np.array_equal(out, O.fir_fast(g["coefs"], g["input"], g["scale"]))


def test_fir_decim(ref):
    g = ref["fir_decim_execute"]
    f = O.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"])
    for x, exp in zip(g["input"], g["expect_per_call"]):
        assert list(f.execute(x)) == [complex(e, 0.0) for e in exp]
    g = ref["fir_decim_execute_block"]
    f = O.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"])
    out = f.execute_block(g["input"])
    assert list(out) == [complex(e, 0.0) for e in g["expect"]]
    assert np.array_equal(out, O.fir_fast(g["coefs"], g["input"], g["scale"], g["decimation"]))


def test_sos(ref):
    g = ref["sos_execute"]
    ff, fb = _design(g["design"])
    s = O.SecondOrderFilter(ff, fb)
    assert s.execute(g["input"][0]) == complex(g["expect"][0], 0.0)
    g = ref["sos_numerator_coefs"]
    n = O.SecondOrderFilter(*_design(g["design"])).numerator_coefs()
    assert len(n) == g["len"] and n[g["index"]] == g["expect"]
    g = ref["sos_denominator_coefs"]
    d = O.SecondOrderFilter(*_design(g["design"])).denominator_coefs()
    assert len(d) == g["len"] and d[g["index"]] == g["expect"]


def test_iir_sos(ref):
    g = ref["iir_sos_execute"]
    ff, fb = _design(g["design"])
    assert O.IIRFilter(ff, fb, O.SECOND_ORDER).execute(1.0)[0] == complex(g["expect"][0], 0)
    g = ref["iir_sos_execute_block"]
    out = O.IIRFilter(ff, fb, O.SECOND_ORDER).execute_block(g["input"])
    assert list(out) == [complex(e, 0.0) for e in g["expect"]]
    y, _ = O.sos_cascade_fast(ff, fb, g["input"])
    assert np.array_equal(out, y)


def test_iir_decim_interp(ref):
    g = ref["iir_decim_execute"]
    ff, fb = _design(g["design"])
    f = O.DecimatingIIRFilter(ff, fb, O.SECOND_ORDER, g["decimation"])
    for x, exp in zip(g["input"], g["expect_per_call"]):
        assert list(f.execute(x)) == [complex(e, 0.0) for e in exp]
    g = ref["iir_decim_execute_block"]
    f = O.DecimatingIIRFilter(ff, fb, O.SECOND_ORDER, g["decimation"])
    assert list(f.execute_block(g["input"])) == [complex(e, 0.0) for e in g["expect"]]
    g = ref["iir_interp_execute"]
    f = O.InterpolatingIIRFilter(ff, fb, O.SECOND_ORDER, g["interpolation"])
    assert list(f.execute(1.0)) == [complex(e, 0.0) for e in g["expect"]]
    g = ref["iir_interp_execute_block_len"]
    f = O.InterpolatingIIRFilter(ff, fb, O.SECOND_ORDER, g["interpolation"])
    assert len(f.execute_block(g["input"])) == g["expect_len"]


def test_firdes(ref):
    g = ref["firdes_autocorrelation"]
    taps = O.firdes_notch(*g["notch"])
    assert np.float32(O.filter_autocorrelation(taps, g["lag"])) == np.float32(g["expect"])
    assert O.filter_autocorrelation(taps, g["lag"]) == O.filter_autocorrelation(taps, -g["lag"])
    g = ref["firdes_crosscorrelation"]
    h = O.firdes_kaiser(*g["kaiser"])
    n = O.firdes_notch(*g["notch"])
    assert np.float32(O.filter_crosscorrelation(h, n, g["lag"])) == np.float32(g["expect"])
    # the two analysis routines behind the taps: filter_isi, and filter_energy -- the crate's own caller of
    # DotProduct::execute (firdes/mod.rs:620-629)
    g = ref["firdes_isi"]
    rms, mx = O.filter_isi(O.firdes_notch(*g["notch"]), g["samples_per_symbol"], g["filter_delay"])
    assert np.float32(rms) == np.float32(g["expect"][0]) and np.float32(mx) == np.float32(g["expect"][1])
    assert O.filter_isi(O.firdes_notch(*g["notch"]), 2, g["filter_delay"]) == (0.0, 0.0)   # length mismatch, :554-561
    g = ref["firdes_energy"]
    assert np.float32(O.filter_energy(O.firdes_notch(*g["notch"]), g["cutoff"], g["fft_size"])) == np.float32(g["expect"])
    for bad, code in (((0.6, 128), "Bandwidth"), ((0.35, 0), "FFTSize")):
        with pytest.raises(ValueError) as e:
            O.filter_energy(O.firdes_notch(*g["notch"]), *bad)
        assert str(e.value) == code
    assert len(O.firdes_kaiser(*ref["firdes_kaiser_len"]["kaiser"])) == 8
    assert len(O.firdes_notch(*ref["firdes_notch_len"]["notch"])) == 17


def test_construction_errors():
    for ctor, code in [
        (lambda: O.FIRFilter([], 1.0), "FIRErrorCode::CoefficientsLengthZero"),
        (lambda: O.DecimatingFIRFilter([1.0], 1.0, 0), "FIRErrorCode::DecimationLessThanOne"),
        (lambda: O.InterpolatingFIRFilter([1.0], 0), "FIRErrorCode::InterpolationLessThanOne"),
        (lambda: O.PolyPhaseFilterBank([1.0], 0), "FIRErrorCode::NotEnoughFilters"),
        (lambda: O.IIRFilter([1.0] * 3, [1.0] * 6, O.SECOND_ORDER),
         "IIRErrorCode::SecondOrderSectionSizeMismatch"),
        (lambda: O.IIRFilter([], [], O.SECOND_ORDER), "IIRErrorCode::SecondOrderSectionSizeZero"),
        (lambda: O.IIRFilter([1.0] * 4, [1.0] * 4, O.SECOND_ORDER),
         "IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3"),
        (lambda: O.IIRFilter([], [1.0], O.NORMAL), "IIRErrorCode::NumeratorLengthZero"),
        (lambda: O.IIRFilter([1.0], [], O.NORMAL), "IIRErrorCode::DenominatorLengthZero"),
        (lambda: O.DecimatingIIRFilter([1.0] * 3, [1.0] * 3, O.SECOND_ORDER, 0),
         "IIRErrorCode::DecimationLessThanOne"),
        (lambda: O.InterpolatingIIRFilter([1.0] * 3, [1.0] * 3, O.SECOND_ORDER, 0),
         "IIRErrorCode::InterpolationLessThanOne"),
        (lambda: O.SecondOrderFilter([1.0, 2.0], [1.0, 2.0, 3.0]),
         "SecondOrderErrorCode::CoefficientsNotInRange"),
    ]:
        with pytest.raises(O.OracleError) as e:
            ctor()
        assert e.value.code == code


# ------------------------------------------------------------------ derived (Appendix B)
def test_derived_fir(der):
    x = _cx(der["x"])
    for key in ("fir_123", "fir_123_scale_half"):
        g = der[key]
        out = O.FIRFilter(g["coefs"], g["scale"]).execute_block(x)
        assert np.array_equal(out, _cx(g["expect"]))
    g = der["decim_123_m2"]
    out = O.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"]).execute_block(x)
    assert np.array_equal(out, _cx(g["expect"]))


def test_derived_interp(der):
    x = _cx(der["x"])
    for key in ("interp_6taps_l2", "interp_5taps_l2_padded"):
        g = der[key]
        out = O.InterpolatingFIRFilter(g["coefs"], g["interpolation"]).execute_block(x[: g["n_in"]])
        assert np.array_equal(out, _cx(g["expect"]))
        assert np.array_equal(out, O.firinterp_fast(g["coefs"], g["interpolation"], x[: g["n_in"]]))
    g = der["interp_5taps_l4_impulse"]
    f = O.InterpolatingFIRFilter(g["coefs"], g["interpolation"])
    assert f.sub_len() == 2
    f.set_scale(7.0)  # stored, never applied -- pfb.rs:85-90
    assert np.array_equal(f.execute_block(_cx(g["input"])), _cx(g["expect"]))


def test_derived_iir(der):
    x = _cx(der["x"])
    g = der["iir_sos_2sections"]
    out = O.IIRFilter(g["ff"], g["fb"], O.SECOND_ORDER).execute_block(x)
    assert np.array_equal(out, _cx(g["expect"]))
    y, st = O.sos_cascade_fast(g["ff"], g["fb"], x)
    assert np.array_equal(out, y)
    g = der["iir_normal_vs_sos"]
    n = O.IIRFilter(g["b"], g["a"], O.NORMAL).execute_block(x)
    s = O.IIRFilter(g["b"], g["a"], O.SECOND_ORDER).execute_block(x)
    assert np.array_equal(n, _cx(g["expect"])) and np.array_equal(s, _cx(g["expect"]))


def test_streaming_split_calls():
    """concat(execute_block(a), execute_block(b)) == execute_block(a ++ b) for every type."""
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, 97) + 1j * rng.uniform(-1, 1, 97)
    h = rng.uniform(-1, 1, 13)
    cuts = [0, 10, 11, 40, 97]
    mk = [
        lambda: O.FIRFilter(h, 0.7),
        lambda: O.DecimatingFIRFilter(h, 1.3, 4),
        lambda: O.InterpolatingFIRFilter(h, 3),
        lambda: O.IIRFilter([0.2, 0.4, 0.2, 0.5, 0, -0.5], [1, -0.5, 0.25, 2, 0.6, 0.2], O.SECOND_ORDER),
        lambda: O.DecimatingIIRFilter([0.2, 0.4, 0.2], [1, -0.5, 0.25], O.SECOND_ORDER, 3),
    ]
    for make in mk:
        whole = make().execute_block(x)
        f = make()
        parts = np.concatenate([f.execute_block(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        assert np.array_equal(whole, parts)


def test_fast_equals_structural_random():
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, 300) + 1j * rng.uniform(-1, 1, 300)
    for T in (1, 2, 7, 64):
        h = rng.uniform(-1, 1, T)
        assert np.array_equal(O.FIRFilter(h, 0.5).execute_block(x), O.fir_fast(h, x, 0.5))
        hc = h + 1j * rng.uniform(-1, 1, T)
        assert np.array_equal(O.FIRFilter(hc, 0.5 + 0.25j).execute_block(x),
                              O.fir_fast(hc, x, 0.5 + 0.25j))
        for M in (1, 3, 8):
            assert np.array_equal(O.DecimatingFIRFilter(h, 2.0, M).execute_block(x),
                                  O.fir_fast(h, x, 2.0, M))
        for L in (1, 2, 4, 5):
            assert np.array_equal(O.InterpolatingFIRFilter(h, L).execute_block(x),
                                  O.firinterp_fast(h, L, x))
    # decimator write() advances the counter without output -- fir/decim.rs:136-139
    f = O.DecimatingFIRFilter(h, 1.0, 4)
    f.write(x[:6])
    assert f.current_item() == 2
    a = f.execute_block(x[6:])
    b = O.fir_fast(h, x[6:], 1.0, 4, count0=2, hist=np.concatenate([np.zeros(63 - 6), x[:6]])[-63:])
    assert np.array_equal(a, b)


def test_auto_correlator(ref):
    """auto_correlator/mod.rs:199-211 -- the reference's only numeric assertion for this type -- and
    the Window(capacity, delay) quirk that the restatement must carry: the delayed window's tail is
    never written (window/mod.rs:17-71), so only W-d lags contribute and delay >= window gives 0."""
    g = ref["auto_correlator_get_energy"]
    x = np.array([complex(np.cos(float(k)) * 0.05, np.sin(float(k)) * 0.05) for k in range(-250, 250)])
    a = O.AutoCorrelator(g["construct"]["window_size"], g["construct"]["delay"])
    out = a.execute_block(x)
    assert round(a.get_energy() * 10000.0) == g["expect_rounded"]
    assert np.all(out == 0)                                   # delay (10) >= window (5)
    rng = np.random.default_rng(3)
    y = rng.normal(size=200) + 1j * rng.normal(size=200)
    for W, d in [(1, 0), (7, 0), (10, 5), (16, 15), (16, 16)]:
        s = O.AutoCorrelator(W, d).execute_block(y)
        assert np.array_equal(s, O.autocorr_fast(W, d, y))   # closed form == structural, bit for bit
        n = 150                                                # hand check of one output
        want = sum(y[n - i] * np.conj(y[n - d - i]) for i in range(max(W - d, 0)))
        assert abs(s[n] - want) <= 1e-12 * max(1.0, abs(want))
    # split calls == one call (the windows carry the state)
    b = O.AutoCorrelator(10, 5)
    two = np.concatenate([b.execute_block(y[:77]), b.execute_block(y[77:])])
    assert np.array_equal(two, O.AutoCorrelator(10, 5).execute_block(y))
