"""Host-side API types kept from the reference: Window, CircularBuffer, design helpers, analysis
methods.  CPU only; checked against the oracle / the reference's doc-test values."""
import numpy as np
import pytest

import oracle as O
from solid_dsp_b200.circular_buffer import BufferError, BufferErrorCode, CircularBuffer
from solid_dsp_b200.filter import firdes, group_delay, iirdes
from solid_dsp_b200.window import Window


def test_window_shift_register():
    w = Window(4, 0, np.float64)
    w.write([1.0, 2.0, 3.0])
    assert list(w.to_vec()) == [3.0, 2.0, 1.0, 0.0]  # newest at index 0 (window/mod.rs:63-71)
    w.write([4.0, 5.0])
    assert list(w.to_vec()) == [5.0, 4.0, 3.0, 2.0]
    assert w.capacity() == 4
    assert list(w.to_history(3)) == [3.0, 4.0, 5.0]
    c = w.clone()
    w.reset()
    assert list(w.to_vec()) == [0.0] * 4 and list(c.to_vec()) == [5.0, 4.0, 3.0, 2.0]
    assert list(Window.from_history([7.0, 8.0, 9.0]).to_vec()) == [9.0, 8.0, 7.0]
    with pytest.raises(AssertionError):
        Window(0)


def test_circular_buffer_reference_behaviour():
    # doc-tests of circular_buffer/mod.rs:425-431, 461-467, 504-510, 536-545
    b = CircularBuffer(4, np.uint8)
    b.append([2, 3, 4, 5])
    with pytest.raises(BufferError) as e:
        b.append([6])
    assert e.value.code == BufferErrorCode.NotEnoughBuffer
    with pytest.raises(BufferError) as e:
        b.push(1)
    assert e.value.code == BufferErrorCode.FullBuffer
    assert b.pop() == 2 and b.len() == 3 and b.read_index() == 1
    b.push(9)
    assert b.write_index() == 1 and b.is_full()
    assert list(b.to_vec()) == [3, 4, 5, 9]
    b.release(2)
    assert b.len() == 2 and b.reserved() == 2
    with pytest.raises(BufferError) as e:
        b.release(-1)
    assert e.value.code == BufferErrorCode.NegativeBuffer
    with pytest.raises(BufferError) as e:
        b.release(5)
    assert e.value.code == BufferErrorCode.NotEnoughBuffer
    b.reset()
    assert b.is_empty()
    with pytest.raises(BufferError) as e:
        b.pop()
    assert e.value.code == BufferErrorCode.EmptyBuffer
    c = CircularBuffer.from_vec(np.array([1, 2, 3], dtype=np.int32))
    assert c.capacity() == 3 and c.is_full() and list(c.deref()) == [1, 2, 3]
    c.pop()
    c.linearize()
    assert c.read_index() == 0 and list(c.as_ptr()) == [2, 3, 1]


def test_design_helpers_match_oracle_bitwise():
    for args in [(64, 0.25, 60.0, 0.0), (512, 0.1, 80.0, 0.0), (256, 0.5 / 8 * 0.9, 80.0, 0.0),
                 (128, 0.5 / 4 * 0.9, 80.0, 0.0), (51, 0.35, 120.0, 0.0), (8, 0.35, 120.0, 0.25)]:
        assert np.array_equal(np.array(firdes.firdes_kaiser(*args)), O.firdes_kaiser(*args))
    assert np.array_equal(np.array(firdes.firdes_notch(25, 0.2, 30.0)), O.firdes_notch(25, 0.2, 30.0))
    assert firdes.kaiser_beta(60.0) == O.kaiser_beta(60.0) and firdes.kaiser_beta(30.0) == O.kaiser_beta(30.0)
    n, d = iirdes.active_lag(0.02, 1.0 / 2 ** 0.5, 1000.0)
    on, od = O.pll_active_lag(0.02, 1.0 / 2 ** 0.5, 1000.0)
    assert list(n) == list(on) and list(d) == list(od)


def test_design_goldens(golden):
    ref = golden["reference_doctests"]
    g = ref["firdes_autocorrelation"]
    taps = firdes.firdes_notch(*g["notch"])
    assert np.float32(firdes.filter_autocorrelation(taps, g["lag"])) == np.float32(g["expect"])
    g = ref["firdes_crosscorrelation"]
    v = firdes.filter_crosscorrelation(firdes.firdes_kaiser(*g["kaiser"]), firdes.firdes_notch(*g["notch"]), g["lag"])
    assert np.float32(v) == np.float32(g["expect"])
    g = ref["firdes_isi"]
    rms, mx = firdes.filter_isi(firdes.firdes_notch(*g["notch"]), g["samples_per_symbol"], g["filter_delay"])
    assert (np.float32(rms), np.float32(mx)) == (np.float32(g["expect"][0]), np.float32(g["expect"][1]))
    assert (rms, mx) == O.filter_isi(O.firdes_notch(*g["notch"]), g["samples_per_symbol"], g["filter_delay"])
    g = ref["firdes_energy"]
    e = firdes.filter_energy(firdes.firdes_notch(*g["notch"]), g["cutoff"], g["fft_size"])
    assert np.float32(e) == np.float32(g["expect"]) and e == O.filter_energy(O.firdes_notch(*g["notch"]), g["cutoff"], g["fft_size"])
    for args, code in (((0.6, 128), "Bandwidth"), ((0.35, 0), "FFTSize")):
        with pytest.raises(firdes.FirdesError) as err:
            firdes.filter_energy(firdes.firdes_notch(*g["notch"]), *args)
        assert str(err.value) == code
    with pytest.raises(firdes.FirdesError) as err:
        firdes.filter_energy([], 0.35, 128)
    assert str(err.value) == "FilterSize"
    with pytest.raises(firdes.FirdesError):
        firdes.firdes_kaiser(8, 0.7, 60.0)
    with pytest.raises(iirdes.IirdesError):
        iirdes.active_lag(0.0, 1.0, 1.0)


def test_group_delay_goldens():
    # iir/sos.rs:206 : SecondOrderFilter::group_delay(0.0) == 17.6774211296624 for active_lag(...)
    from solid_dsp_b200.filter.iir import SecondOrderFilter
    ff, fb = iirdes.active_lag(0.02, 1.0 / 2 ** 0.5, 1000.0)
    s = SecondOrderFilter(ff, fb)
    assert s.group_delay(0.0) == pytest.approx(17.6774211296624, rel=1e-12)
    assert list(s.numerator_coefs())[1] == 0.99999840000128          # sos.rs:129
    assert list(s.denominator_coefs())[1] == 0.003199997440002048    # sos.rs:149
    # fir/mod.rs:291 : (delay + 0.5) as usize == 12 for a 25-tap symmetric filter
    taps = firdes.firdes_notch(12, 0.35, 120.0)
    assert int(group_delay.fir_group_delay(taps, 0.0) + 0.5) == 12
    with pytest.raises(group_delay.DelayError):
        group_delay.fir_group_delay([], 0.0)
    with pytest.raises(group_delay.DelayError):
        group_delay.fir_group_delay([1.0], 0.7)


def test_stable_section_table():
    ff, fb = iirdes.stable_lowpass_sections(8)
    assert len(ff) == 24 and len(fb) == 24
    for i in range(8):
        a1, a2 = fb[3 * i + 1], fb[3 * i + 2]
        r = np.abs(np.roots([1.0, a1, a2]))
        assert np.all(r <= 0.951) and fb[3 * i] == 1.0
        assert abs(sum(ff[3 * i:3 * i + 3]) / (1 + a1 + a2) - 1.0) < 1e-6  # unit DC gain
    assert np.array_equal(ff.astype(np.float32).astype(np.float64), ff)  # f32-representable
