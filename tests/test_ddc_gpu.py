"""GPU parity of the NCO bank (csrc/nco.cu) and of the digital down-converter (sgpu_ddc_*: NCO mix-down fused into the
decimating FIR's tile loader) against the oracle's restatement of nco/mod.rs + filter/fir/decim.rs."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


@pytest.fixture(scope="module")
def NCO():
    from solid_dsp_b200.nco import NCO
    return NCO


@pytest.fixture(scope="module")
def DDC():
    from solid_dsp_b200.filter.ddc import DigitalDownConverter
    return DigitalDownConverter


def test_nco_scalar_accessors_match_the_oracle(NCO):
    n, o = NCO(), O.NCO()
    n.set_frequency(0.1)
    o.set_frequency(0.1)
    for _ in range(2000):
        assert n.raw() == o.raw()
        assert n.sincos() == o.sincos()
        n.step()
        o.step()
    n.adjust_phase(-1.0)
    o.adjust_phase(-1.0)
    n.adjust_frequency(6.0)
    o.adjust_frequency(6.0)
    assert n.raw() == o.raw()
    assert n.mix_down(0.5 - 2j) == o.mix_down(0.5 - 2j)
    assert n.get_frequency() == 0.0 and n.get_phase() == 0.0  # the reference's integer division (nco/mod.rs:69-91)


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("device", [False, True])
def test_nco_mix_block(NCO, torch, up, device):
    rng = np.random.default_rng(11)
    Cn, n = 3, 50001
    x = rand_cf32(rng, (Cn, n))
    raw = [(0, O.nco_constrain(0.1)), (0xFFFF0000, 0x7FFFFFFF), (123456789, O.nco_constrain(-2.5))]
    g = NCO(Cn)
    for c, (th, dl) in enumerate(raw):
        g.set_raw(th, dl, channel=c)
    fn = g.mix_up_block if up else g.mix_down_block
    cuts = [0, 1, 4098, 30000, n]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        xi = torch.from_numpy(x[:, a:b]).cuda() if device else x[:, a:b]
        yi = fn(xi)
        parts.append(yi.cpu().numpy() if device else yi)
    y = np.concatenate(parts, axis=1)
    ref = O.nco_mix_down_block(x, raw=raw, up=up)
    for c in range(Cn):
        assert nerr(y[c], ref[c]) <= 2e-7  # f32 table and products
        th, dl = g.raw(c)
        assert (th, dl) == ((raw[c][0] + n * raw[c][1]) & 0xFFFFFFFF, raw[c][1])


@pytest.mark.parametrize("M,T,fused", [(8, 256, True), (4, 37, True), (2, 64, True), (8, 16, True), (1, 40, False),
                                       (3, 50, False), (16, 128, False), (64, 256, False)])
def test_ddc_matches_mix_then_decimate(DDC, torch, M, T, fused):
    rng = np.random.default_rng(100 * M + T)
    Cn, n = 4, 70003
    h = f32_taps(O.firdes_kaiser(T, 0.45 / max(M, 1), 60.0, 0.0))
    x = rand_cf32(rng, (Cn, n))
    d = DDC(h, 0.75, M, frequency=0.1234, n_channels=Cn)
    d.nco.set_phase(1.0, channel=2)
    d.nco.set_frequency(-2.9, channel=3)
    raw = [d.nco.raw(c) for c in range(Cn)]
    cuts = [0, 5, 5 + M - 1 if M > 1 else 6, 20011, n]  # includes calls shorter than one decimation period
    outs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        y = d.execute_block(torch.from_numpy(x[:, a:b]).cuda())
        if b - a >= 4096:
            assert d.last_fused == fused
        outs.append(y.cpu().numpy())
    y = np.concatenate(outs, axis=1)
    for c in range(Cn):
        ref = O.ddc_fast(h, x[c], 0.75, M, raw=raw[c])
        assert y[c].shape == ref.shape
        assert nerr(y[c], ref) <= TOL
    # host pointers (chunked pipeline inside the library): same results
    d2 = DDC(h, 0.75, M, frequency=0.1234, n_channels=Cn)
    d2.nco.set_phase(1.0, channel=2)
    d2.nco.set_frequency(-2.9, channel=3)
    yh = d2.execute_block(x)
    assert nerr(yh, y) <= 1e-6


def test_ddc_complex_taps_write_clone_reset(DDC, torch):
    rng = np.random.default_rng(77)
    T, M, n = 48, 4, 30000
    h = f32_taps(rng.uniform(-1, 1, T)) + 1j * f32_taps(rng.uniform(-1, 1, T))
    x = rand_cf32(rng, (1, n))
    d = DDC(h, 1.0 - 0.25j, M, frequency=0.7)
    raw = d.nco.raw(0)
    xd = torch.from_numpy(x).cuda()
    d.write(xd[:, :1001])                      # mixed and pushed, no output; NCO and counter advance
    c = d.clone()
    y = d.execute_block(xd[:, 1001:]).cpu().numpy()
    ref = O.ddc_fast(h, x[0], 1.0 - 0.25j, M, raw=raw)
    n_skipped = 1001 // M
    assert nerr(y[0], ref[n_skipped:]) <= TOL
    assert np.array_equal(c.execute_block(xd[:, 1001:]).cpu().numpy(), y)
    assert not d.last_fused
    d.reset()
    assert d.nco.raw(0) == (0, 0)
    y0 = d.execute_block(xd[:, :4000]).cpu().numpy()
    assert nerr(y0[0], O.fir_fast(h, x[0, :4000], 1.0 - 0.25j, M)) <= TOL  # zero frequency: the plain decimator


def test_ddc_frequency_change_mid_stream_and_filter_view(DDC, torch):
    rng = np.random.default_rng(78)
    T, M, n = 256, 8, 1 << 16
    h = f32_taps(O.firdes_kaiser(T, 0.05, 80.0, 0.0))
    x = rand_cf32(rng, (2, n))
    d = DDC(h, 1.0, M, frequency=0.3, n_channels=2)
    assert d.get_decimation() == M and d.len() == T and d.filter.get_scale() == 1.0
    xd = torch.from_numpy(x).cuda()
    ya = d.execute_block(xd[:, : n // 2]).cpu().numpy()
    th = [d.nco.raw(c)[0] for c in range(2)]
    d.nco.set_frequency(-0.9, channel=1)      # phase continues, step changes
    yb = d.execute_block(xd[:, n // 2:]).cpu().numpy()
    assert d.last_fused
    for c in range(2):
        m1 = O.nco_mix_down_block(x[c, : n // 2], 0.3)
        dl = O.nco_constrain(-0.9) if c == 1 else O.nco_constrain(0.3)
        m2 = O.nco_mix_down_block(x[c, n // 2:], raw=(th[c], dl))
        ref = O.fir_fast(h, np.concatenate([m1, m2]), 1.0, M)
        assert nerr(np.concatenate([ya[c], yb[c]]), ref) <= TOL
    hist, cur = d.filter.get_state()
    assert hist.shape == (2, T - 1) and cur == 0
    assert nerr(hist[1], m2[-(T - 1):]) <= 2e-7  # the decimator's window holds MIXED samples
