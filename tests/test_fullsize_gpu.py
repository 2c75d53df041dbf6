"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to walk the
whole buffer: impulse responses at large indices (exact sample counts, phase alignment, 64-bit index
arithmetic), linearity, and oracle windows whose history is regenerated from the same inputs.

Sizes: config 2 = one 512-tap FIR stream of 2^30 samples (8 GiB in, 8 GiB out); config 3 = decimator
M=8 / 256 taps on 4096 channels x 2^20 (32 GiB in); config 4 = interpolator L=4 / 128 taps on 1024
channels x 2^20 (32 GiB out); config 5 = 8 biquads on 65536 channels x 2^14 and on one stream of 2^28."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if torch.cuda.mem_get_info()[1] < 100 * (1 << 30):
        pytest.skip("full-size buffers need a 180 GB B200")
    yield torch
    torch.cuda.empty_cache()


def _window(t, lo, hi):
    return t[lo:hi].cpu().numpy()


def test_config2_fir_full_stream(torch_cuda):
    torch = torch_cuda
    from solid_dsp_b200.filter.fir import FIRFilter
    h = f32_taps(O.firdes_kaiser(512, 0.1, 80.0, 0.0))
    n = 1 << 30
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    y = FIRFilter(h, 1.0).execute_block(x)
    assert y.shape == (n,)
    # oracle windows, also past 2^29 and at the very end (history regenerated from the same buffer)
    for start in (0, 511, (1 << 29) + 12345, (1 << 30) - 4096):
        lo = max(0, start - 511)
        ref = O.fir_fast(h, _window(x, lo, start + 4096))[start - lo:]
        assert nerr(_window(y, start, start + 4096), ref) <= TOL
    # impulse response: exact alignment at large indices.  y[n0 + k] = h[T-1-k] (fir/mod.rs:86, reversed taps)
    pos = [7, (1 << 29) + 1, (1 << 30) - 600]
    x.zero_()
    for p in pos:
        x[p] = 1.0
    y = FIRFilter(h, 1.0).execute_block(x)
    hr = h[::-1].astype(np.complex64)
    for p in pos:
        got = _window(y, p, p + 512)
        # tensor-core path: taps enter as TF32 hi + lo, exact to 2^-22; a one-sample shift is off by > 1e-3
        assert np.max(np.abs(got - hr)) <= 1e-6 * np.max(np.abs(hr))
        assert np.max(np.abs(got[1:] - hr[:-1])) > 1e-3 * np.max(np.abs(hr))
    assert int(torch.count_nonzero(y).item()) == 3 * int(np.count_nonzero(hr))
    # linearity on the full stream: F(a*u + b*v) = a*F(u) + b*F(v), checked on reductions of the whole output
    del y
    u = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(u).uniform_(-1, 1, generator=g)
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    a, b = 0.75, -1.5
    yu = FIRFilter(h, 1.0).execute_block(u)
    yv = FIRFilter(h, 1.0).execute_block(x)
    u.mul_(a).add_(x, alpha=b)
    ymix = FIRFilter(h, 1.0).execute_block(u)
    yu.mul_(a).add_(yv, alpha=b).sub_(ymix)
    assert float(yu.abs().max().item()) <= TOL * float(ymix.abs().max().item())


def test_config3_decimator_full_batch(torch_cuda):
    torch = torch_cuda
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter
    h = f32_taps(O.firdes_kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
    C, n, M = 4096, 1 << 20, 8
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    y = DecimatingFIRFilter(h, 1.0, M, n_channels=C).execute_block(x)
    assert y.shape == (C, n // M)                       # exact sample count (decim.rs:221-228)
    for c in (0, 2047, 4095):                           # element offsets beyond 2^32 in the last channel
        for start in (0, n - (1 << 14)):
            xs = _window(x[c], max(0, start - 256), start + (1 << 14))
            pre = start - max(0, start - 256)
            ref = O.fir_fast(h, xs, 1.0, M, count0=(start - pre) % M)[-(1 << 14) // M:]
            got = _window(y[c], start // M, start // M + (1 << 14) // M)
            assert nerr(got, ref) <= TOL
    # phase alignment: an impulse at input M*m0 + M-1 produces h[T-1] at output m0 (decim.rs:115-118)
    x.zero_()
    m0 = (n // M) - 40
    x[4095, M * m0 + M - 1] = 1.0
    y = DecimatingFIRFilter(h, 1.0, M, n_channels=C).execute_block(x)
    got = _window(y[4095], m0, m0 + 32)
    ref = h[::-1][::M][:32].astype(np.complex64)       # y[m0 + q] = g[q*M], g = reversed taps
    assert np.array_equal(got, ref)
    assert int(torch.count_nonzero(y).item()) == int(np.count_nonzero(ref))


def test_config4_interpolator_full_batch(torch_cuda):
    torch = torch_cuda
    from solid_dsp_b200.filter.fir import InterpolatingFIRFilter
    h = f32_taps(O.firdes_kaiser(128, 0.5 / 4 * 0.9, 80.0, 0.0))
    C, n, L = 1024, 1 << 20, 4
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    y = InterpolatingFIRFilter(h, L, n_channels=C).execute_block(x)
    assert y.shape == (C, n * L)                        # exact sample count (interp.rs:102-111)
    for c in (0, 511, 1023):
        for start in (0, n - (1 << 13)):
            lo = max(0, start - 32)
            ref = O.firinterp_fast(h, L, _window(x[c], lo, start + (1 << 13)))[(start - lo) * L:]
            assert nerr(_window(y[c], start * L, (start + (1 << 13)) * L), ref) <= TOL
    # an impulse at input n0 of the last channel: y[(n0 + j)*L + p] = hpad[p + (S-1-j)*L] -- every
    # sub-filter is applied reversed (pfb.rs:24-49,85-90)
    x.zero_()
    n0 = n - 50
    x[1023, n0] = 1.0
    y = InterpolatingFIRFilter(h, L, n_channels=C).execute_block(x)
    got = _window(y[1023], n0 * L, n0 * L + 128)
    assert np.array_equal(got, h.reshape(32, L)[::-1].ravel().astype(np.complex64))
    assert int(torch.count_nonzero(y).item()) == int(np.count_nonzero(h))


def test_config5_iir_full_batch_and_stream(torch_cuda):
    torch = torch_cuda
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    ff, fb = stable_lowpass_sections(8)
    C, n = 65536, 1 << 14
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    f = IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=C)
    y = f.execute_block(x)
    assert y.shape == (C, n)
    for c in (0, 32767, 65535):
        ref, st = O.sos_cascade_fast(ff, fb, x[c].cpu().numpy())
        assert nerr(y[c].cpu().numpy(), ref) <= TOL
    state, _ = f.get_state()
    assert nerr(state[65535], st.ravel()) <= 1e-4
    del x, y
    # one stream of 2^28: windows checked against the oracle warmed up over the preceding 4096 samples
    n = 1 << 28
    s = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(s).uniform_(-1, 1, generator=g)
    ys = IIRFilter(ff, fb, IIRFilterType.SecondOrder).execute_block(s)
    assert ys.shape == (n,)
    for start in (0, (1 << 27) + 777, n - 8192):
        lo = max(0, start - 4096)
        ref, _ = O.sos_cascade_fast(ff, fb, _window(s, lo, start + 8192))
        assert nerr(_window(ys, start, start + 8192), ref[start - lo:]) <= TOL


def test_ddc_full_batch(torch_cuda):
    """the NCO mix-down fused into config 3's decimator, at config 3's full size."""
    torch = torch_cuda
    from solid_dsp_b200.filter.ddc import DigitalDownConverter
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter
    h = f32_taps(O.firdes_kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
    C, n, M = 4096, 1 << 20, 8
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.empty((C, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    # zero frequency, zero phase: the phasor is exactly (1, 0), so the DDC IS the decimator -- bit for bit
    d0 = DigitalDownConverter(h, 1.0, M, frequency=0.0, n_channels=C)
    y0 = d0.execute_block(x)
    assert d0.last_fused and y0.shape == (C, n // M)
    yd = DecimatingFIRFilter(h, 1.0, M, n_channels=C).execute_block(x)
    assert torch.equal(y0, yd)
    del y0, yd
    # a real mix: oracle windows at the start and at the end of the first, a middle and the last channel (the 32-bit
    # phase accumulator has wrapped tens of thousands of times by then), channels with their own frequency and phase
    d = DigitalDownConverter(h, 1.0, M, frequency=0.6183, n_channels=C)
    d.nco.set_frequency(-2.9, channel=2047)
    d.nco.set_phase(1.0, channel=4095)
    raw = {c: d.nco.raw(c) for c in (0, 2047, 4095)}
    y = d.execute_block(x)
    assert d.last_fused and y.shape == (C, n // M)
    for c in (0, 2047, 4095):
        theta, delta = raw[c]
        for start in (0, n - (1 << 14)):
            pre = start - max(0, start - 256)
            lo = start - pre
            xs = _window(x[c], lo, start + (1 << 14))
            ref = O.ddc_fast(h, xs, 1.0, M, count0=lo % M, raw=((theta + lo * delta) & 0xFFFFFFFF, delta))[-(1 << 14) // M:]
            got = _window(y[c], start // M, start // M + (1 << 14) // M)
            assert nerr(got, ref) <= TOL
    # the accumulators have advanced by exactly n steps (nco/mod.rs:93-96 per sample)
    for c in (0, 2047, 4095):
        theta, delta = raw[c]
        assert d.nco.raw(c) == ((theta + n * delta) & 0xFFFFFFFF, delta)
