"""GPU parity for the tensor-core FIR (csrc/fir_tc.cu): long filters on tcgen05 as a banded-Toeplitz product, in
both operand formats (F16x2 block floating point with 3 products, the default; BF16x3 with 6 products).  Same bar as
the FFMA2 kernels: max normalised error <= 1e-5 against the f64 oracle on identical f32-rounded inputs, exact output
counts, exact alignment, streaming across calls, and the reference's handling of non-finite samples."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr

pytestmark = pytest.mark.gpu

N_MIN = 1 << 21  # comfortably above the dispatcher's threshold (2^15 samples: shorter calls go to the FFMA2 kernel)


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


@pytest.fixture(scope="module", params=["f16", "bf16"])
def FIR(request):
    """Both operand formats of the fused kernel: F16x2 block floating point (default) and BF16x3 (SGPU_FIR_TC_FMT=bf16)."""
    import os
    from solid_dsp_b200.filter.fir import FIRFilter
    os.environ["SGPU_FIR_TC_FMT"] = request.param
    yield FIRFilter
    del os.environ["SGPU_FIR_TC_FMT"]


def _rand(torch, n, seed, lo=-1.0, hi=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(lo, hi, generator=g)
    return x


def _check_windows(h, x, y, starts, width=2048, hist=None):
    T = len(h)
    worst = 0.0
    for start in starts:
        lo = max(0, start - (T - 1))
        seg = x[lo:start + width].cpu().numpy()
        if hist is not None and start - (T - 1) < 0:
            need = (T - 1) - start
            seg = np.concatenate([hist[len(hist) - need:], seg])
            ref = O.fir_fast(h, seg)[need + start - lo:]
        else:
            ref = O.fir_fast(h, seg)[start - lo:]
        got = y[start:start + width].cpu().numpy()
        worst = max(worst, nerr(got, ref[:len(got)]))
    return worst


@pytest.mark.parametrize("T", [112, 128, 200, 512, 777, 2048])
def test_parity_random(torch, FIR, T):
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    # ragged: last tile partial, not a multiple of 128; 310 tiles = up to 3 tiles per CTA (ring + barrier phases)
    n = (16384 * 310 if T in (200, 512) else N_MIN) + 12345
    x = _rand(torch, n, 100 + T)
    f = FIR(h, 0.5)
    y = f.execute_block(x)
    assert f.last_path == "tensor"
    assert y.shape == (n,)
    h_scaled = h  # scale applied by the oracle below
    starts = (0, T - 1, 16384 - 7, 16384 * 100 - 100, n // 2 + 3, n - 2048)
    worst = 0.0
    for s in starts:
        lo = max(0, s - (T - 1))
        ref = O.fir_fast(h_scaled, x[lo:s + 2048].cpu().numpy(), scale=0.5)[s - lo:]
        worst = max(worst, nerr(y[s:s + 2048].cpu().numpy(), ref))
    assert worst <= TOL, worst
    # the FFMA2 kernel on the same input: two independent implementations agree over the whole stream
    import os
    os.environ["SGPU_FIR_TC"] = "0"
    try:
        f2 = FIR(h, 0.5)
        y2 = f2.execute_block(x)
        assert f2.last_path == "ffma"
    finally:
        del os.environ["SGPU_FIR_TC"]
    assert (y - y2).abs().max().item() <= TOL * y2.abs().max().item()


def test_dc_and_positive_taps_bias(torch, FIR):
    """Worst case for the tensor core's truncating accumulator: all products of one sign (tools/tc_accum_probe.py)."""
    for T in (512, 4096):
        h = f32_taps(np.hanning(T + 2)[1:-1] / T)
        n = N_MIN
        x = torch.full((n,), 0.7 - 0.3j, dtype=torch.complex64, device="cuda")
        f = FIR(h, 1.0)
        y = f.execute_block(x)
        assert f.last_path == "tensor"
        assert _check_windows(h, x, y, (0, n // 2, n - 2048)) <= TOL / 4


def test_streaming_split_calls(torch, FIR):
    """concat(execute_block(a), execute_block(b)) == execute_block(a ++ b): the history tail feeds the split planes."""
    T = 512
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n1, n2, n3 = N_MIN + 5, 1000, N_MIN + 16384 * 3 + 1
    x = _rand(torch, n1 + n2 + n3, 7)
    f = FIR(h, 1.0)
    y1 = f.execute_block(x[:n1])
    assert f.last_path == "tensor"
    y2 = f.execute_block(x[n1:n1 + n2])  # short call: FFMA2 kernel, same history buffers
    assert f.last_path == "ffma"
    y3 = f.execute_block(x[n1 + n2:])  # odd sample offset: 8-byte aligned input, scalar split path
    assert f.last_path == "tensor"
    y = torch.cat([y1, y2, y3])
    starts = (0, n1 - 300, n1 + n2 - 100, n1 + n2 + 100, n1 + n2 + n3 - 2048)
    assert _check_windows(h, x, y, starts) <= TOL
    # state after the calls equals the last T-1 inputs
    hist, _ = f.get_state()
    assert np.array_equal(np.asarray(hist).reshape(-1), x[-(T - 1):].cpu().numpy())


def test_impulse_alignment_and_zeros(torch, FIR):
    T = 512
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n = 16384 * 160 + 999
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    pos = [0, 1151, 2176, 16383, 16384 * 149 + 5, n - 600]  # more than T apart; block and tile boundaries
    for p in pos:
        x[p] = 1.0 - 2.0j
    f = FIR(h, 1.0)
    y = f.execute_block(x)
    assert f.last_path == "tensor"
    hr = h[::-1]
    for p in pos:
        got = y[p:p + T].cpu().numpy()
        ref = hr * (1.0 - 2.0j)
        assert np.max(np.abs(got - ref)) <= 1e-6 * np.max(np.abs(ref))
        # a one-sample shift would be unmistakable
        assert np.max(np.abs(got[1:] - ref[:-1])) > 1e-3 * np.max(np.abs(ref))
    # exact zeros everywhere else: zero products sum to zero in the tensor core too
    mask = torch.ones(n, dtype=torch.bool, device="cuda")
    for p in pos:
        mask[p:p + T] = False
    assert int(torch.count_nonzero(y[mask]).item()) == 0


def test_two_channels_and_linearity(torch, FIR):
    T = 256
    h = f32_taps(O.firdes_kaiser(T, 0.2, 70.0, 0.0))
    n = N_MIN
    u, v = _rand(torch, n, 11), _rand(torch, n, 12)
    f = FIR(h, 1.0, n_channels=2)
    y = f.execute_block(torch.stack([u, v]))
    assert f.last_path == "tensor" and y.shape == (2, n)
    assert _check_windows(h, u, y[0], (0, n - 2048)) <= TOL
    assert _check_windows(h, v, y[1], (0, n - 2048)) <= TOL
    w = FIR(h, 1.0).execute_block(0.75 * u - 1.5 * v)
    lin = 0.75 * y[0] - 1.5 * y[1]
    assert (w - lin).abs().max().item() <= 4 * TOL * lin.abs().max().item()


def test_block_floating_dynamic_range(torch, FIR):
    """F16x2 scales every group of 64 samples by its own power of two: a stream whose level wanders over 2^+-40 from
    one stretch to the next, a burst after silence, tiny and huge taps, zeros -- each window is checked against the
    oracle relative to ITS OWN peak (a global-scale fp16 split would lose the quiet stretches entirely)."""
    T = 512
    n = 16384 * 150 + 77
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1, 1, generator=g)
    # level: piecewise constant over stretches of 50 000 samples, 2^k with k in [-40, 40]
    k = torch.randint(-40, 41, ((n + 49999) // 50000,), generator=g, device="cuda").repeat_interleave(50000)[:n]
    x = x * torch.pow(torch.tensor(2.0, device="cuda"), k.to(torch.float32))
    x[200000:260000] = 0  # exact silence
    x[300000] = 3e20      # one huge sample
    for hs in (1.0, 2.0 ** -20, 2.0 ** -60, 2.0 ** 50):
        h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0) * hs)
        f = FIR(h, 1.0)
        y = f.execute_block(x)
        assert f.last_path == "tensor"
        starts = [0, 49000, 100000 + 17, 199000, 230000, 259000, 299000, 1234567, n - 2048]
        for s0 in starts:
            lo = max(0, s0 - (T - 1))
            ref = O.fir_fast(h, x[lo:s0 + 2048].cpu().numpy())[s0 - lo:]
            got = y[s0:s0 + 2048].cpu().numpy()
            if np.max(np.abs(ref)) == 0:
                assert np.max(np.abs(got)) == 0
            else:
                assert nerr(got, ref) <= TOL, (hs, s0, nerr(got, ref))


@pytest.mark.parametrize("T", [512, 500, 130])
def test_non_finite_samples_match_the_reference(torch, FIR, T):
    """fir/mod.rs:209-212: a NaN / Inf sample reaches exactly the outputs whose T-sample window holds it, component by
    component (real taps scale re and im separately).  The banded GEMM would smear it over whole 128-output blocks, so
    the kernel flags such tiles and fir_tc_post_kernel recomputes them in the reference's order."""
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n = (1 << 21) + 12345
    x = _rand(torch, n, 4242 + T)
    rng = np.random.default_rng(T)
    pos = np.unique(np.concatenate([rng.integers(0, n, 40), [0, 1, 127, 128, 16383, 16384, 16385, n - 1, n - T, 777777]]))
    vals = [complex(np.nan, 0.5), complex(np.inf, -1.0), complex(0.25, -np.inf), complex(np.nan, np.nan), complex(-np.inf, np.inf)]
    xv = x.cpu().numpy()
    for i, p0 in enumerate(pos):
        xv[p0] = vals[i % len(vals)]
    x = torch.from_numpy(xv).cuda()
    f = FIR(h, 0.5)
    y = f.execute_block(x).cpu().numpy()
    assert f.last_path == "tensor"
    # expected masks straight from the definition: component c of output n is non-finite iff component c of some
    # x[n-T+1 .. n] is non-finite (finite taps; 0-valued taps still multiply: 0 * Inf = NaN)
    bad_re = np.convolve((~np.isfinite(xv.real)).astype(np.float64), np.ones(T))[:n] > 0
    bad_im = np.convolve((~np.isfinite(xv.imag)).astype(np.float64), np.ones(T))[:n] > 0
    assert np.array_equal(~np.isfinite(y.real), bad_re)
    assert np.array_equal(~np.isfinite(y.imag), bad_im)
    # ... and the oracle agrees with that definition on a window, and with the finite values everywhere else
    for s0 in (0, int(pos[5]) - 100 if pos[5] > 100 else 0, 16384 - 600, n - 3000):
        s0 = max(s0, 0)
        lo = max(0, s0 - (T - 1))
        ref = O.fir_fast(h, xv[lo:s0 + 3000], scale=0.5)[s0 - lo:]
        got = y[s0:s0 + 3000]
        assert np.array_equal(np.isfinite(ref.real), np.isfinite(got.real))
        assert np.array_equal(np.isfinite(ref.imag), np.isfinite(got.imag))
        ok = np.isfinite(ref.real) & np.isfinite(ref.imag)
        if ok.any():
            assert np.max(np.abs(got[ok] - ref[ok])) <= TOL * np.max(np.abs(ref[ok]))
    # the flags are cleared: the next call on clean data is clean
    x2 = _rand(torch, 1 << 20, 99)
    f2 = FIR(h, 0.5)
    y2 = f.execute_block(x2)  # same handle: its history still holds non-finite samples at most T-1 deep
    y3 = f2.execute_block(x2)
    assert bool(torch.isfinite(torch.view_as_real(y2[T:])).all())
    assert (y2[T:] - y3[T:]).abs().max().item() <= TOL * y3.abs().max().item()


# ------------------------------------------------------------------ polyphase interpolator on the same kernel
@pytest.mark.parametrize("L,T", [(4, 256), (2, 128), (4, 200), (4, 1024), (2, 777), (4, 128), (2, 64), (4, 90), (2, 30)])
def test_interpolator_parity(torch, L, T, monkeypatch):
    """InterpolatingFIRFilter (interp.rs:102-111, pfb.rs:85-90): block rows of 128 / L inputs, exact output count,
    padded sub-filters (T not a multiple of L), several channels in one launch, ragged tails."""
    from solid_dsp_b200.filter.fir import InterpolatingFIRFilter
    # sub-filters of <= 32 taps (one accumulation chain per tile, the alternating-group epilogue) are dispatched to the
    # FP32 walking kernel by default: force them through the tensor kernel here
    monkeypatch.setenv("SGPU_INTERP_TC_MIN_SUB", "1")
    h = f32_taps(O.firdes_kaiser(T, 0.5 / L * 0.9, 80.0, 0.0))
    C = 3
    n = (1 << 22) // L + 4321  # n * L * C >= 2^23 outputs: the dispatcher's threshold for the tensor path
    x = torch.stack([_rand(torch, n, 40 + c) for c in range(C)])
    f = InterpolatingFIRFilter(h, L, n_channels=C)
    y = f.execute_block(x)
    assert f.last_path == "tensor"
    assert y.shape == (C, n * L)
    S = f.sub_len()
    for c in range(C):
        for start in (0, 5000, n // 2 + 11, n - 3000):
            lo = max(0, start - S)
            ref = O.firinterp_fast(h, L, x[c, lo:start + 3000].cpu().numpy())[(start - lo) * L:]
            got = y[c, start * L:(start + 3000) * L].cpu().numpy()
            assert nerr(got, ref) <= TOL
    # same input through the FP32 walking / tile kernels
    import os
    os.environ["SGPU_FIR_TC"] = "0"
    try:
        f2 = InterpolatingFIRFilter(h, L, n_channels=C)
        y2 = f2.execute_block(x)
        assert f2.last_path == "ffma"
    finally:
        del os.environ["SGPU_FIR_TC"]
    assert (y - y2).abs().max().item() <= TOL * y2.abs().max().item()


def test_interpolator_streaming_and_impulse(torch):
    from solid_dsp_b200.filter.fir import InterpolatingFIRFilter
    L, T = 4, 384
    h = f32_taps(O.firdes_kaiser(T, 0.5 / L * 0.9, 80.0, 0.0))
    n1, n2 = (1 << 21) + 3, (1 << 21) + 4096 * 5 + 1
    x = _rand(torch, n1 + n2, 77)
    f = InterpolatingFIRFilter(h, L)
    y = torch.cat([f.execute_block(x[:n1]), f.execute_block(x[n1:])])
    assert f.last_path == "tensor"
    S = f.sub_len()
    for start in (n1 - 100, n1, n1 + 7, n1 + n2 - 2000):
        lo = start - S
        ref = O.firinterp_fast(h, L, x[lo:start + 2000].cpu().numpy())[(start - lo) * L:]
        assert nerr(y[start * L:(start + 2000) * L].cpu().numpy(), ref) <= TOL
    # impulse at input n0: exact phase alignment against the oracle (a one-sample shift would be unmistakable)
    f = InterpolatingFIRFilter(h, L)
    z = torch.zeros(n1, dtype=torch.complex64, device="cuda")
    n0 = 4096 * 37 + 5
    z[n0] = 2.0
    yz = f.execute_block(z)
    ref = O.firinterp_fast(h, L, z[n0 - S:n0 + S + 8].cpu().numpy())[S * L:]
    got = yz[n0 * L:(n0 + S + 8) * L].cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-6 * np.max(np.abs(ref))
    assert np.max(np.abs(got[1:] - ref[:-1])) > 1e-3 * np.max(np.abs(ref))
    assert int(torch.count_nonzero(yz).item()) == int(np.count_nonzero(ref))


# ------------------------------------------------------------------ complex taps (Coef = Complex<f64>) on the tensor kernel
@pytest.mark.parametrize("T", [56, 64, 128, 512, 777])
def test_complex_taps(torch, T):
    """FIRFilter<Complex<f64>, Complex<f64>> (fir/mod.rs:181-186): complex taps and complex scale; the cross terms
    run as N = 128 MMAs on the re / im halves of the split planes (D_re -= Gi Xim, D_im += Gi Xre)."""
    from solid_dsp_b200.filter.fir import FIRFilter
    k = np.arange(T)
    hr = O.firdes_kaiser(T, 0.1, 80.0, 0.0) * np.exp(2j * np.pi * 0.05 * k)  # config 2's complex-tap variant
    h = f32_taps(hr.real) + 1j * f32_taps(hr.imag)
    scale = 0.5 - 0.25j
    n = 16384 * 300 + 4321
    x = _rand(torch, n, 900 + T)
    f = FIRFilter(h, scale)
    y1 = f.execute_block(x[:n // 2])
    assert f.last_path == "tensor"
    y = torch.cat([y1, f.execute_block(x[n // 2:])])
    for s in (0, T - 1, n // 2 - 100, n // 2 + 5, n - 2048):
        lo = max(0, s - (T - 1))
        ref = O.fir_fast(h, x[lo:s + 2048].cpu().numpy(), scale)[s - lo:]
        assert nerr(y[s:s + 2048].cpu().numpy(), ref) <= TOL
    import os
    os.environ["SGPU_FIR_TC"] = "0"
    try:
        f2 = FIRFilter(h, scale)
        y2 = f2.execute_block(x)
        assert f2.last_path == "ffma"
    finally:
        del os.environ["SGPU_FIR_TC"]
    assert (y - y2).abs().max().item() <= TOL * y2.abs().max().item()


def test_default_dispatch_rule(torch, FIR, monkeypatch):
    """Without SGPU_FIR_TC_MIN_SAMPLES the library picks the kernel by the size of the call (csrc/fir.cu:
    tc_call_is_long_enough, measured with tools/call_size_crossover.py): short calls of a long filter stay on the FFMA2
    kernel, long calls go to the tensor cores, and a stream may cross the rule in either direction without a seam."""
    monkeypatch.delenv("SGPU_FIR_TC_MIN_SAMPLES", raising=False)
    T = 512
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n_short, n_long = 1 << 17, (1 << 20) + 3
    x = _rand(torch, n_short + n_long + n_short, 11)
    f = FIR(h, 1.0)
    y1 = f.execute_block(x[:n_short])
    assert f.last_path == "ffma"      # 2^17 samples x 512 taps: 18.5 us on FFMA2, 28.7 us on the tensor kernel
    y2 = f.execute_block(x[n_short:n_short + n_long])
    assert f.last_path == "tensor"    # 2^20 x 512 = 2^29 sample-taps
    y3 = f.execute_block(x[n_short + n_long:])
    assert f.last_path == "ffma"
    y = torch.cat([y1, y2, y3]).cpu().numpy()
    xs = x.cpu().numpy()
    for s0 in (0, n_short - 200, n_short + n_long - 300, y.size - 4096):
        lo = max(0, s0 - (T - 1))
        ref = O.fir_fast(h, xs[lo:s0 + 4096])[s0 - lo:]
        assert nerr(y[s0:s0 + 4096], ref) <= TOL
    # a short filter needs a longer call: 128 taps x 2^20 samples is still FFMA2, x 2^22 is tensor
    h128 = f32_taps(O.firdes_kaiser(128, 0.1, 80.0, 0.0))
    f = FIR(h128, 1.0)
    f.execute_block(x[:1 << 20])
    assert f.last_path == "ffma"
    xx = _rand(torch, 1 << 22, 12)
    f.execute_block(xx)
    assert f.last_path == "tensor"
