"""GPU parity: FIRFilter / DecimatingFIRFilter / InterpolatingFIRFilter / PolyPhaseFilterBank
through the C ABI versus the CPU oracle, on identical f32-rounded inputs."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fir():
    from solid_dsp_b200.filter import fir
    return fir


def _cx(pairs):
    return np.array([complex(a, b) for a, b in pairs])


# ------------------------------------------------------------------ reference doc-test goldens
def test_reference_goldens(fir, golden):
    ref = golden["reference_doctests"]
    g = ref["fir_execute"]
    out = fir.FIRFilter(g["coefs"], g["scale"]).execute(g["input"][0])
    assert abs(out[0] - g["expect_first"]) <= TOL * abs(g["expect_first"])
    g = ref["fir_execute_block"]
    out = fir.FIRFilter(g["coefs"], g["scale"]).execute_block(g["input"])
    assert len(out) == 5 and abs(out[4] - g["expect_value"]) <= TOL * g["expect_value"]
    g = ref["fir_decim_execute"]
    f = fir.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"])
    a = f.execute(g["input"][0])
    b = f.execute(g["input"][1])
    assert len(a) == 0 and len(b) == 1 and abs(b[0] - 28.28) <= TOL * 28.28
    g = ref["fir_decim_execute_block"]
    out = fir.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"]).execute_block(g["input"])
    assert nerr(out, g["expect"]) <= TOL and len(out) == 2


def test_derived_vectors(fir, golden):
    der = golden["derived_vectors"]
    x = _cx(der["x"])
    for key in ("fir_123", "fir_123_scale_half"):
        g = der[key]
        assert nerr(fir.FIRFilter(g["coefs"], g["scale"]).execute_block(x), _cx(g["expect"])) <= TOL
    g = der["decim_123_m2"]
    assert nerr(fir.DecimatingFIRFilter(g["coefs"], g["scale"], g["decimation"]).execute_block(x),
                _cx(g["expect"])) <= TOL
    for key in ("interp_6taps_l2", "interp_5taps_l2_padded"):
        g = der[key]
        out = fir.InterpolatingFIRFilter(g["coefs"], g["interpolation"]).execute_block(x[: g["n_in"]])
        assert nerr(out, _cx(g["expect"])) <= TOL
    g = der["interp_5taps_l4_impulse"]
    f = fir.InterpolatingFIRFilter(g["coefs"], g["interpolation"])
    f.set_scale(7.0)  # stored, never applied
    assert f.get_scale() == 7.0 and f.sub_len() == 2
    out = f.execute_block(_cx(g["input"]))
    assert np.array_equal(out, _cx(g["expect"]).astype(np.complex64))  # exact: integers


# ------------------------------------------------------------------ FIR
@pytest.mark.parametrize("T", [1, 2, 5, 31, 32, 33, 64, 257, 512, 1000])
@pytest.mark.parametrize("n", [1, 7, 2047, 2048, 2049, 6000])
def test_fir_random(fir, T, n):
    rng = np.random.default_rng(1000 * T + n)
    h = f32_taps(rng.uniform(-1, 1, T))
    x = rand_cf32(rng, n)
    got = fir.FIRFilter(h, 0.75).execute_block(x)
    ref = O.fir_fast(h, x, 0.75)
    assert got.shape == (n,)
    assert nerr(got, ref) <= TOL


def test_fir_kaiser_config1_shape(fir):
    """BASELINE config 1 at reduced length: 64-tap Kaiser low-pass, single channel."""
    rng = np.random.default_rng(1)
    h = f32_taps(O.firdes_kaiser(64, 0.25, 60.0, 0.0))
    x = rand_cf32(rng, 1 << 16)
    got = fir.FIRFilter(h, 1.0).execute_block(x)
    assert nerr(got, O.fir_fast(h, x)) <= TOL


def test_fir_multichannel_and_stride(fir):
    rng = np.random.default_rng(3)
    h = f32_taps(rng.uniform(-1, 1, 48))
    x = rand_cf32(rng, (5, 3001))
    got = fir.FIRFilter(h, 1.0, n_channels=5).execute_block(x)
    assert got.shape == (5, 3001)
    for c in range(5):
        assert nerr(got[c], O.fir_fast(h, x[c])) <= TOL


def test_fir_streaming_split_calls(fir):
    """concat(execute_block(a), execute_block(b)) == execute_block(a ++ b)."""
    rng = np.random.default_rng(5)
    h = f32_taps(rng.uniform(-1, 1, 100))
    x = rand_cf32(rng, (2, 5000))
    whole = fir.FIRFilter(h, 1.0, n_channels=2).execute_block(x)
    f = fir.FIRFilter(h, 1.0, n_channels=2)
    cuts = [0, 1, 2, 50, 99, 100, 2500, 2501, 5000]
    parts = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(whole, parts)  # same kernel arithmetic -> bit-identical
    ref = np.stack([O.fir_fast(h, x[c]) for c in range(2)])
    assert nerr(parts, ref) <= TOL


def test_fir_state_roundtrip_and_clone(fir):
    rng = np.random.default_rng(6)
    h = f32_taps(rng.uniform(-1, 1, 33))
    x = rand_cf32(rng, 400)
    f = fir.FIRFilter(h, 1.0)
    f.execute_block(x[:150])
    hist, cur = f.get_state()
    assert cur == 0 and np.array_equal(hist[0], x[150 - 32:150])
    g = f.clone()
    a = f.execute_block(x[150:])
    b = g.execute_block(x[150:])
    assert np.array_equal(a, b)
    k = fir.FIRFilter(h, 1.0)
    k.set_state(hist, 0)
    assert np.array_equal(k.execute_block(x[150:]), a)
    k.reset()
    assert nerr(k.execute_block(x), O.fir_fast(h, x)) <= TOL
    assert np.array_equal(f.coefficients(), h[::-1])  # stored (reversed) order
    assert f.len() == 33 and not f.is_empty()
    f.set_scale(2.5)
    assert f.get_scale() == 2.5


def test_fir_impulse_alignment(fir):
    """Asymmetric taps + impulse: a one-sample shift or a tap reversal is unmistakable."""
    h = np.arange(1, 41, dtype=np.float64)
    x = np.zeros(100, dtype=np.complex64)
    x[3] = 1.0
    got = fir.FIRFilter(h, 1.0).execute_block(x)
    ref = O.fir_fast(h, x)
    assert np.array_equal(got, ref.astype(np.complex64))
    assert got[3] == 40.0 and got[42] == 1.0 and got[43] == 0.0


def test_fir_device_pointers(fir):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    h = f32_taps(rng.uniform(-1, 1, 512))
    x = rand_cf32(rng, (3, 20000))
    xt = torch.from_numpy(x).cuda()
    got = fir.FIRFilter(h, 1.0, n_channels=3).execute_block(xt)
    assert got.is_cuda and got.shape == (3, 20000)
    host = fir.FIRFilter(h, 1.0, n_channels=3).execute_block(x)
    assert np.array_equal(got.cpu().numpy(), host)
    for c in range(3):
        assert nerr(host[c], O.fir_fast(h, x[c])) <= TOL


# ------------------------------------------------------------------ decimator
@pytest.mark.parametrize("M", [1, 2, 3, 8, 13])
@pytest.mark.parametrize("T", [1, 5, 64, 256, 300])
def test_decim_random(fir, M, T):
    rng = np.random.default_rng(100 * M + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    for n in (1, M - 1, M, 5 * M + 3, 9001):
        if n == 0:
            continue
        x = rand_cf32(rng, n)
        got = fir.DecimatingFIRFilter(h, 1.5, M).execute_block(x)
        ref = O.fir_fast(h, x, 1.5, M)
        assert len(got) == n // M == len(ref)
        assert nerr(got, ref) <= TOL


def test_decim_streaming_counter_and_write(fir):
    rng = np.random.default_rng(17)
    h = f32_taps(rng.uniform(-1, 1, 40))
    x = rand_cf32(rng, (2, 3000))
    M = 8
    whole = fir.DecimatingFIRFilter(h, 1.0, M, n_channels=2).execute_block(x)
    f = fir.DecimatingFIRFilter(h, 1.0, M, n_channels=2)
    cuts = [0, 3, 10, 11, 700, 701, 3000]
    parts = [f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert [p.shape[1] for p in parts] == [0, 1, 0, 86, 0, 288]
    assert np.array_equal(whole, np.concatenate(parts, axis=1))
    # write() advances the counter without output -- fir/decim.rs:136-139
    g = fir.DecimatingFIRFilter(h, 1.0, M)
    g.write(x[0, :6])
    assert g.get_state()[1] == 6
    a = g.execute_block(x[0, 6:])
    o = O.DecimatingFIRFilter(h, 1.0, M)
    o.write(x[0, :6])
    b = o.execute_block(x[0, 6:])
    assert len(a) == len(b) and nerr(a, b) <= TOL
    assert g.get_decimation() == M


def test_decim_impulse_phase(fir):
    """Output m is produced by input m*M + M-1: an impulse at index M-1 hits tap h[T-1] at m=0."""
    h = np.arange(1, 25, dtype=np.float64)
    M = 4
    x = np.zeros(64, dtype=np.complex64)
    x[M - 1] = 1.0
    got = fir.DecimatingFIRFilter(h, 1.0, M).execute_block(x)
    ref = O.fir_fast(h, x, 1.0, M)
    assert np.array_equal(got, ref.astype(np.complex64)) and got[0] == 24.0 and got[1] == 20.0


def test_decim_config3_shape_small(fir):
    """BASELINE config 3 reduced: M=8, 256-tap Kaiser, 16 channels x 2^14."""
    rng = np.random.default_rng(3)
    h = f32_taps(O.firdes_kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
    x = rand_cf32(rng, (16, 1 << 14))
    got = fir.DecimatingFIRFilter(h, 1.0, 8, n_channels=16).execute_block(x)
    assert got.shape == (16, 2048)
    for c in (0, 7, 15):
        assert nerr(got[c], O.fir_fast(h, x[c], 1.0, 8)) <= TOL


# ------------------------------------------------------------------ interpolator / PFB
@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("T", [1, 3, 7, 128, 131])
def test_interp_random(fir, L, T):
    rng = np.random.default_rng(100 * L + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    for n in (1, 2, 1023, 1024, 1025, 3000):
        x = rand_cf32(rng, n)
        f = fir.InterpolatingFIRFilter(h, L)
        got = f.execute_block(x)
        ref = O.firinterp_fast(h, L, x)
        assert len(got) == n * L
        assert nerr(got, ref) <= TOL
        assert f.sub_len() == O.interp_sub_len(T, L) and f.interpolation() == L


def test_interp_streaming_and_channels(fir):
    rng = np.random.default_rng(23)
    h = f32_taps(O.firdes_kaiser(128, 0.5 / 4 * 0.9, 80.0, 0.0))
    x = rand_cf32(rng, (3, 2500))
    whole = fir.InterpolatingFIRFilter(h, 4, n_channels=3).execute_block(x)
    f = fir.InterpolatingFIRFilter(h, 4, n_channels=3)
    cuts = [0, 1, 31, 32, 33, 1500, 2500]
    parts = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(whole, parts)
    for c in range(3):
        assert nerr(whole[c], O.firinterp_fast(h, 4, x[c])) <= TOL
    # coefficents(): per-phase stored order, flattened (interp.rs:77-79)
    o = O.InterpolatingFIRFilter(h, 4)
    import ctypes
    ref = np.zeros(128)
    O.lib().so_firinterp_coefficients(o._h, ref.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert np.array_equal(f.coefficents(), ref)


def test_pfb_push_execute(fir):
    rng = np.random.default_rng(29)
    h = f32_taps(rng.uniform(-1, 1, 23))  # 23 / 4 -> sub_len 5, 3 trailing taps dropped (pfb.rs:32)
    bank = fir.PolyPhaseFilterBank(h, 4, 2.0)
    ref = O.PolyPhaseFilterBank(h, 4, 2.0)
    assert bank.sub_len() == ref.sub_len() == 5 and bank.len() == 4
    assert np.array_equal(bank.coefficents(), ref.coefficents())
    x = rand_cf32(rng, 9)
    for s in x:
        bank.push(s)
        ref.push(s)
        for p in range(4):
            assert abs(bank.execute(p) - ref.execute(p)) <= TOL * 5


# ------------------------------------------------------------------ complex taps (Coef = Complex<f64>)
def _ctaps(rng, T):
    return f32_taps(rng.uniform(-1, 1, T)) + 1j * f32_taps(rng.uniform(-1, 1, T))


@pytest.mark.parametrize("T", [1, 3, 16, 17, 64, 200])
def test_fir_complex_taps(fir, T):
    """FIRFilter<Complex<f64>, Complex<f64>> is a legal instantiation (fir/mod.rs:181-186):
    complex taps, complex scale (Out * Coef, fir/mod.rs:211)."""
    rng = np.random.default_rng(T)
    h = _ctaps(rng, T)
    x = rand_cf32(rng, (2, 3001))
    scale = 0.5 - 0.25j
    f = fir.FIRFilter(h, scale, n_channels=2)
    got = np.concatenate([f.execute_block(x[:, :1000]), f.execute_block(x[:, 1000:])], axis=1)
    for c in range(2):
        assert nerr(got[c], O.fir_fast(h, x[c], scale)) <= TOL
    assert np.array_equal(f.coefficients(), h[::-1]) and f.get_scale() == scale
    g = f.clone()
    y = rand_cf32(rng, (2, 100))
    assert np.array_equal(f.execute_block(y), g.execute_block(y))


@pytest.mark.parametrize("M", [2, 4, 8, 5])
def test_decim_complex_taps(fir, M):
    rng = np.random.default_rng(M)
    h = _ctaps(rng, 96)
    x = rand_cf32(rng, 5003)
    got = fir.DecimatingFIRFilter(h, 1.0 + 0.5j, M).execute_block(x)
    ref = O.fir_fast(h, x, 1.0 + 0.5j, M)
    assert len(got) == len(ref) and nerr(got, ref) <= TOL


@pytest.mark.parametrize("L", [1, 2, 4, 3])
def test_interp_complex_taps(fir, L):
    rng = np.random.default_rng(L)
    h = _ctaps(rng, 50)
    x = rand_cf32(rng, (2, 1500))
    f = fir.InterpolatingFIRFilter(h, L, n_channels=2)
    got = np.concatenate([f.execute_block(x[:, :7]), f.execute_block(x[:, 7:])], axis=1)
    for c in range(2):
        assert nerr(got[c], O.firinterp_fast(h, L, x[c])) <= TOL
    bank = fir.PolyPhaseFilterBank(h, 4, 1.0)
    ref = O.PolyPhaseFilterBank(h, 4, 1.0)
    assert np.array_equal(bank.coefficents(), ref.coefficents())
    for s_ in x[0, :6]:
        bank.push(s_)
        ref.push(s_)
        for p_ in range(4):
            assert abs(bank.execute(p_) - ref.execute(p_)) <= TOL * 20


# ------------------------------------------------------------------ warp-private tile kernels (fir_walk.cuh)
@pytest.mark.parametrize("T", [33, 64, 512, 3000, 9000])
def test_fir_many_warp_tiles(fir, T):
    """Streams that span several warps, blocks and TPW groups of fir_warp_kernel, ragged end, split
    calls whose boundaries fall inside tiles; two channels with a row stride."""
    rng = np.random.default_rng(T)
    h = f32_taps(rng.uniform(-1, 1, T))
    n = 4 * 4096 * 4 + 777
    x = rand_cf32(rng, (2, n))
    f = fir.FIRFilter(h, 0.75, n_channels=2)
    cuts = [0, 5000, 5001, 40000, n]
    got = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    for c in range(2):
        assert nerr(got[c], O.fir_fast(h, x[c], 0.75)) <= TOL


@pytest.mark.parametrize("M,T", [(2, 40), (4, 128), (8, 256), (8, 700), (16, 512), (16, 100), (32, 1024), (32, 4000)])
def test_decim_many_warp_tiles(fir, M, T):
    rng = np.random.default_rng(M * 1000 + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    n = 150001
    x = rand_cf32(rng, (3, n))
    f = fir.DecimatingFIRFilter(h, 1.25, M, n_channels=3)
    cuts = [0, 3, 50003, 50004, n]
    got = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert got.shape == (3, n // M)
    for c in range(3):
        assert nerr(got[c], O.fir_fast(h, x[c], 1.25, M)) <= TOL


@pytest.mark.parametrize("L,T", [(2, 64), (4, 128), (8, 256), (4, 100), (8, 17), (16, 512), (16, 40), (32, 1000)])
def test_interp_many_warp_tiles(fir, L, T):
    """Sub-filters of <= 32 taps take fir_interp_walk_kernel: several warps / blocks / tiles per warp,
    ragged end inside a run, split calls."""
    rng = np.random.default_rng(L * 1000 + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    n = 60001
    x = rand_cf32(rng, (2, n))
    f = fir.InterpolatingFIRFilter(h, L, n_channels=2)
    cuts = [0, 7, 20000, 20013, n]
    got = np.concatenate([f.execute_block(x[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert got.shape == (2, n * L)
    for c in range(2):
        assert nerr(got[c], O.firinterp_fast(h, L, x[c])) <= TOL
    # phase alignment: an impulse at input n0 must put hpad[p + j*L] at output (n0 + j)*L + p
    imp = np.zeros(2000, dtype=np.complex64)
    imp[777] = 1.0
    y = fir.InterpolatingFIRFilter(h, L).execute_block(imp)
    ref = O.firinterp_fast(h, L, imp)
    assert np.array_equal(np.nonzero(y)[0], np.nonzero(ref)[0]) and nerr(y, ref) <= 1e-7


def test_more_than_65535_channels(fir):
    """Channels ride in grid.y (<= 65535): larger handles are launched in channel blocks.  The first,
    the 65535th / 65536th and the last channel against the oracle, FIR + decimator + interpolator."""
    rng = np.random.default_rng(65536)
    C, n = 70000, 300
    x = rand_cf32(rng, (C, n))
    h = f32_taps(rng.uniform(-1, 1, 40))
    picks = (0, 65534, 65535, 65536, C - 1)
    f = fir.FIRFilter(h, 1.0, n_channels=C)
    y = np.concatenate([f.execute_block(x[:, :100]), f.execute_block(x[:, 100:])], axis=1)
    d = fir.DecimatingFIRFilter(h, 1.0, 4, n_channels=C).execute_block(x)
    i = fir.InterpolatingFIRFilter(h, 2, n_channels=C).execute_block(x)
    for c in picks:
        assert nerr(y[c], O.fir_fast(h, x[c])) <= TOL
        assert nerr(d[c], O.fir_fast(h, x[c], 1.0, 4)) <= TOL
        assert nerr(i[c], O.firinterp_fast(h, 2, x[c])) <= TOL
