"""GPU parity for the boundary features added in round 2: per-channel taps, any decimation / interpolation factor,
one launch per execute_block (history written by the main kernel), rejected in-place calls, and one stream split
over several handles (the multi-GPU seam) on both arithmetic paths."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fir():
    from solid_dsp_b200.filter import fir
    return fir


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


# ------------------------------------------------------------------ per-channel taps (one reference object per channel)
@pytest.mark.parametrize("T", [7, 64, 200])
def test_fir_per_channel_taps(fir, T):
    rng = np.random.default_rng(T)
    Cn, n = 5, 20000
    h = f32_taps(rng.uniform(-1, 1, (Cn, T)))
    x = rand_cf32(rng, (Cn, n))
    f = fir.FIRFilter(h, 0.75)
    assert f.n_channels == Cn
    y = np.concatenate([f.execute_block(x[:, :7777]), f.execute_block(x[:, 7777:])], axis=1)
    for c in range(Cn):
        assert nerr(y[c], O.fir_fast(h[c], x[c], 0.75)) <= TOL
        assert np.array_equal(f.coefficients(c), h[c][::-1])
    g = f.clone()
    x2 = rand_cf32(rng, (Cn, 3000))
    assert np.array_equal(f.execute_block(x2), g.execute_block(x2))


def test_fir_per_channel_complex_taps_and_long_calls(fir, torch):
    """Long per-channel filters stay on the FP32 kernels (the tensor band matrix is shared by all channels)."""
    rng = np.random.default_rng(3)
    Cn, T, n = 3, 160, 1 << 17
    h = f32_taps(rng.uniform(-1, 1, (Cn, T))) + 1j * f32_taps(rng.uniform(-1, 1, (Cn, T)))
    x = rand_cf32(rng, (Cn, n))
    f = fir.FIRFilter(h, 1.0 - 0.5j)
    y = f.execute_block(torch.from_numpy(x).cuda()).cpu().numpy()
    assert f.last_path == "ffma"
    for c in range(Cn):
        assert nerr(y[c, :5000], O.fir_fast(h[c], x[c, :5000], 1.0 - 0.5j)) <= TOL
        assert nerr(y[c, -3000:], O.fir_fast(h[c], x[c, -3000 - T:], 1.0 - 0.5j)[T:]) <= TOL


@pytest.mark.parametrize("M,T", [(8, 256), (4, 37), (3, 50)])
def test_decimator_per_channel_taps(fir, M, T):
    rng = np.random.default_rng(10 * M + T)
    Cn, n = 4, 30011
    h = f32_taps(rng.uniform(-1, 1, (Cn, T)))
    x = rand_cf32(rng, (Cn, n))
    f = fir.DecimatingFIRFilter(h, 1.0, M)
    y = np.concatenate([f.execute_block(x[:, :10001]), f.execute_block(x[:, 10001:])], axis=1)
    assert y.shape == (Cn, n // M)
    for c in range(Cn):
        assert nerr(y[c], O.fir_fast(h[c], x[c], 1.0, M)) <= TOL


@pytest.mark.parametrize("L,T", [(4, 128), (2, 24), (5, 33)])
def test_interpolator_per_channel_taps(fir, L, T):
    rng = np.random.default_rng(100 * L + T)
    Cn, n = 3, 9000
    h = f32_taps(rng.uniform(-1, 1, (Cn, T)))
    x = rand_cf32(rng, (Cn, n))
    f = fir.InterpolatingFIRFilter(h, L)
    y = np.concatenate([f.execute_block(x[:, :4001]), f.execute_block(x[:, 4001:])], axis=1)
    assert y.shape == (Cn, n * L)
    for c in range(Cn):
        assert nerr(y[c], O.firinterp_fast(h[c], L, x[c])) <= TOL
    g = f.clone()
    x2 = rand_cf32(rng, (Cn, 500))
    assert np.array_equal(f.execute_block(x2), g.execute_block(x2))


# ------------------------------------------------------------------ any factor the reference takes
@pytest.mark.parametrize("M,T,cx", [(64, 256, False), (100, 256, False), (1000, 3000, False), (50, 300, True), (4096, 5000, False)])
def test_decimator_large_factors(fir, M, T, cx):
    """fir/decim.rs:27-42 takes any decimation: shapes whose M phase planes do not fit one SM's shared memory run on the
    direct kernel -- counts, phase and values as the oracle's, across split calls."""
    rng = np.random.default_rng(M + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    if cx:
        h = h + 1j * f32_taps(rng.uniform(-1, 1, T))
    n = 20 * M + 777
    x = rand_cf32(rng, (2, n))
    f = fir.DecimatingFIRFilter(h, 0.5, M, n_channels=2)
    cut = 3 * M + 5
    y = np.concatenate([f.execute_block(x[:, :cut]), f.execute_block(x[:, cut:])], axis=1)
    assert y.shape == (2, n // M)
    for c in range(2):
        assert nerr(y[c], O.fir_fast(h, x[c], 0.5, M)) <= TOL
    hist, cur = f.get_state()
    assert cur == n % M
    assert np.array_equal(hist[0][-min(T - 1, n):], x[0][-min(T - 1, n):])


@pytest.mark.parametrize("L,T,cx", [(100, 400, False), (28, 300, True), (64, 5000, False), (7, 100, False), (1000, 1000, False)])
def test_interpolator_large_factors(fir, L, T, cx):
    rng = np.random.default_rng(L + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    if cx:
        h = h + 1j * f32_taps(rng.uniform(-1, 1, T))
    n = 700
    x = rand_cf32(rng, (2, n))
    f = fir.InterpolatingFIRFilter(h, L, n_channels=2)
    y = np.concatenate([f.execute_block(x[:, :301]), f.execute_block(x[:, 301:])], axis=1)
    assert y.shape == (2, n * L)
    for c in range(2):
        assert nerr(y[c], O.firinterp_fast(h, L, x[c])) <= TOL


# ------------------------------------------------------------------ in-place calls are refused
def test_overlapping_in_out_is_rejected(fir, torch):
    from solid_dsp_b200 import _ffi
    h = f32_taps(O.firdes_kaiser(64, 0.25, 60.0, 0.0))
    f = fir.FIRFilter(h, 1.0)
    buf = torch.zeros(1 << 16, dtype=torch.complex64, device="cuda")
    got = _ffi.c_size()
    s = torch.cuda.current_stream().cuda_stream
    for off in (0, 100, (1 << 15) - 1):
        st = _ffi.lib.sgpu_fir_execute_block(f._h, buf.data_ptr(), 1 << 15, 1 << 15, buf.data_ptr() + 8 * off, 1 << 15,
                                             C.byref(got), _ffi.DEVICE, s)
        assert st == _ffi.ERR_INVALID_ARGUMENT
        assert b"overlap" in _ffi.lib.sgpu_last_error()
    st = _ffi.lib.sgpu_fir_execute_block(f._h, buf.data_ptr(), 1 << 15, 1 << 15, buf.data_ptr() + 8 * (1 << 15), 1 << 15,
                                         C.byref(got), _ffi.DEVICE, s)
    assert st == _ffi.OK
    fi = fir.InterpolatingFIRFilter(h, 2)
    st = _ffi.lib.sgpu_interp_execute_block(fi._h, buf.data_ptr(), 1000, 1000, buf.data_ptr() + 8 * 500, 2000,
                                            C.byref(got), _ffi.DEVICE, s)
    assert st == _ffi.ERR_INVALID_ARGUMENT


# ------------------------------------------------------------------ one launch per execute_block
def test_one_launch_per_execute_block(fir, torch):
    """the history update is fused into the kernel that computes the outputs (window/mod.rs:63-71)."""
    from solid_dsp_b200 import launch_count
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rand_cf32(rng, (4, 1 << 14))).cuda()
    h64 = f32_taps(O.firdes_kaiser(64, 0.25, 60.0, 0.0))
    h256 = f32_taps(O.firdes_kaiser(256, 0.05, 80.0, 0.0))
    cases = [fir.FIRFilter(h64, 1.0, n_channels=4), fir.DecimatingFIRFilter(h256, 1.0, 8, n_channels=4),
             fir.DecimatingFIRFilter(h256, 1.0, 5, n_channels=4), fir.InterpolatingFIRFilter(h64, 4, n_channels=4),
             fir.InterpolatingFIRFilter(h256, 3, n_channels=4), fir.DecimatingFIRFilter(h256, 1.0, 100, n_channels=4)]
    for f in cases:
        f.execute_block(x)
        l0 = launch_count()
        f.execute_block(x)
        assert launch_count() - l0 == 1, type(f).__name__
    # and the state it leaves is the reference's: the last T-1 inputs
    f = cases[0]
    hist, _ = f.get_state()
    assert np.array_equal(hist, x[:, -63:].cpu().numpy())
    # tensor path: the tcgen05 kernel + the post kernel (fix-up of non-finite tiles, history)
    h512 = f32_taps(O.firdes_kaiser(512, 0.1, 80.0, 0.0))
    f = fir.FIRFilter(h512, 1.0)
    xs = torch.from_numpy(rand_cf32(rng, 1 << 18)).cuda()
    f.execute_block(xs)
    l0 = launch_count()
    f.execute_block(xs)
    assert f.last_path == "tensor" and launch_count() - l0 == 2


# ------------------------------------------------------------------ one stream over several handles (the multi-GPU seam)
@pytest.mark.parametrize("path", ["tensor", "ffma"])
def test_stream_split_over_two_handles(fir, torch, path, monkeypatch):
    """SURVEY 8(e) config 2 on one GPU: sgpu_shard_stream cuts the stream, handle r > 0 is primed with the T-1 samples
    in front of its segment (sgpu_fir_write) and runs its segment; the seam -- the first T-1 outputs of segment 1, which
    depend on the halo only -- and the interior must equal the oracle's unbroken stream."""
    from solid_dsp_b200 import sharding
    if path == "ffma":
        monkeypatch.setenv("SGPU_FIR_TC", "0")
    T = 512
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n = (1 << 22) + 4097
    rng = np.random.default_rng(8)
    x = rand_cf32(rng, n)
    xd = torch.from_numpy(x).cuda()
    world = 3
    outs = []
    for r in range(world):
        first, count = sharding.shard_stream(n, 1, world, r)
        f = fir.FIRFilter(h, 1.0)
        if r > 0:
            f.write(xd[first - (T - 1):first])
        y = f.execute_block(xd[first:first + count])
        assert f.last_path == path
        outs.append((first, count, y.cpu().numpy()))
    for first, count, y in outs:
        for s0 in (0, T - 1, count // 2, count - 2048):
            lo = max(0, first + s0 - (T - 1))
            ref = O.fir_fast(h, x[lo:first + s0 + 2048])[first + s0 - lo:]
            assert nerr(y[s0:s0 + 2048], ref[:len(y[s0:s0 + 2048])]) <= TOL, (first, s0)


def test_decimator_stream_split_over_two_handles(fir, torch):
    """Decimator segments start at multiples of M (sgpu_shard_stream's align), so every handle starts at phase 0."""
    from solid_dsp_b200 import sharding
    T, M = 256, 8
    h = f32_taps(O.firdes_kaiser(T, 0.5 / M * 0.9, 80.0, 0.0))
    n = (1 << 21) + 1003
    rng = np.random.default_rng(9)
    x = rand_cf32(rng, n)
    xd = torch.from_numpy(x).cuda()
    ref = O.fir_fast(h, x, 1.0, M)
    world = 2
    got = []
    for r in range(world):
        first, count = sharding.shard_stream(n, M, world, r)
        assert first % M == 0
        f = fir.DecimatingFIRFilter(h, 1.0, M)
        if r > 0:
            f.write(xd[first - (T - 1) - ((T - 1) % M and (M - (T - 1) % M)):first])  # a multiple of M samples: the phase stays 0
        got.append(f.execute_block(xd[first:first + count]).cpu().numpy())
    y = np.concatenate(got)
    assert y.shape == ref.shape
    assert nerr(y, ref) <= TOL
