"""CPU checks of the tensor-core formulation (csrc/fir_tc.cu), independent of the GPU:

* the banded-Toeplitz GEMM  y[128 b + m] = sum_k A[m][k] B[b][k]  with A[m][k] = tp[m mod L][m / L + Koff - k] and
  B[b][k] = x[R b - Koff + k] (R = 128 / L) reproduces the oracle's FIR (L = 1) and polyphase interpolator (L = 2, 4)
  -- the index arithmetic the host code uses to build the band and the TMA boxes use to fetch the rows;
* the operand splits carry f32 accuracy: BF16x3 with the six products b1h1, b1h2, b2h1, b2h2, b1h3, b3h1 and TF32x3 with
  hi*hi + lo*hi + hi*lo, evaluated in f64 so that only the split itself is under test (the accumulator's behaviour is
  measured on the GPU: tools/tc_accum_probe.py)."""
import numpy as np
import pytest

import oracle as O
from tests._util import f32_taps, nerr, rand_cf32


def _band(tp, L, S):
    """A [128][K] exactly as fir_tc_create_pfb builds it."""
    R = 128 // L
    koff = max(S - 1, 1)
    koff = (koff + 31) // 32 * 32
    K = koff + R
    A = np.zeros((128, K))
    for m in range(128):
        for k in range(K):
            j = m // L + koff - k
            if 0 <= j < S:
                A[m, k] = tp[m % L][j]
    return A, koff, R, K


def _rows(x, hist, n_blocks, koff, R, K):
    """B [n_blocks][K]: row b = x[R b - koff .. R b - koff + K) with the history in front and zeros behind."""
    ext = np.concatenate([np.zeros(koff, dtype=x.dtype), x, np.zeros(K, dtype=x.dtype)])
    if hist is not None and len(hist):
        ext[koff - len(hist):koff] = hist
    return np.stack([ext[R * b:R * b + K] for b in range(n_blocks)])


@pytest.mark.parametrize("T", [5, 64, 130, 512])
def test_fir_as_banded_gemm(T):
    rng = np.random.default_rng(T)
    h = f32_taps(rng.uniform(-1, 1, T))
    x = rand_cf32(rng, 1000).astype(np.complex128)
    tp = [h[::-1]]  # g[j] = h[T-1-j] (fir/mod.rs:86)
    A, koff, R, K = _band(tp, 1, T)
    nb = (len(x) + 127) // 128
    y = (_rows(x, None, nb, koff, R, K) @ A.T).reshape(-1)[:len(x)]
    assert nerr(y, O.fir_fast(h, x)) <= 1e-12


@pytest.mark.parametrize("L,T", [(2, 64), (4, 128), (4, 100), (2, 31), (4, 384)])
def test_interpolator_as_banded_gemm(L, T):
    rng = np.random.default_rng(10 * L + T)
    h = f32_taps(rng.uniform(-1, 1, T))
    x = rand_cf32(rng, 700).astype(np.complex128)
    S = -(-T // L)
    hpad = np.concatenate([h, np.zeros(S * L - T)])
    tp = [[hpad[p + (S - 1 - j) * L] for j in range(S)] for p in range(L)]  # pfb.rs:85-90, newest first
    A, koff, R, K = _band(tp, L, S)
    nb = (len(x) * L + 127) // 128
    y = (_rows(x, None, nb, koff, R, K) @ A.T).reshape(-1)[:len(x) * L]
    assert nerr(y, O.firinterp_fast(h, L, x)) <= 1e-12


def _bf16(a):
    """round-to-nearest-even to bfloat16, returned as f32 (cvt.rn.bf16.f32)"""
    u = np.asarray(a, dtype=np.float32).view(np.uint32)
    u = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return u.view(np.float32)


def _tf32(a):
    """round-to-nearest (ties away) to TF32 (cvt.rna.tf32.f32)"""
    u = np.asarray(a, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _parts(v, rnd, n):
    out, r = [], np.asarray(v, dtype=np.float32).copy()
    for _ in range(n):
        p = rnd(r)
        out.append(p.astype(np.float64))
        r = (r - p).astype(np.float32)  # exact in f32, as on the device
    return out


@pytest.mark.parametrize("T", [512, 2048])
def test_operand_splits_keep_f32_accuracy(T):
    rng = np.random.default_rng(T)
    h = np.asarray(O.firdes_kaiser(T, 0.1, 80.0, 0.0), dtype=np.float32)
    x = rng.uniform(-1, 1, 6000).astype(np.float32)
    ref = np.convolve(x.astype(np.float64), h.astype(np.float64))[:len(x)]
    den = np.max(np.abs(ref))
    hb, xb = _parts(h, _bf16, 3), _parts(x, _bf16, 3)
    y = sum(np.convolve(xb[j], hb[i])[:len(x)] for i, j in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)))
    assert np.max(np.abs(y - ref)) / den <= 3e-8  # dropped: b2 h3, b3 h2, b3 h3 (<= 2^-24 per product)
    # without the 2^-16 terms the split would NOT be f32-accurate: the six products are all needed
    y4 = sum(np.convolve(xb[j], hb[i])[:len(x)] for i, j in ((0, 0), (1, 0), (0, 1), (1, 1)))
    assert np.max(np.abs(y4 - ref)) / den > 3e-7
    ht, xt = _parts(h, _tf32, 2), _parts(x, _tf32, 2)
    y = sum(np.convolve(xt[j], ht[i])[:len(x)] for i, j in ((0, 0), (1, 0), (0, 1)))
    assert np.max(np.abs(y - ref)) / den <= 1e-7  # dropped: lo * lo (<= 2^-22 per product)


def _trunc_f32(v):
    """round toward zero to f32"""
    f = np.float32(v)
    if abs(float(f)) > abs(v):
        f = np.nextafter(f, np.float32(0))
    return f


def _dc_error(T, chain_ksteps):
    """Relative error of sum_k h[k] * 0.7 (positive Hann taps, TF32-exact operands: tools/tc_accum_probe.py's worst case)
    when an accumulator is truncated toward zero after every K = 8 step and chains of `chain_ksteps` steps are summed
    in f32 with round-to-nearest (chain_ksteps = None: one chain over the whole filter)."""
    h = _tf32((np.hanning(T + 2)[1:-1] / T).astype(np.float32))
    x = float(_tf32(np.float32(0.7)))
    prods = h.astype(np.float64) * x
    total, acc, cnt = np.float32(0), np.float32(0), 0
    for s in range((T + 7) // 8):
        acc = _trunc_f32(float(acc) + prods[8 * s:8 * s + 8].sum())
        cnt += 1
        if chain_ksteps and cnt == chain_ksteps:
            total, acc, cnt = np.float32(total + acc), np.float32(0), 0
    total = np.float32(total + acc)
    return (float(total) - prods.sum()) / prods.sum()


def test_truncating_accumulator_model():
    """Why the kernel sums short chains in registers: a round-toward-zero accumulator biases a same-sign sum by an
    amount that grows linearly with the number of sequential MMAs (measured on the B200: -3.1e-6 at 512 taps, -1.4e-5
    at 2048; this per-instruction model gives -1.2e-6 / -6.2e-6, the hardware drops a little more than one bit more),
    while chains of 64 taps keep it at the 1e-7 level whatever the filter length."""
    one = {T: _dc_error(T, None) for T in (512, 1024, 2048)}
    assert all(e < 0 for e in one.values())                      # the sum shrinks: truncation toward zero
    assert 1.7 < one[1024] / one[512] < 2.8 and 1.7 < one[2048] / one[1024] < 2.8  # linear in T
    for T in (256, 512, 1024, 2048, 4096):
        assert abs(_dc_error(T, 8)) <= 3e-7
