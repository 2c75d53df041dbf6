"""In-library multi-GPU context (sgpu_ctx_* / sgpu_sharded_*): channel ranges and time segments over several shards give
the results and the streaming state of the single handle.  On a one-GPU box the shards all sit on device 0 (a device may
appear more than once in a context); with more GPUs visible the same tests spread over them."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx3():
    import torch
    from solid_dsp_b200.context import Context
    n = torch.cuda.device_count()
    return Context([i % n for i in range(3)])


def test_context_devices(ctx3):
    import torch
    from solid_dsp_b200.context import Context
    assert ctx3.n_devices == 3
    assert Context().n_devices == torch.cuda.device_count()
    assert Context(1).n_devices == 1


@pytest.mark.parametrize("T,path", [(64, "ffma"), (512, "tensor")])
def test_fir_stream_segments(ctx3, T, path):
    """One FIR stream over three shards: halo from the caller's buffer, state handed back to shard 0 between calls."""
    rng = np.random.default_rng(T)
    h = f32_taps(O.firdes_kaiser(T, 0.1, 80.0, 0.0))
    n1, n2, n3 = 3 * (1 << 17) + 1234, 777, 1 << 18
    x = rand_cf32(rng, n1 + n2 + n3)
    f = ctx3.fir(h, 0.5)
    assert len(f.shards) == 3
    y1 = f.execute_block(x[:n1])
    assert f.last_segments == 3
    y2 = f.execute_block(x[n1:n1 + n2])          # too short to split: shard 0 alone, from the handed-over history
    assert f.last_segments == 1
    y3 = f.execute_block(x[n1 + n2:])
    assert f.last_segments == 3
    y = np.concatenate([y1, y2, y3])
    ref = O.fir_fast(h, x, 0.5)
    assert y.shape == ref.shape
    for lo in (0, n1 // 3 - 600, 2 * (n1 // 3) - 600, n1 - 600, n1 + n2 - 300, n1 + n2 + n3 // 3 - 600, x.size - 4096):
        lo = max(lo, 0)
        assert nerr(y[lo:lo + 4096], ref[lo:lo + 4096]) <= TOL, lo


@pytest.mark.parametrize("M,T", [(8, 256), (5, 41)])
def test_decimator_stream_segments_keep_the_phase(ctx3, M, T):
    rng = np.random.default_rng(M)
    h = f32_taps(O.firdes_kaiser(T, 0.4 / M, 60.0, 0.0))
    cuts = [0, 3, 3 + 3 * (1 << 17) + 11, 3 + 3 * (1 << 17) + 11 + 5, 3 + 6 * (1 << 17) + 29]
    x = rand_cf32(rng, cuts[-1])
    f = ctx3.fir(h, 1.0, decimation=M)
    outs = [f.execute_block(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    y = np.concatenate(outs)
    ref = O.fir_fast(h, x, 1.0, M)
    assert y.shape == ref.shape          # exact sample counts and phase alignment (decim.rs:221-228)
    assert nerr(y, ref) <= TOL


def test_channel_ranges(ctx3):
    rng = np.random.default_rng(9)
    Cn, n = 7, 40000
    x = rand_cf32(rng, (Cn, n))
    hd = f32_taps(O.firdes_kaiser(64, 0.05, 60.0, 0.0))
    d = ctx3.fir(hd, 1.0, n_channels=Cn, decimation=4)
    assert [s[1:] for s in d.shards] == [(0, 3), (3, 2), (5, 2)]
    yd = np.concatenate([d.execute_block(x[:, :12345]), d.execute_block(x[:, 12345:])], axis=1)
    hi = f32_taps(O.firdes_kaiser(24, 0.1, 60.0, 0.0))
    it = ctx3.interpolator(hi, 3, n_channels=Cn)
    yi = it.execute_block(x[:, :5000])
    from solid_dsp_b200.filter import iirdes
    from solid_dsp_b200.filter.iir import IIRFilterType
    ff, fb = (f32_taps(v) for v in iirdes.stable_lowpass_sections(4))
    for c in range(Cn):
        assert nerr(yd[c], O.fir_fast(hd, x[c], 1.0, 4)) <= TOL
        assert nerr(yi[c], O.firinterp_fast(hi, 3, x[c, :5000])) <= TOL
    q = ctx3.iir(ff, fb, IIRFilterType.SecondOrder, n_channels=Cn)
    yq = np.concatenate([q.execute_block(x[:, :999]), q.execute_block(x[:, 999:20000])], axis=1)
    for c in range(Cn):
        ref, _ = O.sos_cascade_fast(ff, fb, x[c, :20000])
        assert nerr(yq[c], ref) <= TOL
    # fewer channels than devices: the empty shards are not created
    assert len(ctx3.fir(hd, 1.0, n_channels=2).shards) == 2
