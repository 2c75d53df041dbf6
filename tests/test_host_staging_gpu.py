"""SGPU_HOST calls on PAGEABLE caller memory (ordinary numpy arrays -- what the reference's callers hand over): from 4 MiB
on the library gathers / scatters through its own pinned staging buffers with host threads instead of leaving it to the
driver's synchronous staged copies (csrc/sgpu_common.cuh: host_pipeline_body).  Many small chunks (SGPU_HOST_CHUNK_MB=1)
exercise every buffer hand-over; results must equal the device-pointer path and the oracle."""
import numpy as np
import pytest

import oracle as O
from tests._util import TOL, f32_taps, nerr, rand_cf32

pytestmark = pytest.mark.gpu


@pytest.fixture()
def small_chunks(monkeypatch):
    monkeypatch.setenv("SGPU_HOST_CHUNK_MB", "1")
    monkeypatch.setenv("SGPU_HOST_STAGING", "1")


@pytest.mark.parametrize("threads", ["1", "5"])
def test_pageable_fir_family(small_chunks, monkeypatch, threads):
    import torch
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter, FIRFilter, InterpolatingFIRFilter
    monkeypatch.setenv("SGPU_HOST_COPY_THREADS", threads)
    rng = np.random.default_rng(41)
    C, n = 3, 700_001                      # 16.8 MB in: staged; 1 MiB chunks: 17 chunks, ragged last one
    x = rand_cf32(rng, (C, n))
    h = f32_taps(O.firdes_kaiser(64, 0.2, 60.0, 0.0))
    for make, ref in ((lambda: FIRFilter(h, 0.5, n_channels=C), lambda c: O.fir_fast(h, x[c], 0.5)),
                      (lambda: DecimatingFIRFilter(h, 1.0, 5, n_channels=C), lambda c: O.fir_fast(h, x[c], 1.0, 5)),
                      (lambda: InterpolatingFIRFilter(h, 3, n_channels=C), lambda c: O.firinterp_fast(h, 3, x[c]))):
        y_host = np.asarray(make().execute_block(x))                         # pageable in, fresh pageable out
        y_dev = make().execute_block(torch.from_numpy(x).cuda()).cpu().numpy()
        assert y_host.shape == y_dev.shape
        assert nerr(y_host, y_dev) <= 1e-6
        for c in (0, C - 1):
            assert nerr(y_host[c], ref(c)) <= TOL
    # split calls keep streaming state across staged calls
    f = DecimatingFIRFilter(h, 1.0, 5, n_channels=C)
    y = np.concatenate([np.asarray(f.execute_block(x[:, :600_003])), np.asarray(f.execute_block(x[:, 600_003:]))], axis=1)
    assert nerr(y[1], O.fir_fast(h, x[1], 1.0, 5)) <= TOL


def test_pageable_iir_and_strided_rows(small_chunks):
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    rng = np.random.default_rng(42)
    ff, fb = (f32_taps(v) for v in stable_lowpass_sections(4))
    C, n = 2, 600_000
    x = rand_cf32(rng, (C, n))
    y = np.asarray(IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=C).execute_block(x))
    for c in range(C):
        assert nerr(y[c], O.sos_cascade_fast(ff, fb, x[c])[0]) <= TOL


def test_pinned_and_pageable_mixes(small_chunks):
    """Pinned input with pageable output and the reverse: each side chooses its own path."""
    from solid_dsp_b200.filter.fir import FIRFilter
    from solid_dsp_b200.hostmem import PinnedArray
    rng = np.random.default_rng(43)
    n = 1_000_003
    h = f32_taps(O.firdes_kaiser(48, 0.2, 60.0, 0.0))
    xp = PinnedArray(1, n)
    xp.array[0] = rand_cf32(rng, (n,))
    ref = O.fir_fast(h, xp.array[0].astype(np.complex128))
    y1 = np.asarray(FIRFilter(h, 1.0).execute_block(xp.array[0]))            # pinned in (numpy view), pageable out
    y2 = np.asarray(FIRFilter(h, 1.0).execute_block(xp.array[0].copy()))     # pageable in, pageable out
    assert nerr(y1, ref) <= TOL and np.array_equal(y1, y2)
