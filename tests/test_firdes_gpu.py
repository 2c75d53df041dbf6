"""Tap design on the device (sgpu_firdes_kaiser, csrc/firdes.cu) against the host restatement of
firdes/mod.rs:278-305, which is pinned by the reference's f32-rounded goldens (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

# f64 on both sides, the same operations in the same order: only the last bits of log / exp / sin / cos differ between
# the device's math library and the host's.  The 64-term series accumulates a few hundred of them.
RTOL = 1e-12


@pytest.fixture(scope="module")
def firdes():
    from solid_dsp_b200.filter import firdes
    return firdes


@pytest.mark.parametrize("n,fc,att,mu", [
    (64, 0.25, 60.0, 0.0),                 # BASELINE configs[0]
    (512, 0.1, 80.0, 0.0),                 # configs[1]
    (256, 0.5 / 8 * 0.9, 80.0, 0.0),       # configs[2]
    (128, 0.5 / 4 * 0.9, 80.0, 0.0),       # configs[3]
    (8, 0.35, 120.0, 0.0),                 # firdes/mod.rs:271 doc-test
    (33, 0.2, 40.0, 0.3), (33, 0.2, 15.0, -0.5), (1, 0.1, 60.0, 0.0), (2, 0.5, 60.0, 0.5), (4097, 0.01, 100.0, 0.0),
])
def test_device_design_matches_the_host_design(firdes, n, fc, att, mu):
    dev = np.array(firdes.firdes_kaiser_device(n, fc, att, mu))
    host = np.array(firdes.firdes_kaiser(n, fc, att, mu))
    orc = np.array(O.firdes_kaiser(n, fc, att, mu))
    assert dev.shape == host.shape == (n,)
    if n > 1:  # n = 1: (n - 1) = 0 in windows/kaiser.rs:43 -> 0/0 = NaN in the reference, and here
        assert np.array_equal(host, orc)
        scale = np.max(np.abs(host))
        assert np.max(np.abs(dev - host)) <= RTOL * scale
        # the f32 taps the filters are built from are the same (bar one-ulp ties)
        assert np.max(np.abs(dev.astype(np.float32) - host.astype(np.float32))) <= 1.2e-7 * scale
    else:
        assert np.isnan(dev).all() and np.isnan(host).all()


def test_reference_golden_through_the_device_design(firdes):
    """The reference's doc-test that consumes a firdes_kaiser design (firdes/mod.rs:470-485: cross-correlation with a notch
    filter, compared as f32) holds when the Kaiser design comes from the device."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "reference_doctests.json").read_text())["firdes_crosscorrelation"]
    h = firdes.firdes_kaiser_device(*g["kaiser"])
    n = firdes.firdes_notch(*g["notch"])
    assert np.float32(firdes.filter_crosscorrelation(h, n, g["lag"])) == np.float32(g["expect"])


def test_bank_of_designs_in_one_launch_feeds_per_channel_filters(firdes):
    import torch
    from solid_dsp_b200.filter.fir import FIRFilter
    C, T = 37, 96
    fcs = [0.05 + 0.4 * c / C for c in range(C)]
    bank = firdes.firdes_kaiser_device(T, fcs, 70.0, 0.0)
    assert len(bank) == C and all(len(r) == T for r in bank)
    for c in (0, 5, C - 1):
        host = np.array(firdes.firdes_kaiser(T, fcs[c], 70.0, 0.0))
        assert np.max(np.abs(np.array(bank[c]) - host)) <= RTOL * np.max(np.abs(host))
    # device-resident output: designed and kept on the GPU
    d = torch.empty((C, T), dtype=torch.float64, device="cuda")
    firdes.firdes_kaiser_device(T, fcs, 70.0, 0.0, device_out=d)
    assert np.max(np.abs(d.cpu().numpy() - np.array(bank))) == 0.0
    # ... and the bank drives C independent filter objects (fir/mod.rs:79-88: every object owns its taps)
    rng = np.random.default_rng(3)
    x = (rng.uniform(-1, 1, (C, 2000)) + 1j * rng.uniform(-1, 1, (C, 2000))).astype(np.complex64)
    f = FIRFilter(np.array(bank), 1.0)
    y = f.execute_block(torch.from_numpy(x).cuda()).cpu().numpy()
    for c in (0, 11, C - 1):
        taps32 = np.array(bank[c], dtype=np.float32).astype(np.float64)
        ref = np.array(O.FIRFilter(taps32, 1.0).execute_block(x[c].astype(np.complex128)))
        assert np.max(np.abs(y[c] - ref)) / np.max(np.abs(ref)) <= 1e-5


def test_errors_follow_the_reference_order(firdes):
    for args, code in [((8, 0.2, 60.0, 0.6), "Mu"), ((8, 0.6, 60.0, 0.0), "Bandwidth"), ((8, 0.2, 0.0, 0.0), "StopBandLevel"),
                       ((8, 0.7, -1.0, 0.9), "Mu"), ((8, 0.7, -1.0, 0.0), "Bandwidth"), ((8, float("nan"), 60.0, 0.0), "Bandwidth")]:
        with pytest.raises(firdes.FirdesError) as e:
            firdes.firdes_kaiser_device(*args)
        assert str(e.value) == code
        with pytest.raises(firdes.FirdesError) as e2:
            firdes.firdes_kaiser(*args)
        assert str(e2.value) == code
    assert firdes.firdes_kaiser_device(0, 0.2, 60.0, 0.0) == []


def test_filter_energy_with_the_dot_products_on_the_gpu(firdes):
    """firdes::filter_energy is the reference crate's own caller of DotProduct::execute (firdes/mod.rs:620-629): 128
    sample vectors e^{j 2 pi f k} against FORWARD coefficients.  Here they are one batched sgpu_dot_execute; the f64 host
    form reproduces the reference's golden (0.3152318 as f32) exactly, the device form within f32 rounding."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "reference_doctests.json").read_text())["firdes_energy"]
    h = firdes.firdes_notch(*g["notch"])
    host = firdes.filter_energy(h, g["cutoff"], g["fft_size"])
    assert np.float32(host) == np.float32(g["expect"])
    dev = firdes.filter_energy_device(h, g["cutoff"], g["fft_size"])
    assert abs(dev - host) <= 1e-5 * host
    # a long filter and a fine grid: 4096 dot products of 2049 terms in one call; the cut-off sits inside the pass band so
    # that the ratio is not dominated by the f32 rounding of an 80 dB stop band
    h2 = firdes.firdes_kaiser(2049, 0.1, 80.0, 0.0)
    want = O.filter_energy(h2, 0.05, 4096)
    assert 0.3 < want < 0.7
    assert abs(firdes.filter_energy_device(h2, 0.05, 4096) - want) <= 1e-5 * want
