"""GPU exploration helper (not part of the product): FP32 peaks and raw kernel timings.
usage: python tools/explore.py [peaks] [fir] [decim] [interp] [iir]"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from solid_dsp_b200 import _ffi  # noqa: E402


def ev_time(fn, reps=None, warm=None):
    import os
    reps = int(os.environ.get('EXPLORE_REPS', 5)) if reps is None else reps
    warm = int(os.environ.get('EXPLORE_WARM', 2)) if warm is None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def peaks():
    P = _ffi.peak_lib()
    names = {0: "ffma_scalar_3reg", 1: "ffma2_packed", 2: "ffma2_fir_shape", 3: "ffma_scalar_fir_shape"}
    for v, nm in names.items():
        for bps in (2, 4, 8):
            ms, fl = C.c_double(), C.c_double()
            rc = P.sgpu_peak_fma(v, bps, 2000, 5, C.byref(ms), C.byref(fl))
            print(json.dumps({"peak": nm, "blocks_per_sm": bps, "rc": rc, "ms": ms.value,
                              "tflops": fl.value / ms.value / 1e9}))
    ms, by = C.c_double(), C.c_double()
    P.sgpu_peak_copy(4 << 30, 5, C.byref(ms), C.byref(by))
    print(json.dumps({"peak": "copy", "ms": ms.value, "GBps": by.value / ms.value / 1e6}))


def occ():
    """FMA throughput vs resident warps per SM (128-thread blocks)."""
    P = _ffi.peak_lib()
    for v, nm in ((1, "ffma2_packed"), (2, "ffma2_fir_shape"), (0, "ffma_scalar")):
        for bps in (1, 2, 3, 4, 6, 8):
            ms, fl = C.c_double(), C.c_double()
            P.sgpu_peak_fma_ex(v, bps, 128, 2000, 3, C.byref(ms), C.byref(fl))
            print(json.dumps({"peak": nm, "warps_per_sm": bps * 4, "tflops": round(fl.value / ms.value / 1e9, 2)}))


def fir(T=512, logn=26):
    from solid_dsp_b200.filter.fir import FIRFilter
    n = 1 << logn
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    h = np.random.default_rng(0).uniform(-1, 1, T)
    f = FIRFilter(h, 1.0)
    best, med = ev_time(lambda: f.execute_block(x))
    print(json.dumps({"kernel": "fir", "T": T, "n": n, "best_ms": best, "med_ms": med,
                      "Gsamp_s": n / best / 1e6, "tflops": n * T * 4 / best / 1e9}))


def decim(T=256, M=8, Cn=256, logn=20):
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter
    n = 1 << logn
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    h = np.random.default_rng(0).uniform(-1, 1, T)
    f = DecimatingFIRFilter(h, 1.0, M, n_channels=Cn)
    best, med = ev_time(lambda: f.execute_block(x))
    tot = Cn * n
    print(json.dumps({"kernel": "decim", "T": T, "M": M, "C": Cn, "n": n, "best_ms": best, "med_ms": med,
                      "Gsamp_in_s": tot / best / 1e6, "tflops": tot / M * T * 4 / best / 1e9,
                      "GBps": tot * 9 / best / 1e6}))


def interp(T=128, L=4, Cn=256, logn=18):
    from solid_dsp_b200.filter.fir import InterpolatingFIRFilter
    n = 1 << logn
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    h = np.random.default_rng(0).uniform(-1, 1, T)
    f = InterpolatingFIRFilter(h, L, n_channels=Cn)
    best, med = ev_time(lambda: f.execute_block(x))
    tot = Cn * n * L
    print(json.dumps({"kernel": "interp", "T": T, "L": L, "C": Cn, "n": n, "best_ms": best, "med_ms": med,
                      "Gsamp_out_s": tot / best / 1e6, "tflops": tot * (T / L) * 4 / best / 1e9,
                      "GBps": tot * 10 / best / 1e6}))


def iir(Cn=65536, logn=12, nsec=8):
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    n = 1 << logn
    ff, fb = stable_lowpass_sections(nsec)
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    f = IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=Cn)
    best, med = ev_time(lambda: f.execute_block(x))
    tot = Cn * n
    print(json.dumps({"kernel": "iir", "C": Cn, "n": n, "nsec": nsec, "best_ms": best, "med_ms": med,
                      "Gsamp_s": tot / best / 1e6, "GBps": tot * 16 / best / 1e6}))


def iirn(Cn=65536, n=16400, nsec=8):
    """batch IIR with an arbitrary (non power of two) row length"""
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    ff, fb = stable_lowpass_sections(nsec)
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    f = IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=Cn)
    best, med = ev_time(lambda: f.execute_block(x))
    tot = Cn * n
    print(json.dumps({"kernel": "iirn", "C": Cn, "n": n, "nsec": nsec, "best_ms": best, "med_ms": med,
                      "Gsamp_s": tot / best / 1e6, "GBps": tot * 16 / best / 1e6}))


def iirwrap(Cn=65536, logn=12, factor=4, nsec=8):
    """decimating / interpolating IIR wrappers (iir/decim.rs, iir/interp.rs)"""
    from solid_dsp_b200.filter.iir import DecimatingIIRFilter, InterpolatingIIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    n = 1 << logn
    ff, fb = stable_lowpass_sections(nsec)
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    f = DecimatingIIRFilter(ff, fb, IIRFilterType.SecondOrder, factor, n_channels=Cn)
    best, _ = ev_time(lambda: f.execute_block(x))
    print(json.dumps({"kernel": "iir_decim", "C": Cn, "n": n, "M": factor, "best_ms": best, "Gsamp_in_s": Cn * n / best / 1e6}))
    xi = x[:, : n // factor].contiguous()
    g = InterpolatingIIRFilter(ff, fb, IIRFilterType.SecondOrder, factor, n_channels=Cn)
    best, _ = ev_time(lambda: g.execute_block(xi))
    print(json.dumps({"kernel": "iir_interp", "C": Cn, "n_in": n // factor, "L": factor, "best_ms": best,
                      "Gsamp_out_s": Cn * n / best / 1e6}))


def autocorr(W=64, d=16, Cn=1024, logn=20):
    from solid_dsp_b200.filter.auto_correlator import AutoCorrelator
    n = 1 << logn
    x = torch.randn((Cn, n), dtype=torch.complex64, device="cuda")
    f = AutoCorrelator(W, d, n_channels=Cn)
    best, med = ev_time(lambda: f.execute_block(x))
    print(json.dumps({"kernel": "autocorr", "W": W, "d": d, "C": Cn, "n": n, "best_ms": best,
                      "Gsamp_s": Cn * n / best / 1e6, "GBps": Cn * n * 16 / best / 1e6}))


def iirscan(logn=26, nsec=8):
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
    from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
    n = 1 << logn
    ff, fb = stable_lowpass_sections(nsec)
    x = torch.randn(n, dtype=torch.complex64, device="cuda")
    f = IIRFilter(ff, fb, IIRFilterType.SecondOrder)
    best, med = ev_time(lambda: f.execute_block(x))
    print(json.dumps({"kernel": "iirscan", "n": n, "nsec": nsec, "best_ms": best, "med_ms": med,
                      "Gsamp_s": n / best / 1e6, "GBps": n * 16 / best / 1e6}))


if __name__ == "__main__":
    what = sys.argv[1:] or ["peaks", "fir"]
    print(json.dumps({"device": torch.cuda.get_device_name(0), "info": str(_ffi.device_info())}))
    for w in what:
        t0 = time.time()
        if ":" in w:
            name, args = w.split(":", 1)
            globals()[name](*[int(a) for a in args.split(",")])
        else:
            globals()[w]()
        print(f"# {w} took {time.time() - t0:.1f}s", flush=True)
