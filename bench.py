#!/usr/bin/env python
"""bench.py -- throughput of the filtering hot path on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fir|fir64|decim|interp|iir_batch|iir_scan|autocorr|ddc]
    python bench.py --impl reference ...        # the reference's CPU path (restated oracle) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...   (one rank per GPU)

Default workload = BASELINE.json configs[1]: one complex stream of 2^30 samples through a 512-tap
real-coefficient Kaiser low-pass (firdes_kaiser(512, 0.1, 80 dB)), contiguous segments sharded over
the ranks with a 511-sample halo (strong scaling: total work fixed).  A step is one pass over the
whole stream: halo write + FIR execute_block on every rank.

`value`   : complex Gsamples/s with inputs resident in HBM (device pointers through the C ABI).
`e2e`     : same metric through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region.
`roofline`: the dominant kernel -- timed with CUDA events around execute_block INSIDE the timed loop that
            gives `value` -- against the binding peak: FP32 FMA measured live (libsgpu_peakbench), the measured
            HBM copy peak or the measured bf16 tensor rate (MEASURED_PEAKS.json).
`cpu_baseline`: the restated reference CPU path (oracle/, structural: Window memmove + to_vec +
            sequential DotProduct per sample) on the box's host cores, bounded sample.
`workloads`: (default run at N = 1 only) the other BASELINE configs, each with value, kernel time, both roofline
            fractions, oracle parity, clocks and a 1-core cpu_baseline.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import gc
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "complex Gsamples/s"
UNIT = "Gsamples/s"

WORKLOADS = {
    # name: description, flop per unit, bytes per unit, unit
    "fir": dict(desc="BASELINE configs[1]: single-stream 512-tap FIR (real Kaiser taps, complex f32 samples) over 2^30 samples",
                flop_per_unit=2048.0, bytes_per_unit=16.0, unit="input sample"),
    "fir64": dict(desc="BASELINE configs[0]: single-channel 64-tap windowed-sinc low-pass over 2^20 samples per call (the reference's CPU-runnable case)",
                  flop_per_unit=256.0, bytes_per_unit=16.0, unit="input sample"),
    "decim": dict(desc="BASELINE configs[2]: decimator M=8, 256 taps, 4096 channels x 2^20 samples",
                  flop_per_unit=128.0, bytes_per_unit=9.0, unit="input sample"),
    "decim_pc": dict(desc="SURVEY 8d: BASELINE configs[2] with per-channel taps -- 4096 different 256-tap Kaiser designs (computed on the device), M=8, 4096 channels x 2^20 samples",
                     flop_per_unit=128.0, bytes_per_unit=9.0, unit="input sample"),
    "interp": dict(desc="BASELINE configs[3]: interpolator L=4, 128 taps, 1024 channels x 2^20 inputs",
                   flop_per_unit=128.0, bytes_per_unit=10.0, unit="output sample"),
    "iir_batch": dict(desc="BASELINE configs[4a]: 8-section biquad cascade, 65536 channels x 2^14 samples",
                      flop_per_unit=144.0, bytes_per_unit=16.0, unit="sample"),
    "iir_scan": dict(desc="BASELINE configs[4b]: 8-section biquad cascade, one stream of 2^28 samples (chunked scan)",
                     flop_per_unit=144.0, bytes_per_unit=16.0, unit="sample"),
    # widening, not BASELINE configs
    "autocorr": dict(desc="SURVEY 8f: AutoCorrelator window 64, delay 16, 1024 channels x 2^20 samples",
                     flop_per_unit=16.0, bytes_per_unit=16.0, unit="sample"),
    "ddc": dict(desc="SURVEY 8f: NCO mix-down fused into the decimator (DDC): M=8, 256 taps, 4096 channels x 2^20 samples",
                flop_per_unit=136.0, bytes_per_unit=9.0, unit="input sample"),
}
AUTOCORR_SHAPE = (64, 16)
DDC_FREQ = 0.1234  # radians per sample of the mix-down NCO
FIR64_ROWS = 32    # fir64: the calls rotate over 32 buffers of 2^20 samples (512 MB in + out: every call reads cold data)
SHAPES = {  # channels, log2(samples per channel)
    "decim": (4096, 20), "interp": (1024, 20), "iir_batch": (65536, 14), "iir_scan": (1, 28), "autocorr": (1024, 20),
    "ddc": (4096, 20), "decim_pc": (4096, 20),
}

# (pole radius, pole angle / pi) of the benchmark's biquad sections: conjugate pole pairs, double zero at z = -1, unit DC
# gain (the reference has no SOS designer); the same table as solid_dsp_b200.filter.iirdes, kept
# here so that the reference arm imports nothing of the product
_SECTION_TABLE = [(0.50, 0.10), (0.60, 0.14), (0.70, 0.18), (0.78, 0.22), (0.84, 0.26), (0.88, 0.30), (0.92, 0.34), (0.95, 0.38)]


def f32_taps(h):
    return np.asarray(h, dtype=np.float32).astype(np.float64)


def bench_sections():
    ff, fb = [], []
    for r, th in _SECTION_TABLE:
        a1 = -2.0 * r * math.cos(math.pi * th)
        a2 = r * r
        g = (1.0 + a1 + a2) / 4.0
        ff += [g, 2.0 * g, g]
        fb += [1.0, a1, a2]
    return f32_taps(ff), f32_taps(fb)


def workload_taps(name, impl="ours"):
    """Taps of a workload.  The reference arm designs them with the oracle's restatement of firdes_kaiser (bit-identical
    to the product's, tests/test_oracle_golden.py) so that process never loads the product library."""
    if impl == "reference":
        import oracle as O
        kaiser = O.firdes_kaiser
    else:
        from solid_dsp_b200.filter import firdes
        kaiser = firdes.firdes_kaiser
    if name == "fir":
        return f32_taps(kaiser(512, 0.1, 80.0, 0.0))
    if name == "fir64":
        return f32_taps(kaiser(64, 0.25, 60.0, 0.0))
    if name in ("decim", "ddc"):
        return f32_taps(kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
    if name == "decim_pc":
        # one design per channel: cut-offs spread over [0.5, 0.9] of the decimated band.  Ours: all 4096 designs in one
        # launch of sgpu_firdes_kaiser; the reference arm only needs one of them (the cost does not depend on the values)
        chans = SHAPES[name][0]
        fcs = [0.5 / 8 * (0.5 + 0.4 * c / chans) for c in range(chans)]
        if impl == "reference":
            return f32_taps(kaiser(256, fcs[0], 80.0, 0.0))
        return f32_taps(firdes.firdes_kaiser_device(256, fcs, 80.0, 0.0))
    if name == "interp":
        return f32_taps(kaiser(128, 0.5 / 4 * 0.9, 80.0, 0.0))
    if name == "autocorr":
        return AUTOCORR_SHAPE
    return bench_sections()


def workload_config(name, world, log2_samples):
    """The `config` object: identical in both arms."""
    W = WORKLOADS[name]
    if name == "fir":
        shape = {"samples": 1 << log2_samples, "taps": 512, "segments": world, "halo": 511}
    elif name == "fir64":
        shape = {"samples_per_call": 1 << 20, "taps": 64, "buffers": FIR64_ROWS}
    else:
        chans, lg = SHAPES[name]
        if log2_samples != 30:
            lg = log2_samples
        shape = {"channels": chans, "samples_per_channel": 1 << lg}
    l2 = ("the calls rotate over 32 input / output buffer pairs (512 MB): every call reads data that is not in the 126 MB L2"
          if name == "fir64" else "inputs (>= 2 GiB per rank) exceed the 126 MB L2; no explicit flush")
    par = (f"stream segments x{world} with halo" if name in ("fir", "iir_scan") else
           ("replicas" if name == "fir64" else f"channels x{world}"))
    return {"workload": W["desc"], "name": name, **shape, "l2": l2, "parallelism": par}


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def mark(self):
        """Evaluate only the samples delivered from now on."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs a few hundred ms to start: wait until it has delivered a sample."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# --------------------------------------------------------------------------------- CPU baseline
def cpu_reference_path(name: str, n_threads: int, budget_s: float, native: bool = True):
    """Time the restated reference CPU path (structural oracle objects) on a bounded sample of the
    workload.  Returns (units_per_second, sample description, seconds)."""
    import oracle as O
    rng = np.random.default_rng(1234)
    taps = workload_taps(name, "reference")

    def run(n_units, n_in):
        if name in ("fir", "fir64"):
            T = len(taps)
            # stream segments, each primed with T-1 preroll samples through the public API
            x = (rng.uniform(-1, 1, n_units * n_in + T) + 1j * rng.uniform(-1, 1, n_units * n_in + T))
            out = np.zeros(n_units * n_in, dtype=np.complex128)
            t0 = time.perf_counter()
            O.run_units("fir", x, n_in, n_units, n_in, out, n_in, n_threads, coefs=taps, preroll=T - 1,
                        x_offset=T - 1, native=native)
            return time.perf_counter() - t0, n_units * n_in
        if name in ("decim", "ddc", "decim_pc"):
            x = rng.uniform(-1, 1, n_units * n_in) + 1j * rng.uniform(-1, 1, n_units * n_in)
            out = np.zeros(n_units * (n_in // 8 + 1), dtype=np.complex128)
            t0 = time.perf_counter()
            if name == "ddc":  # NCO::mix_down per sample (nco/mod.rs:141-172 semantics, oracle restatement) in front of the decimator
                x = O.nco_mix_down_block(x.reshape(n_units, n_in), DDC_FREQ, n_threads=n_threads).reshape(-1)
            O.run_units("decim", x, n_in, n_units, n_in, out, n_in // 8 + 1, n_threads, coefs=taps, decimation=8,
                        native=native)
            return time.perf_counter() - t0, n_units * n_in
        if name == "interp":
            x = rng.uniform(-1, 1, n_units * n_in) + 1j * rng.uniform(-1, 1, n_units * n_in)
            out = np.zeros(n_units * (n_in * 4 + 1), dtype=np.complex128)
            t0 = time.perf_counter()
            O.run_units("interp", x, n_in, n_units, n_in, out, n_in * 4 + 1, n_threads, coefs=taps, interpolation=4,
                        native=native)
            return time.perf_counter() - t0, n_units * n_in * 4
        if name == "autocorr":
            # structural oracle objects, one per unit, on a thread pool (ctypes releases the GIL)
            from concurrent.futures import ThreadPoolExecutor
            x = rng.uniform(-1, 1, (n_units, n_in)) + 1j * rng.uniform(-1, 1, (n_units, n_in))
            objs = [O.AutoCorrelator(*taps, native=native) for _ in range(n_units)]
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max(n_threads, 1)) as ex:
                list(ex.map(lambda i: objs[i].execute_block(x[i]), range(n_units)))
            return time.perf_counter() - t0, n_units * n_in
        ff, fb = taps
        x = rng.uniform(-1, 1, n_units * n_in) + 1j * rng.uniform(-1, 1, n_units * n_in)
        out = np.zeros(n_units * n_in, dtype=np.complex128)
        t0 = time.perf_counter()
        O.run_units("iir", x, n_in, n_units, n_in, out, n_in, n_threads, ff=ff, fb=fb, native=native)
        return time.perf_counter() - t0, n_units * n_in

    units = max(n_threads, 1)
    n_in = 1 << 14
    dt, done = run(units, n_in)  # calibration (also warms the allocator)
    rate = done / dt
    per_unit_rate = rate / (4 if name == "interp" else 1)  # input samples per second
    n_in = int(max(1 << 14, min(1 << 24, budget_s * per_unit_rate / units)))
    n_in -= n_in % 64
    dt, done = run(units, n_in)
    what = (f"{units} independent {name} objects x {n_in} input samples each, {n_threads} thread(s), "
            f"f64 structural oracle (-O3 -march={'native' if native else 'x86-64-v3'})")
    return done / dt, what, dt


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (restated oracle; the
    reference is Rust and cannot be built here), all host threads, bounded sample per step.  This process imports
    oracle/ only -- never the product package or its library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    name = args.workload
    per_step_budget = max(1.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    rates = []
    what = ""
    t_total = 0.0
    for i in range(args.warmup + args.steps):
        rate, what, dt = cpu_reference_path(name, cores, per_step_budget)
        if i >= args.warmup:
            rates.append(rate)
            t_total += dt
    value = float(np.mean(rates)) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, world, args.log2_samples),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- GPU arm
_PEAKS = None


def measure_peaks():
    global _PEAKS
    if _PEAKS is None:
        from solid_dsp_b200 import _ffi
        P = _ffi.peak_lib()
        best = 0.0
        detail = {}
        for variant, nm in ((0, "ffma_scalar"), (1, "ffma2_packed")):
            ms, fl = C.c_double(), C.c_double()
            if P.sgpu_peak_fma(variant, 8, 400, 3, C.byref(ms), C.byref(fl)) == 0:
                tf = fl.value / ms.value / 1e9
                detail[nm] = tf
                best = max(best, tf)
        _PEAKS = (best, detail)
    return _PEAKS


def file_peaks():
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    out = {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "bf16": 1590.0, "bf16_src": "fallback (B200_PROFILING.md)"}
    if peaks_file.exists():
        pk = json.loads(peaks_file.read_text())
        out["hbm"] = float(pk["hbm_gbs"])
        out["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        out["bf16"] = float(pk.get("bf16_tflops_sustained") or pk["bf16_tflops"])
        out["bf16_src"] = ("measured (MEASURED_PEAKS.json, bf16_tflops_sustained: the kernel runs for several ms back to "
                           "back under the power cap)")
    return out


class _LazyRaw(dict):
    """NCO words of a channel, read on demand -- valid because every channel of the DDC bench steps alike: the words
    read AFTER the step are moved back by the step's length."""

    def __init__(self, nco):
        super().__init__()
        self.nco = nco
        self.rewind = 0

    def __missing__(self, c):
        th, dl = self.nco.raw(c)
        return ((th - self.rewind * dl) & 0xFFFFFFFF, dl)


class Workload:
    """One workload on this rank: inputs in HBM, the filter handle, step() and the timed loop."""

    def __init__(self, name, args, world, rank, dev, dist):
        import torch
        from solid_dsp_b200 import sharding
        from solid_dsp_b200.filter.fir import DecimatingFIRFilter, FIRFilter, InterpolatingFIRFilter
        from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
        self.name, self.args, self.world, self.rank, self.dev, self.dist = name, args, world, rank, dev, dist
        self.torch = torch
        self.taps = workload_taps(name)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + rank)

        def rand_c(shape):
            t = torch.empty(tuple(shape) + (2,), dtype=torch.float32, device=dev)
            t.uniform_(-1.0, 1.0, generator=gen)
            return torch.view_as_complex(t)

        self.halo_prev = None
        self.seg_first = 0
        self.kernel_events = []
        taps = self.taps
        if name == "fir":
            n_total = 1 << args.log2_samples
            self.seg_first, n_loc = sharding.shard_stream(n_total, 1, world, rank)
            self.x = rand_c((n_loc,))
            self.filt = FIRFilter(taps, 1.0)
            T = len(taps)
            self.halo_prev = torch.zeros(T - 1, dtype=torch.complex64, device=dev)
            self.units_total = n_total

            def step():
                # halo: last T-1 samples of the previous rank's segment (rank 0: zeros = fresh filter)
                sharding.exchange_halo(self.x, self.halo_prev, rank, world, dist)
                self.filt.write(self.halo_prev)
                return self._timed(lambda: self.filt.execute_block(self.x))
        elif name == "fir64":
            self.x = rand_c((FIR64_ROWS, 1 << 20))
            self.filt = FIRFilter(taps, 1.0)
            self.units_total = (1 << 20) * world  # every rank runs its own replica of the single-channel case
            self.it = 0
            self.out = torch.empty_like(self.x)
            from solid_dsp_b200 import _ffi
            got = _ffi.c_size()
            stream = torch.cuda.current_stream(dev).cuda_stream
            n = 1 << 20

            # straight through the C ABI with device pointers, as a C / Rust caller would: arguments prepared once, no
            # per-call output allocation, no per-call CUDA events (the step IS the call: kernel time = step time)
            fn, h, gotp = _ffi.lib.sgpu_fir_execute_block, self.filt._h, C.byref(got)
            calls = [(self.x[r].data_ptr(), self.out[r].data_ptr()) for r in range(FIR64_ROWS)]

            def step():
                xi, yi = calls[self.it % FIR64_ROWS]
                self.it += 1
                st = fn(h, xi, n, n, yi, n, gotp, _ffi.DEVICE, stream)
                if st:
                    _ffi.check(st)
                return self.out[(self.it - 1) % FIR64_ROWS]
        else:
            chans, lg = SHAPES[name]
            n_per = 1 << (lg if args.log2_samples == 30 else args.log2_samples)
            if name == "iir_scan":
                c_loc = 1  # one stream: time segments per rank, warm-up halo from the previous rank
                self.units_total = n_per
            else:
                c_first, c_loc = sharding.shard_channels(chans, world, rank)
                self.units_total = chans * n_per * (4 if name == "interp" else 1)
            n_loc = n_per
            if name == "iir_scan" and world > 1:
                self.seg_first, n_loc = sharding.shard_stream(n_per, 32, world, rank)
            self.x = rand_c((c_loc, n_loc))
            if name == "decim":
                self.filt = DecimatingFIRFilter(taps, 1.0, 8, n_channels=c_loc)
            elif name == "decim_pc":  # [C][T] taps: every channel owns its coefficients (fir/mod.rs:79-88)
                self.taps = taps = taps[c_first:c_first + c_loc]
                self.filt = DecimatingFIRFilter(taps, 1.0, 8)
            elif name == "ddc":
                from solid_dsp_b200.filter.ddc import DigitalDownConverter
                self.filt = DigitalDownConverter(taps, 1.0, 8, DDC_FREQ, n_channels=c_loc)
            elif name == "interp":
                self.filt = InterpolatingFIRFilter(taps, 4, n_channels=c_loc)
            elif name == "autocorr":
                from solid_dsp_b200.filter.auto_correlator import AutoCorrelator
                self.filt = AutoCorrelator(*taps, n_channels=c_loc)
            else:
                self.filt = IIRFilter(taps[0], taps[1], IIRFilterType.SecondOrder, n_channels=c_loc)
            if name == "iir_scan" and world > 1:
                warm = self.filt.decay_length()
                assert warm > 0, "the benchmark cascade decays"
                self.halo_prev = torch.zeros(warm, dtype=torch.complex64, device=dev)

                def step():
                    sharding.exchange_halo(self.x[0], self.halo_prev, rank, world, dist)
                    return self._timed(lambda: sharding.iir_segment(self.filt, self.x, self.halo_prev.unsqueeze(0), rank))
            else:
                def step():
                    return self._timed(lambda: self.filt.execute_block(self.x))
        self.step = step

    def _timed(self, fn):
        """CUDA events around the execute_block call(s) of a step, on the stream they are launched on."""
        torch = self.torch
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        y = fn()
        k1.record()
        self.kernel_events.append((k0, k1))
        return y

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run_timed(self, warmup, steps, sample_clocks=True):
        """W warm-up steps, K timed steps between barriers; device time, max over ranks."""
        torch, dist = self.torch, self.dist
        from solid_dsp_b200 import launch_count
        # the clock sampler starts BEFORE the warm-up (nvidia-smi needs a few hundred ms to deliver its first line: waiting
        # for it between the warm-up and the timed steps would let the GPU fall idle right before the timed region); only
        # the lines that arrive after the timed region has started are evaluated
        sampler = None
        if self.rank == 0 and sample_clocks:
            sampler = ClockSampler(self.dev.index)
            sampler.start()
            sampler.wait_first()
        # the warm-up keeps its result alive across the next step exactly like the timed loop does: a step's output (32 GiB
        # for the interpolator) is allocated while the previous one is still referenced, so the caching allocator needs TWO
        # blocks -- a warm-up that dropped each result at once left the second cudaMalloc (~100 ms) inside timed step 2
        y = None
        for _ in range(warmup):
            y = self.step()
        self.barrier()
        if sampler is not None:
            sampler.mark()
        self.kernel_events = []
        l0 = launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            y = self.step()
        e1.record()
        self.barrier()
        ms_total = e0.elapsed_time(e1)
        # workloads without per-call events (fir64: the step is one C call) time the kernel as the step
        per_step = [a.elapsed_time(b) for a, b in self.kernel_events]
        ms_kernel = sum(per_step) / steps if per_step else ms_total / steps
        launches = launch_count() - l0
        if self.world > 1:
            t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total, ms_kernel = float(t[0].item()), float(t[1].item())
            lt = torch.tensor([launches], dtype=torch.int64, device=self.dev)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        # short runs (fir64: 200 calls of ~10 us; 8 ranks: 10 steps of ~1 ms) end before nvidia-smi has taken a sample under
        # load: keep the same step going for about 0.5 s (untimed, after the counters were read).  EVERY rank runs the same
        # number of extra steps -- a step of the stream workloads holds a point-to-point halo exchange.
        if sample_clocks and ms_total < 500.0:
            n_extra = min(20000, int(500.0 / max(ms_total / steps, 1e-3)) + 1)
            for i in range(n_extra):
                self.step()
                if i % 50 == 49:
                    torch.cuda.synchronize()
            self.barrier()
        self.kernel_events = []
        clocks = sampler.stop() if sampler is not None else None
        self.last_y = y
        ms_step = ms_total / steps
        return {"ms_step": ms_step, "ms_kernel": ms_kernel, "value": self.units_total / (ms_step * 1e-3) / 1e9,
                "launches": launches, "clocks": clocks, "kernel_ms_steps": [round(v, 4) for v in per_step[:64]]}

    def parity(self):
        """Oracle check of the timed buffers on EVERY rank (the first outputs of rank r > 0 depend on the halo it
        received); the worst error over the ranks is reported."""
        torch, dist = self.torch, self.dist
        nco_raw = None
        if self.name == "ddc":  # (theta, delta_theta) of every channel's NCO entering the checked step
            nco_raw = {int(c): self.filt.nco.raw(int(c)) for c in range(self.x.shape[0])} if self.x.shape[0] <= 64 else \
                _LazyRaw(self.filt.nco)
        y = self.step()  # known entry state (halo / fresh history) for the buffers that get checked
        torch.cuda.synchronize()
        if isinstance(nco_raw, _LazyRaw):
            nco_raw.rewind = self.x.shape[1]
        p = spot_check(self.name, self.taps, self.filt, self.x, y, self.halo_prev, self.rank, nco_raw)
        if self.world > 1:
            t = torch.tensor([p["max_normalised_error"]], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p["max_normalised_error_rank0"] = p["max_normalised_error"]
            p["max_normalised_error"] = float(t.item())
            p["ok"] = bool(p["max_normalised_error"] <= 1e-5)
            p["ranks_checked"] = self.world
        return p

    def roofline(self, ms_kernel):
        name, W = self.name, WORKLOADS[self.name]
        peak_fma, peak_detail = measure_peaks()
        fp = file_peaks()
        units_rank = self.units_total / self.world
        ach_tflops = W["flop_per_unit"] * units_rank / (ms_kernel * 1e-3) / 1e12
        ach_gbs = W["bytes_per_unit"] * units_rank / (ms_kernel * 1e-3) / 1e9
        t_fma = W["flop_per_unit"] / (peak_fma * 1e12)
        t_hbm = W["bytes_per_unit"] / (fp["hbm"] * 1e9)
        bound = "fma" if t_fma >= t_hbm else "hbm"
        # Long FIR: the library runs it on the tcgen05 tensor cores (csrc/fir_tc.cu, a banded-Toeplitz GEMM with a
        # 16-bit operand split), so the binding roofline is the tensor pipe (measured bf16 dense rate).
        tensor = None
        on_tensor = name == "fir" and getattr(self.filt, "last_path", "ffma") == "tensor"
        if on_tensor:
            T_ = len(self.taps)
            koff = (T_ - 1 + 31) // 32 * 32
            band = (koff + 128) / T_  # band of Koff + 128 columns per 128 outputs, zeros included
            fmt = os.environ.get("SGPU_FIR_TC_FMT", "f16")
            products = 6.0 if fmt.startswith("b") else 3.0
            # one f32-accurate real product = `products` 16-bit products (F16x2 block floating point: f1h1, f1h2, f2h1;
            # BF16x3: six), so the pipe's f32-equivalent peak is the measured 16-bit dense rate / products; `executed`
            # counts every 16-bit flop the pipe really does
            executed = products * band
            f32_eq_peak = fp["bf16"] / products
            tensor = {"achieved_tflops": ach_tflops, "peak_tflops": f32_eq_peak, "frac": ach_tflops / f32_eq_peak,
                      "peak_source": f"f32-equivalent tensor peak = 16-bit dense / {products:.0f} "
                                     f"({'BF16x3: 6' if products == 6 else 'F16x2 block floating point: 3'} MMAs per f32 product); 16-bit dense: " + fp["bf16_src"],
                      "bf16_peak_tflops": fp["bf16"], "algorithmic_frac_of_bf16_peak": ach_tflops / fp["bf16"],
                      "executed_16bit_tflops": ach_tflops * executed, "executed_frac_of_bf16_peak": ach_tflops * executed / fp["bf16"],
                      "executed_over_algorithmic": executed, "band_overhead": band, "products_per_f32_product": products,
                      "note": "algorithmic flops = 4 per real x complex tap; the tensor pipe executes `products` 16-bit MMAs per K step "
                              "over a band of (Koff+128)/T columns (Toeplitz zeros), i.e. products x band times the algorithmic flops"}
            bound = "tensor"
        # DRAM bytes (read + write) of one launch of the dominant kernel at this workload's full size,
        # from a committed `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` capture
        # (tools/profile_round.sh -> tools/traffic_json.py); only valid for the default sizes at N = 1.
        traffic, traffic_detail = None, None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists() and self.world == 1 and self.args.log2_samples == 30:
            traffic_detail = json.loads(tf.read_text()).get(name)
            if traffic_detail:
                traffic = traffic_detail.get("dram_bytes")
        kernel = {"fir": "fir_tc_fused_kernel (tcgen05.mma kind::f16, TMA, TMEM) + fir_tc_post_kernel" if on_tensor else "fir_warp_kernel<R=16>",
                  "fir64": "fir_warp_kernel<R=16> (one launch per call, history written by the same kernel)",
                  "decim": "fir_decim_warp_kernel<M=8,PS=4>", "decim_pc": "fir_decim_warp_kernel<M=8,PS=4> (per-channel tap images)", "ddc": "fir_decim_warp_kernel<M=8,PS=4,NCO mix>",
                  "interp": "fir_interp_walk_kernel<L=4,K=5>", "iir_batch": "iir_sos_kernel<8>",
                  "iir_scan": "iir_sos_kernel<8> (chunked scan)", "autocorr": "autocorr_kernel"}[name]
        r = {
            "bound": bound,
            "achieved": ach_gbs if bound == "hbm" else ach_tflops,
            "peak": {"fma": peak_fma, "hbm": fp["hbm"], "tensor": tensor and tensor["peak_tflops"]}[bound],
            "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
            "frac": {"fma": ach_tflops / peak_fma, "hbm": ach_gbs / fp["hbm"], "tensor": tensor and tensor["frac"]}[bound],
            "traffic": traffic,
            "traffic_detail": traffic_detail,
            "kernel": kernel,
            "kernel_ms_per_launch": ms_kernel,
            "timed": "CUDA events around execute_block inside the timed loop that gives `value`",
            "algorithmic": {"flop_per_unit": W["flop_per_unit"], "bytes_per_unit": W["bytes_per_unit"], "unit": W["unit"],
                            "units_per_launch": units_rank},
            "fma": {"achieved_tflops": ach_tflops, "peak_tflops": peak_fma, "frac": ach_tflops / peak_fma,
                    "peak_source": "measured live: libsgpu_peakbench FFMA chains (fp32, non-tensor)", "detail": peak_detail},
            "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": fp["hbm"], "frac": ach_gbs / fp["hbm"], "peak_source": fp["hbm_src"]},
        }
        if tensor:
            r["tensor"] = tensor
        return r, on_tensor


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    wl = Workload(name, args, world, rank, dev, dist)
    res = wl.run_timed(args.warmup, args.steps)
    parity = None if args.no_check else wl.parity()
    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(name, wl.taps, wl.x, world, rank, dev, wl.units_total, args)
    line = None
    if rank == 0:
        roofline, on_tensor = wl.roofline(res["ms_kernel"])
        roofline["kernel_ms_steps"] = res["kernel_ms_steps"]
        cpu = None
        if world == 1 and not args.no_cpu:
            rate1, what1, _ = cpu_reference_path(name, 1, args.cpu_seconds)
            cores = os.cpu_count() or 1
            rateN, whatN, _ = cpu_reference_path(name, cores, args.cpu_seconds)
            cpu = {"value": rate1 / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "sample": what1,
                   "all_cores": {"value": rateN / 1e9, "cores": cores, "sample": whatN}}
        fmt = os.environ.get("SGPU_FIR_TC_FMT", "f16")
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_step"], "higher_is_better": True,
            "scaling": "weak" if name == "fir64" else "strong", "vs_baseline": None,
            "dtype": ("f32 (16-bit operand split on the tensor cores, f32 accumulation: "
                      + ("BF16x3, 6 MMAs per product)" if fmt.startswith("b") else "F16x2 block floating point, 3 MMAs per product)"))
            if on_tensor else "f32", "data": "synthetic",
            "config": workload_config(name, world, args.log2_samples),
            "e2e": e2e, "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity,
        }
    # the workload object sits in a reference cycle (its step closure): collect it NOW so that its handle -- and with it
    # the persisting-L2 carve-out of the tensor FIR -- is gone before the next workload is timed
    del wl
    gc.collect()
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs: driver-visible numbers for every config.  One GPU: all of them, with a 1-core CPU
    # baseline each.  N > 1: the channel-sharded ones (contiguous channel ranges, no data-path collective),
    # oracle parity on EVERY rank -- so that the driver's scaling run carries them too.
    if name == "fir" and args.log2_samples == 30 and not args.no_workloads:
        default_side = "fir64,decim,decim_pc,interp,iir_batch,iir_scan,ddc" if world == 1 else "decim,interp,iir_batch"
        side = [w for w in os.environ.get("SGPU_BENCH_SIDE", default_side).split(",") if w in WORKLOADS and w != "fir"]
        block = {}
        guard = None
        if world > 1:
            # a rank that fails alone would leave the others in a collective: after `side_timeout` seconds every rank gives
            # up on the block, rank 0 still prints the headline line (the block says what happened), exit code 0
            def give_up(why=None):
                if rank == 0:
                    line["workloads"] = dict(block, error=why or f"side workloads did not finish within {args.side_timeout:.0f} s")
                    print(json.dumps(line), flush=True)
                os._exit(0)
            guard = threading.Timer(args.side_timeout, give_up)
            guard.daemon = True
            guard.start()
        for other in side:
            try:
                res_o = run_side_workload(other, args, dev, dist, world, rank)
            except Exception as e:  # noqa: BLE001  (a failing side workload must not take the headline line with it)
                res_o = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:  # the other ranks are inside this workload's collectives: their timers end them, exit code 0
                    give_up(f"rank {rank}, {other}: {res_o['error']}")
            block[other] = res_o
            gc.collect()
            torch.cuda.empty_cache()
        if guard is not None:
            guard.cancel()
        if rank == 0:
            line["workloads"] = block
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_side_workload(name, args, dev, dist, world=1, rank=0):
    """One of the other BASELINE configs at its full size, about 5 s: device-resident value, kernel time, both roofline
    denominators, oracle parity (every rank), clocks, 1-core cpu_baseline (one GPU only)."""
    wl = Workload(name, args, world, rank, dev, dist)
    steps = 200 if name == "fir64" else 10
    res = wl.run_timed(3, steps)
    parity = None if args.no_check else wl.parity()
    if rank != 0:
        return None
    roofline, _ = wl.roofline(res["ms_kernel"])
    cpu = None
    if not args.no_cpu and world == 1:
        rate1, what1, _ = cpu_reference_path(name, 1, min(args.cpu_seconds, 3.0))
        cpu = {"value": rate1 / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "sample": what1}
    out = {"config": workload_config(name, world, args.log2_samples), "value": res["value"], "unit": UNIT, "n_gpus": world,
           "scaling": "weak" if name == "fir64" else "strong", "steps": steps,
           "warmup": 3, "ms_per_step": res["ms_step"], "kernel_ms": res["ms_kernel"], "kernel_ms_steps": res["kernel_ms_steps"],
           "gpu_launches": res["launches"],
           "clocks": res["clocks"],
           "roofline": {k: roofline[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel")},
           "roofline_frac_fma": roofline["fma"]["frac"], "roofline_frac_hbm": roofline["hbm"]["frac"],
           "parity": parity, "cpu_baseline": cpu}
    if name in ("decim", "decim_pc", "interp"):
        out["unit_note"] = "decim: input samples/s; interp: output samples/s"
    return out


def spot_check(name, taps, filt, x, y, halo_prev, rank=0, nco_raw=None):
    """Oracle check of a few windows of the buffers that were just timed."""
    import oracle as O
    rng = np.random.default_rng(7 + rank)
    worst = 0.0
    windows = []

    def nerr(got, ref):
        return float(np.max(np.abs(got.astype(np.complex128) - ref)) / max(np.max(np.abs(ref)), 1e-30))

    if name == "fir":
        T = len(taps)
        n = x.shape[0]
        # window 0 = the first 4096 outputs: on rank r > 0 they depend on the halo received from rank r - 1
        for start in [0] + [int(s) for s in rng.integers(T, max(T + 1, n - 4096), 3)] + [n - 4096]:
            start = max(0, min(start, n - 4096))
            lo = max(0, start - (T - 1))
            xs = x[lo:start + 4096].cpu().numpy()
            if start < T - 1:
                pre = halo_prev.cpu().numpy()[-(T - 1 - start):] if start < T - 1 else np.zeros(0)
                xs = np.concatenate([pre, xs])
            ref = O.fir_fast(taps, xs)[-4096:]
            e = nerr(y[start:start + 4096].cpu().numpy(), ref)
            windows.append([start, e])
            worst = max(worst, e)
    elif name == "fir64":
        # the handle streams over the rotating rows: check the row of the last call on a fresh handle
        from solid_dsp_b200.filter.fir import FIRFilter
        row = int(rng.integers(0, x.shape[0]))
        got = FIRFilter(taps, 1.0).execute_block(x[row]).cpu().numpy()
        xs = x[row].cpu().numpy()
        for start in (0, 500000, (1 << 20) - 8192):
            lo = max(0, start - 63)
            ref = O.fir_fast(taps, xs[lo:start + 8192])[start - lo:]
            e = nerr(got[start:start + 8192], ref)
            windows.append([row, start, e])
            worst = max(worst, e)
    elif name in ("decim", "ddc", "decim_pc"):
        # the timed handle streams (history = tail of the previous step): compare outputs that
        # depend only on this call's inputs, i.e. skip the first ceil(T/M) of every window
        skip = np.shape(taps)[-1] // 8 + 1
        n = x.shape[1]
        for c in rng.integers(0, x.shape[0], 3):
            taps_c = taps[int(c)] if name == "decim_pc" else taps
            for start in (0, (n // 2) & ~7, n - (1 << 15)):
                xs = x[int(c), start:start + (1 << 15)].cpu().numpy()
                if name == "ddc":  # the NCO phase at sample `start` of the checked step
                    th0, dl = nco_raw[int(c)]
                    xs = O.nco_mix_down_block(xs, raw=((th0 + start * dl) & 0xFFFFFFFF, dl))
                ref = O.fir_fast(taps_c, xs, 1.0, 8)[skip:]
                got = y[int(c), start // 8 + skip:start // 8 + skip + len(ref)].cpu().numpy()
                e = nerr(got, ref)
                windows.append([int(c), start, e])
                worst = max(worst, e)
    elif name == "interp":
        skip = (len(taps) // 4 + 1) * 4
        n = x.shape[1]
        for c in rng.integers(0, x.shape[0], 3):
            for start in (0, n // 2, n - (1 << 13)):
                xs = x[int(c), start:start + (1 << 13)].cpu().numpy()
                ref = O.firinterp_fast(taps, 4, xs)[skip:]
                got = y[int(c), start * 4 + skip:start * 4 + skip + len(ref)].cpu().numpy()
                e = nerr(got, ref)
                windows.append([int(c), start, e])
                worst = max(worst, e)
    elif name == "autocorr":
        # outputs further than window_size into the call depend on this call's inputs only
        W, d = taps
        n = x.shape[1]
        for c in rng.integers(0, x.shape[0], 3):
            for start in (0, n // 2, n - (1 << 13)):
                xs = x[int(c), start:start + (1 << 13)].cpu().numpy()
                ref = O.autocorr_fast(W, d, xs)[W:]
                got = y[int(c), start + W:start + W + len(ref)].cpu().numpy()
                e = nerr(got, ref)
                windows.append([int(c), start, e])
                worst = max(worst, e)
    else:
        # the timed handle is streaming (state persists over steps): check a fresh handle instead
        from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType
        ff, fb = taps
        if x.shape[0] == 1:
            n = min(x.shape[1], 1 << 20)
            f = IIRFilter(ff, fb, IIRFilterType.SecondOrder)
            f.set_mode(1)
            got = f.execute_block(x[:, :n]).cpu().numpy()[0]
            ref, _ = O.sos_cascade_fast(ff, fb, x[0, :n].cpu().numpy())
            worst = nerr(got, ref)
        else:
            f = IIRFilter(ff, fb, IIRFilterType.SecondOrder, n_channels=x.shape[0])
            got = f.execute_block(x)
            for c in rng.integers(0, x.shape[0], 3):
                ref, _ = O.sos_cascade_fast(ff, fb, x[int(c)].cpu().numpy())
                worst = max(worst, nerr(got[int(c)].cpu().numpy(), ref))
            del got
    return {"max_normalised_error": worst, "tolerance": 1e-5, "ok": bool(worst <= 1e-5), "checker": "oracle (f64)",
            "windows": windows}


def measure_e2e(name, taps, x, world, rank, dev, units_total, args):
    """Same workload through the C ABI with SGPU_HOST buffers in pinned memory."""
    import torch
    import torch.distributed as dist
    from solid_dsp_b200 import _ffi
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter, FIRFilter, InterpolatingFIRFilter
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType

    if name == "fir64":
        x = x[0]
    two_d = x.dim() == 2
    c_loc = x.shape[0] if two_d else 1
    n_in = x.shape[-1]
    if name in ("fir", "fir64"):
        f = FIRFilter(taps, 1.0)
        n_out = n_in
    elif name == "decim":
        f = DecimatingFIRFilter(taps, 1.0, 8, n_channels=c_loc)
        n_out = n_in // 8
    elif name == "ddc":
        from solid_dsp_b200.filter.ddc import DigitalDownConverter
        f = DigitalDownConverter(taps, 1.0, 8, DDC_FREQ, n_channels=c_loc)
        n_out = n_in // 8
    elif name == "interp":
        f = InterpolatingFIRFilter(taps, 4, n_channels=c_loc)
        n_out = n_in * 4
    elif name == "autocorr":
        from solid_dsp_b200.filter.auto_correlator import AutoCorrelator
        f = AutoCorrelator(*taps, n_channels=c_loc)
        n_out = n_in
    else:
        f = IIRFilter(taps[0], taps[1], IIRFilterType.SecondOrder, n_channels=c_loc)
        n_out = n_in
    # pinned staging buffers on the NUMA node of this rank's GPU (sgpu_host_alloc): with every rank's buffers on node 0
    # the 8-GPU run moved 8.3 GB/s per GPU per direction instead of the 47.8 GB/s one GPU gets (round 1)
    from solid_dsp_b200.hostmem import PinnedArray
    pin_in, pin_out = PinnedArray(c_loc, n_in, dev.index), PinnedArray(c_loc, n_out, dev.index)
    hin, hout = torch.from_numpy(pin_in.array), torch.from_numpy(pin_out.array)
    hin.copy_(x.reshape(c_loc, n_in))
    fn = {"fir": _ffi.lib.sgpu_fir_execute_block, "fir64": _ffi.lib.sgpu_fir_execute_block,
          "decim": _ffi.lib.sgpu_fir_execute_block, "interp": _ffi.lib.sgpu_interp_execute_block,
          "autocorr": _ffi.lib.sgpu_autocorr_execute_block}.get(name, _ffi.lib.sgpu_iir_execute_block)
    if name == "ddc":
        fn = _ffi.lib.sgpu_ddc_execute_block
    got = _ffi.c_size()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step():
        _ffi.check(fn(f._h, hin.data_ptr(), n_in, n_in, hout.data_ptr(), n_out, C.byref(got), _ffi.HOST, stream))

    step()  # warm-up (allocates the handle's staging buffers)
    steps = max(1, min(args.steps, 3)) if name != "fir64" else 50
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()  # returns after the result is in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    h2d = c_loc * n_in * 8 * world
    d2h = c_loc * n_out * 8 * world
    checksum = complex(hout[0, :16].sum().item())
    del hin, hout
    pin_in.free()
    pin_out.free()
    return {"value": units_total / (dt / steps) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "api": "sgpu_*_execute_block(mem=SGPU_HOST), pinned host buffers from sgpu_host_alloc (NUMA node of the GPU)", "result_checksum": [checksum.real, checksum.imag]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fir", choices=list(WORKLOADS))
    ap.add_argument("--log2-samples", type=int, default=30, help="fir: total stream length; others: per-channel override")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline budget per measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the block of the other BASELINE configs")
    ap.add_argument("--side-timeout", type=float, default=240.0,
                    help="N > 1: seconds the block of side workloads may take before the headline line is printed without it")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="seconds after which a stuck run dumps its stacks and exits 3")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)  # a hung collective must not eat the box's time limit
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
