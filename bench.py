#!/usr/bin/env python
"""bench.py -- throughput of the filtering hot path on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fir|decim|interp|iir_batch|iir_scan|autocorr]
    python bench.py --impl reference ...        # the reference's CPU path (restated oracle) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...   (one rank per GPU)

Default workload = BASELINE.json configs[1]: one complex stream of 2^30 samples through a 512-tap
real-coefficient Kaiser low-pass (firdes_kaiser(512, 0.1, 80 dB)), contiguous segments sharded over
the ranks with a 511-sample halo (strong scaling: total work fixed).  A step is one pass over the
whole stream: halo write + FIR execute_block on every rank.

`value`   : complex Gsamples/s with inputs resident in HBM (device pointers through the C ABI).
`e2e`     : same metric through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region.
`roofline`: the dominant kernel against the FP32-FMA peak measured live (libsgpu_peakbench) and
            against the measured HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline`: the restated reference CPU path (oracle/, structural: Window memmove + to_vec +
            sequential DotProduct per sample) on the box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
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
    "decim": dict(desc="BASELINE configs[2]: decimator M=8, 256 taps, 4096 channels x 2^20 samples",
                  flop_per_unit=128.0, bytes_per_unit=9.0, unit="input sample"),
    "interp": dict(desc="BASELINE configs[3]: interpolator L=4, 128 taps, 1024 channels x 2^20 inputs",
                   flop_per_unit=128.0, bytes_per_unit=10.0, unit="output sample"),
    "iir_batch": dict(desc="BASELINE configs[4a]: 8-section biquad cascade, 65536 channels x 2^14 samples",
                      flop_per_unit=144.0, bytes_per_unit=16.0, unit="sample"),
    "iir_scan": dict(desc="BASELINE configs[4b]: 8-section biquad cascade, one stream of 2^28 samples (chunked scan)",
                     flop_per_unit=144.0, bytes_per_unit=16.0, unit="sample"),
    # widening, not a BASELINE config: 8 flop per lag product + 2 complex adds per output
    "autocorr": dict(desc="SURVEY 8f: AutoCorrelator window 64, delay 16, 1024 channels x 2^20 samples",
                     flop_per_unit=16.0, bytes_per_unit=16.0, unit="sample"),
}
AUTOCORR_SHAPE = (64, 16)


def f32_taps(h):
    return np.asarray(h, dtype=np.float32).astype(np.float64)


def workload_taps(name):
    from solid_dsp_b200.filter import firdes, iirdes
    if name == "fir":
        return f32_taps(firdes.firdes_kaiser(512, 0.1, 80.0, 0.0))
    if name == "decim":
        return f32_taps(firdes.firdes_kaiser(256, 0.5 / 8 * 0.9, 80.0, 0.0))
    if name == "interp":
        return f32_taps(firdes.firdes_kaiser(128, 0.5 / 4 * 0.9, 80.0, 0.0))
    if name == "autocorr":
        return AUTOCORR_SHAPE
    return iirdes.stable_lowpass_sections(8)


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

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
        for ln in self.lines:
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
    taps = workload_taps(name)

    def run(n_units, n_in):
        if name == "fir":
            T = len(taps)
            # stream segments, each primed with T-1 preroll samples through the public API
            x = (rng.uniform(-1, 1, n_units * n_in + T) + 1j * rng.uniform(-1, 1, n_units * n_in + T))
            out = np.zeros(n_units * n_in, dtype=np.complex128)
            t0 = time.perf_counter()
            O.run_units("fir", x, n_in, n_units, n_in, out, n_in, n_threads, coefs=taps, preroll=T - 1,
                        x_offset=T - 1, native=native)
            return time.perf_counter() - t0, n_units * n_in
        if name == "decim":
            x = rng.uniform(-1, 1, n_units * n_in) + 1j * rng.uniform(-1, 1, n_units * n_in)
            out = np.zeros(n_units * (n_in // 8 + 1), dtype=np.complex128)
            t0 = time.perf_counter()
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
    reference is Rust and cannot be built here), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
        "config": {"workload": WORKLOADS[name]["desc"], "name": name},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- GPU arm
def measure_peaks():
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
    return best, detail


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
    from solid_dsp_b200 import _ffi, launch_count, sharding
    from solid_dsp_b200.filter.fir import DecimatingFIRFilter, FIRFilter, InterpolatingFIRFilter
    from solid_dsp_b200.filter.iir import IIRFilter, IIRFilterType

    name = args.workload
    W = WORKLOADS[name]
    taps = workload_taps(name)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)

    def rand_c(shape):
        t = torch.empty(tuple(shape) + (2,), dtype=torch.float32, device=dev)
        t.uniform_(-1.0, 1.0, generator=gen)
        return torch.view_as_complex(t)

    halo_prev = None
    if name == "fir":
        n_total = 1 << args.log2_samples
        _, n_loc = sharding.shard_stream(n_total, 1, world, rank)
        x = rand_c((n_loc,))
        filt = FIRFilter(taps, 1.0)
        T = len(taps)
        halo_prev = torch.zeros(T - 1, dtype=torch.complex64, device=dev)
        units_total = n_total
        shape_desc = {"samples": n_total, "taps": T, "segments": world, "halo": T - 1}

        def step():
            # halo: last T-1 samples of the previous rank's segment (rank 0: zeros = fresh filter)
            sharding.exchange_halo(x, halo_prev, rank, world, dist)
            filt.write(halo_prev)
            return filt.execute_block(x)
    else:
        chans = {"decim": 4096, "interp": 1024, "iir_batch": 65536, "iir_scan": 1, "autocorr": 1024}[name]
        n_per = 1 << {"decim": 20, "interp": 20, "iir_batch": 14, "iir_scan": 28, "autocorr": 20}[name]
        if args.log2_samples != 30:
            n_per = 1 << args.log2_samples
        if name == "iir_scan":
            c_loc = 1  # one stream: time segments per rank, warm-up halo from the previous rank
            units_total = n_per
        else:
            _, c_loc = sharding.shard_channels(chans, world, rank)
            units_total = chans * n_per * (4 if name == "interp" else 1)
        n_loc = n_per
        if name == "iir_scan" and world > 1:
            _, n_loc = sharding.shard_stream(n_per, 32, world, rank)
        x = rand_c((c_loc, n_loc))
        if name == "decim":
            filt = DecimatingFIRFilter(taps, 1.0, 8, n_channels=c_loc)
        elif name == "interp":
            filt = InterpolatingFIRFilter(taps, 4, n_channels=c_loc)
        elif name == "autocorr":
            from solid_dsp_b200.filter.auto_correlator import AutoCorrelator
            filt = AutoCorrelator(*taps, n_channels=c_loc)
        else:
            filt = IIRFilter(taps[0], taps[1], IIRFilterType.SecondOrder, n_channels=c_loc)
        shape_desc = {"channels": chans, "samples_per_channel": n_per}
        if name == "iir_scan" and world > 1:
            warm = filt.decay_length()
            assert warm > 0, "the benchmark cascade decays"
            halo_prev = torch.zeros(warm, dtype=torch.complex64, device=dev)
            shape_desc.update({"segments": world, "halo": warm})

            def step():
                sharding.exchange_halo(x[0], halo_prev, rank, world, dist)
                return sharding.iir_segment(filt, x, halo_prev.unsqueeze(0), rank)
        else:
            def step():
                return filt.execute_block(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up, K timed steps between barriers, max over ranks
    y = None
    for _ in range(args.warmup):
        y = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        y = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = launch_count() - l0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = units_total / (ms_step * 1e-3) / 1e9

    # ---- dominant kernel alone (execute_block only) for the roofline, CUDA events on the same stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        y = filt.execute_block(x)
    k1.record()
    torch.cuda.synchronize()
    ms_kernel = k0.elapsed_time(k1) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity spot check on the timed buffers (oracle = checker; rank 0, small windows)
    parity = None
    if not args.no_check:
        y = step()  # known entry state (halo / fresh history) for the buffers that get checked
        torch.cuda.synchronize()
        if rank == 0:
            parity = spot_check(name, taps, filt, x, y, halo_prev)

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(name, taps, x, world, rank, dev, units_total, args)

    if rank == 0:
        peak_fma, peak_detail = measure_peaks()
        units_rank = units_total / world
        ach_tflops = W["flop_per_unit"] * units_rank / (ms_kernel * 1e-3) / 1e12
        ach_gbs = W["bytes_per_unit"] * units_rank / (ms_kernel * 1e-3) / 1e9
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if peaks_file.exists():
            hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
            hbm_src = "measured (MEASURED_PEAKS.json)"
        t_fma = W["flop_per_unit"] / (peak_fma * 1e12)
        t_hbm = W["bytes_per_unit"] / (hbm_peak * 1e9)
        bound = "fma" if t_fma >= t_hbm else "hbm"
        # Long real-tap FIR: the library runs it on the tcgen05 tensor cores (csrc/fir_tc.cu, BF16x3 over a
        # banded-Toeplitz GEMM), so the binding roofline is the tensor pipe (measured bf16 dense rate).
        tensor = None
        if name == "fir" and getattr(filt, "last_path", "ffma") == "tensor":
            bf16_peak, bf16_src = 1590.0, "fallback (B200_PROFILING.md)"
            if peaks_file.exists():
                pk = json.loads(peaks_file.read_text())
                bf16_peak = float(pk.get("bf16_tflops_sustained") or pk["bf16_tflops"])
                bf16_src = ("measured (MEASURED_PEAKS.json, bf16_tflops_sustained: the kernel runs for several ms "
                            "back to back under the power cap)")
            T_ = len(taps)
            koff = (T_ - 1 + 31) // 32 * 32
            band = (koff + 128) / T_            # band of Koff + 128 columns per 128 outputs, zeros included
            # one f32-accurate real product = 6 bf16 products (b1b1, b1b2, b2b1, b2b2, b1b3, b3b1), so the pipe's
            # f32-equivalent peak is the measured bf16 rate / 6; `executed` counts every bf16 flop the pipe really does
            executed = 6.0 * band
            f32_eq_peak = bf16_peak / 6.0
            tensor = {"achieved_tflops": ach_tflops, "peak_tflops": f32_eq_peak, "frac": ach_tflops / f32_eq_peak,
                      "peak_source": "f32-equivalent tensor peak = bf16 dense / 6 (BF16x3 split: 6 bf16 MMAs per f32 product); bf16: " + bf16_src,
                      "bf16_peak_tflops": bf16_peak,
                      "executed_bf16_tflops": ach_tflops * executed, "executed_frac_of_bf16_peak": ach_tflops * executed / bf16_peak,
                      "executed_over_algorithmic": executed, "band_overhead": band,
                      "note": "algorithmic flops = 4 per real x complex tap; the tensor pipe executes 6 bf16 MMAs per K step "
                              "over a band of (Koff+128)/T columns (Toeplitz zeros), i.e. 6 x band times the algorithmic flops"}
            bound = "tensor"
        # DRAM bytes (read + write) of one launch of the dominant kernel at this workload's full size,
        # from a committed `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` capture
        # (tools/profile_round.sh -> tools/traffic_json.py); only valid for the default sizes at N = 1.
        traffic, traffic_detail = None, None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists() and world == 1 and args.log2_samples == 30:
            traffic_detail = json.loads(tf.read_text()).get(name)
            if traffic_detail:
                traffic = traffic_detail.get("dram_bytes")
        roofline = {
            "bound": bound,
            "achieved": ach_gbs if bound == "hbm" else ach_tflops,
            "peak": {"fma": peak_fma, "hbm": hbm_peak, "tensor": tensor and tensor["peak_tflops"]}[bound],
            "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
            "frac": {"fma": ach_tflops / peak_fma, "hbm": ach_gbs / hbm_peak,
                     "tensor": tensor and tensor["frac"]}[bound],
            "traffic": traffic,
            "traffic_detail": traffic_detail,
            "kernel": {"fir": "fir_tc_fused_kernel<BF16x3> (tcgen05.mma kind::f16, TMA, TMEM)" if tensor else "fir_warp_kernel<R=16>", "decim": "fir_decim_kernel<R=16>",
                       "interp": "fir_interp_kernel<R=16>", "iir_batch": "iir_sos_kernel<8>",
                       "iir_scan": "iir_sos_kernel<8> (fused warm-up scan, one launch)",
                       "autocorr": "autocorr_kernel"}[name],
            "kernel_ms_per_launch": ms_kernel,
            "algorithmic": {"flop_per_unit": W["flop_per_unit"], "bytes_per_unit": W["bytes_per_unit"], "unit": W["unit"],
                            "units_per_launch": units_rank},
            "fma": {"achieved_tflops": ach_tflops, "peak_tflops": peak_fma, "frac": ach_tflops / peak_fma,
                    "peak_source": "measured live: libsgpu_peakbench FFMA chains (fp32, non-tensor)", "detail": peak_detail},
            "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": hbm_peak, "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
        }
        if tensor:
            roofline["tensor"] = tensor
        cpu = None
        if world == 1 and not args.no_cpu:
            rate1, what1, _ = cpu_reference_path(name, 1, args.cpu_seconds)
            cores = os.cpu_count() or 1
            rateN, whatN, _ = cpu_reference_path(name, cores, args.cpu_seconds)
            cpu = {"value": rate1 / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "sample": what1,
                   "all_cores": {"value": rateN / 1e9, "cores": cores, "sample": whatN}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32 (BF16x3 split on the tensor cores: 6 bf16 MMAs per product, f32 accumulation)" if tensor else "f32", "data": "synthetic",
            "config": {"workload": W["desc"], "name": name, **shape_desc,
                       "l2": "inputs (>= 2 GiB per rank) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"stream segments x{world} with halo" if name in ("fir", "iir_scan") else f"channels x{world}"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def spot_check(name, taps, filt, x, y, halo_prev):
    """Oracle check of a few windows of the buffers that were just timed."""
    import oracle as O
    import torch
    rng = np.random.default_rng(7)
    worst = 0.0
    windows = []

    def nerr(got, ref):
        return float(np.max(np.abs(got.astype(np.complex128) - ref)) / max(np.max(np.abs(ref)), 1e-30))

    if name == "fir":
        T = len(taps)
        n = x.shape[0]
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
    elif name == "decim":
        # the timed handle streams (history = tail of the previous step): compare outputs that
        # depend only on this call's inputs, i.e. skip the first ceil(T/M) of every window
        skip = len(taps) // 8 + 1
        n = x.shape[1]
        for c in rng.integers(0, x.shape[0], 3):
            for start in (0, (n // 2) & ~7, n - (1 << 15)):
                xs = x[int(c), start:start + (1 << 15)].cpu().numpy()
                ref = O.fir_fast(taps, xs, 1.0, 8)[skip:]
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

    two_d = x.dim() == 2
    c_loc = x.shape[0] if two_d else 1
    n_in = x.shape[-1]
    if name == "fir":
        f = FIRFilter(taps, 1.0)
        n_out = n_in
    elif name == "decim":
        f = DecimatingFIRFilter(taps, 1.0, 8, n_channels=c_loc)
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
    hin = torch.empty((c_loc, n_in), dtype=torch.complex64, pin_memory=True)
    hout = torch.empty((c_loc, n_out), dtype=torch.complex64, pin_memory=True)
    hin.copy_(x.reshape(c_loc, n_in))
    fn = {"fir": _ffi.lib.sgpu_fir_execute_block, "decim": _ffi.lib.sgpu_fir_execute_block,
          "interp": _ffi.lib.sgpu_interp_execute_block,
          "autocorr": _ffi.lib.sgpu_autocorr_execute_block}.get(name, _ffi.lib.sgpu_iir_execute_block)
    got = _ffi.c_size()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step():
        _ffi.check(fn(f._h, hin.data_ptr(), n_in, n_in, hout.data_ptr(), n_out, C.byref(got), _ffi.HOST, stream))

    step()  # warm-up (allocates the handle's staging buffers)
    steps = max(1, min(args.steps, 3))
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
    h2d = c_loc * n_in * 8 * world if name != "fir" else n_in * 8 * world
    d2h = c_loc * n_out * 8 * world if name != "fir" else n_out * 8 * world
    checksum = complex(hout[0, :16].sum().item())
    return {"value": units_total / (dt / steps) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "api": "sgpu_*_execute_block(mem=SGPU_HOST), pinned host buffers", "result_checksum": [checksum.real, checksum.imag]}


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
