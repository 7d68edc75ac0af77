"""bench.py -- channel realizations/s of the README "advanced channel" (2048^2, 5 SS screens, 50 km) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A step is one batch of B independent realizations of the whole hot path (device RNG -> 5 x [leg + screen
synthesis + screen multiply] -> closing leg -> fused measures).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(n=2048, delta=1.5e-3, wvl=808e-9, w0=0.12, Cn2=5e-16, l0=6e-3, L0=1e3, m=2**10, f_min=1 / 1e3 / 15,
          f_max=1 / 6e-3 * 2, length=50e3, count=5, pupil=0.2)
WORKLOAD = "README advanced channel: 2048^2 grid, delta 1.5 mm, 5 SS screens (MVK, 2^10 rings), 50 km, complex64"
METRIC = "channel realizations/sec (2048^2, 5 screens)"
UNIT = "realizations/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the numpy restatement of the reference (oracle, mode="ref"), timed on host cores
# ---------------------------------------------------------------------------------------------------------------
def _cpu_realization(seed):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import splitstep as orc
    p = C3
    x, y = orc.rect_xy(p["n"], p["delta"])
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    psd = _cpu_realization.psd
    np.random.seed(seed)
    u = orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="ref")
    pos = orc.screen_positions(p["length"], p["count"])
    legs = orc.leg_lengths(p["length"], pos)
    for s in range(p["count"]):
        rho, theta, value = orc.draw_spectrum(base, psd)
        fx, fy = orc.spectrum_to_fxy(rho, theta)
        u = np.exp(-1j * orc.ss_screen(x, y, fx, fy, value, mode="ref")) * orc.vacuum_leg(u, legs[s], p["wvl"], p["delta"], "ref")
    u = orc.vacuum_leg(u, legs[-1], p["wvl"], p["delta"], "ref")
    m = orc.moments(u, x, y, p["delta"], pupils=[(p["pupil"], (0, 0))], mode="ref")
    return m["eta_pupil"][0]


def _cpu_init(psd):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    _cpu_realization.psd = psd


def cpu_psd():
    from oracle import splitstep as orc
    p = C3
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    return orc.ring_psd(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"])


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (numpy port in oracle/, pinned against the reference by
    tests/test_oracle_golden.py) with one realization per worker process per step, on all host cores (capped)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = cpu_cores()
    workers = max(1, min(cores, int(os.environ.get("PYATM_REF_WORKERS", "32"))))
    psd = cpu_psd()
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(psd,)) as pool:
        seed = 0
        for _ in range(args.warmup):
            pool.map(_cpu_realization, range(seed, seed + workers))
            seed += workers
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_realization, range(seed, seed + workers))
            seed += workers
        dt = time.perf_counter() - t0
    value = workers * args.steps / dt
    sample = f"{workers} realizations per step (one per worker process), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (numpy: c128 transfer-function product, as the reference)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "realizations_per_step": workers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region.  In-process NVML (nvidia_ml_py)
    every 20 ms (clock + event reasons).  An external `nvidia-smi -lms` process is avoided: its start-up stalls
    kernel launches for tens of milliseconds."""

    def __init__(self, device):
        self.device, self.rows, self.thread, self.stop_flag = device, [], None, False
        self.err = None
        self.period = 0.02      # in-process NVML polling does not perturb the kernels (measured, tools/debug_value.py)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001  (older binding name)
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons, time.perf_counter()))
            except Exception as e:      # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(self.period)

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable: " + str(self.err)]}
        nv = self.nv
        busy = [r for r in self.rows if t0 is not None and t0 <= r[2] <= t1] or self.rows[-3:]      # samples inside the timed region
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in busy))
        return {"sm_mhz": statistics.median(r[0] for r in busy), "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(busy)}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method=args.screen_method, theta_cut=None, rng="philox", seed=1234)
    p = C3
    ch = pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]), source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                                          f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))
    ch.path.init_phase_screens()
    ctx = eng.channel_context(ch)
    lib, h = ctx.lib, ctx.handle
    B, S, M, n = args.batch, p["count"], p["m"], p["n"]
    dev = ctx.tdevice
    desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
    edges_d, psd_d = eng.ring_tables(ctx, ch.path.phase_screens[0])
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    pup = np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32)
    pup_d = torch.as_tensor(pup, device=dev)
    steps_total = args.warmup + args.steps
    table_d = torch.zeros((steps_total, B, stride), dtype=torch.float64, device=dev)
    stream = nat.stream_ptr()
    first = rank * steps_total * B          # disjoint global realization indices per rank

    def step_device(i):
        nat.check(lib.pa_simulate_batch_device(h, desc.ref(), B, 1234, first + i * B, nat.ptr(edges_d), nat.ptr(psd_d),
                                               nat.ptr(pup_d), 1, nat.ptr(table_d[i]), stride, stream))

    def barrier():
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    edges = torch.linspace(0, 1, 201, dtype=torch.float64, device=dev)

    def reduce_stats(lo, hi):
        """The only collective of the path: PDT histogram + beam-statistics sums of the realizations of steps
        [lo, hi), all-reduced over the ranks (NCCL)."""
        tab = table_d[lo:hi].reshape(-1, stride)
        hist = torch.zeros(200, dtype=torch.int64, device=dev)
        nat.check(lib.pa_histogram(h, nat.ptr(tab[:, nat.MEASURE_HEAD:]), stride, tab.shape[0], nat.ptr(edges), 200, nat.ptr(hist), stream))
        sums = torch.stack([tab[:, 1].pow(2).sum(), tab[:, 1].pow(4).sum(), tab[:, 3].sum(), tab[:, 3].pow(2).sum()])
        if world > 1:
            import torch.distributed as td
            td.all_reduce(hist)
            td.all_reduce(sums)
        return hist, sums, tab

    # ---- device-resident throughput ------------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up and given time to come up: nvidia-smi's own start-up stalls
    # kernel launches for tens of milliseconds and must not land inside the timed region
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PYATM_BENCH_NOSAMPLER"):
        sampler.start()
        time.sleep(0.2)
    for i in range(args.warmup):
        step_device(i)
    reduce_stats(0, args.warmup)          # also warms up the lazily loaded torch / NCCL kernels of the reduction
    barrier()
    nat.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    e0.record()
    step_events = []
    for i in range(args.steps):
        step_device(args.warmup + i)
        if os.environ.get("PYATM_BENCH_STEP_TIMES"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            step_events.append(ev)
    hist, sums, tab = reduce_stats(args.warmup, steps_total)
    e1.record()
    barrier()
    launches = nat.launch_count()
    if step_events and rank == 0:
        ts = [e0.elapsed_time(ev) for ev in step_events]
        print("per-step ms:", [round(b - a, 2) for a, b in zip([0.0] + ts[:-1], ts)], file=sys.stderr)
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    value = world * args.steps * B / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    rng = np.random.default_rng(rank)
    base = ch.path.phase_screens[0].f_grid.base
    psd = ch.path.phase_screens[0]._get_psd()
    inner = np.insert(base, 0, 0)[:-1]
    n_sets = 4

    def host_set():
        u = rng.random((S, B, 1), dtype=np.float32)
        rho = np.sqrt(inner**2 + u * (base**2 - inner**2)).astype(np.float32)
        th = (2 * np.pi * rng.random((S, B, M))).astype(np.float32)
        cf = ((rng.standard_normal((S, B, M)) + 1j * rng.standard_normal((S, B, M))).astype(np.complex64) * np.sqrt(psd)).astype(np.complex64)
        return [torch.from_numpy(a).pin_memory() for a in ((rho * np.cos(th)).astype(np.float32), (rho * np.sin(th)).astype(np.float32),
                                                           cf.view(np.float32).copy())]

    sets = [host_set() for _ in range(n_sets)]
    out_host = torch.zeros((B, stride), dtype=torch.float64).pin_memory()
    pup_host = torch.from_numpy(pup).pin_memory()

    def step_e2e(i):
        fx, fy, cf = sets[i % n_sets]
        nat.check(lib.pa_simulate_batch(h, desc.ref(), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), 0, 0, None, None, nat.ptr(pup_host),
                                        1, nat.ptr(out_host), stride, stream))

    for i in range(max(1, min(args.warmup, 3))):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    dt_e2e = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dt_e2e = float(t.item())
    e2e_value = world * args.steps * B / dt_e2e
    h2d = 3 * S * B * M * 4 + S * B * M * 4 + pup.nbytes      # fx, fy (4 B) + coef (8 B) per ring + pupil table
    d2h = B * stride * 8

    # ---- roofline of the FFT passes (algorithmic bytes: 4 N^2 8 B per launch = half a split-step stage) -----
    roof = None
    roof_screen = None
    cpu = None
    if rank == 0:
        field = ctx.empty_field(B)
        field.zero_()
        turns = torch.rand((B, n, n), dtype=torch.float32, device=dev) - 0.5
        leg = float(ch.path.leg_lengths()[1])
        reps = 20
        times = {}
        for kind, name in ((0, "k_cols"), (1, "k_rows(ifft*screen*fft)")):
            for _ in range(3):
                nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), leg, p["wvl"], stream))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), leg, p["wvl"], stream))
            b.record()
            torch.cuda.synchronize()
            times[name] = a.elapsed_time(b) / reps * 1e-3
        peak, peak_src = measured_peaks()
        alg = 4 * n * n * 8 * B
        name = max(times, key=times.get)
        traffic = None
        try:        # dram bytes per launch of the same kernels from the committed ncu capture (scaled to this batch)
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            traffic = int(tj[name] * B / tj["batch"])
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": name, "achieved": alg / times[name] / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg / times[name] / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg,
                "per_kernel_us": {k: v * 1e6 for k, v in times.items()},
                "stage_us_per_realization": sum(times.values()) * 1e6 / B,
                "stage_frac_of_hbm_roofline": (8 * n * n * 8 * B) / sum(times.values()) / 1e9 / peak}
        # ---- tensor-pipe roofline of the screen synthesis (pa_screen_ss, tcgen05 path): B screens per call, preparation
        # kernels included.  Executed MMA flops = 3 split-fp16 products x 2 N^2 K2 (K2 = 2 x high rings, padded to 32);
        # algorithmic flops = 4 N^2 M.  Peak = measured dense bf16 cuBLAS throughput (same pipe, same rate).
        roof_screen = None
        try:
            ps0 = ch.path.phase_screens[0]
            m_split, degree = ps0.low_ring_plan()
            method = eng.screen_method(n)
            if method == nat.PA_SCREEN_TC:
                fx_d = torch.empty((B, M), dtype=torch.float32, device=dev)
                fy_d = torch.empty_like(fx_d)
                cf_d = torch.empty((B, M, 2), dtype=torch.float32, device=dev)
                nat.check(lib.pa_rng_spectrum(h, 99, 0, B, 0, 1, M, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(fx_d), nat.ptr(fy_d),
                                              nat.ptr(cf_d), stream))
                bound = eng.coef_bound(ps0._ring_power(), m_split)

                def screens():
                    nat.check(lib.pa_screen_ss(h, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), M, m_split, degree, 0.0, 0.0, B,
                                               nat.ptr(turns), None, 0, method, bound, stream))
                for _ in range(3):
                    screens()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(reps):
                    screens()
                b.record()
                torch.cuda.synchronize()
                t_scr = a.elapsed_time(b) / reps * 1e-3
                k2 = -(-2 * (M - m_split) // 32) * 32
                mma_flops = 3 * 2.0 * n * n * k2 * B
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                        tpeak, tsrc = float(json.load(f)["bf16_tflops"]), "measured cuBLAS bf16 burst (MEASURED_PEAKS.json)"
                except Exception:
                    tpeak, tsrc = 2250.0, "nominal dense bf16 (no MEASURED_PEAKS.json)"
                roof_screen = {"bound": "tensor", "kernel": "pa_screen_ss (k_factors_tc + polynomial nodes + k_screen_tc)",
                               "achieved": mma_flops / t_scr / 1e12, "peak": tpeak, "unit": "TFLOP/s", "frac": mma_flops / t_scr / 1e12 / tpeak,
                               "peak_source": tsrc, "us_per_screen": t_scr * 1e6 / B, "executed_mma_flops_per_screen": mma_flops / B,
                               "algorithmic_flops_per_screen": 4.0 * n * n * M, "rings_in_contraction": int(M - m_split),
                               "rings_as_polynomial": int(m_split)}
        except Exception as e:          # noqa: BLE001  (an extra, never fatal for the bench line)
            roof_screen = {"error": repr(e)}
        # ---- CPU baseline: numpy port of the reference, one realization on one core ------------------------
        if not args.no_cpu:
            _cpu_init(cpu_psd())
            _cpu_realization(0) if args.cpu_warm else None
            t0 = time.perf_counter()
            for seed in range(1, 1 + args.cpu_samples):
                _cpu_realization(seed)
            dt = time.perf_counter() - t0
            cpu = {"value": args.cpu_samples / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_samples} realizations of the same workload (oracle/splitstep.py mode='ref': numpy "
                             "restatement of the reference, pinned by tests/test_oracle_golden.py), single process",
                   "host_cores_available": cpu_cores()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "realizations_per_step_per_gpu": B, "screen_method": args.screen_method,
                       "rng": "device Philox4x32-10 (value) / host-drawn coefficients in pinned memory (e2e)",
                       "l2": f"working set per step {B * (n * n * 12) / 2**20:.0f} MiB of field+screen > 126 MB L2" if B >= 4 else "L2-resident"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "pa_simulate_batch (C ABI, host buffers in, per-realization table out)"},
            "roofline": roof, "roofline_screen": roof_screen, "cpu_baseline": cpu,
            "stats_check": {"hist_total": int(hist.sum().item()), "mean_eta": float(tab[:, nat.MEASURE_HEAD].mean().item())},
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--screen-method", dest="screen_method", default="auto")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-warm", action="store_true")
    ap.add_argument("--cpu-samples", dest="cpu_samples", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 2
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 40
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        run_gpu(args)


if __name__ == "__main__":
    main()
