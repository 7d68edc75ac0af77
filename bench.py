"""bench.py -- channel realizations/s of the README "advanced channel" (2048^2, 5 SS screens, 50 km) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload c3|c4|c5] [--impl reference]

Workloads (BASELINE.json configs):
  c3 (default, the configuration the metric is quoted on): a step is one batch of B = 160 independent realizations of the
     whole hot path (device RNG -> 5 x [leg + screen synthesis + screen multiply] -> closing leg -> fused measures), one
     pa_simulate_batch_device call (the library works through it in chunks of 32).  `e2e` is the same batch through pa_simulate_batch with HOST coefficient buffers
     (drawn on host threads inside the timed region), copies in and table out.
  c4: Simulation([BeamResult, PDTResult]).run() -- the call a user of the reference makes -- for 6000 device-RNG
     realizations of the same channel, sharded over the ranks, one all-gather of the records at the end.
  c5: long-haul stress, 8192^2, 20 screens over 100 km, batched, complex64 (tensor-core screens) and complex128.
See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import concurrent.futures
import hashlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(n=2048, delta=1.5e-3, wvl=808e-9, w0=0.12, Cn2=5e-16, l0=6e-3, L0=1e3, m=2**10, f_min=1 / 1e3 / 15,
          f_max=1 / 6e-3 * 2, length=50e3, count=5, pupil=0.2)
# config 5: twice the README extent for the twice longer path (6.1 m aperture plane), same spectrum
C5 = dict(C3, n=8192, delta=1.5e-3 * 2048 / 8192 * 2, length=100e3, count=20)
WORKLOAD = "README advanced channel: 2048^2 grid, delta 1.5 mm, 5 SS screens (MVK, 2^10 rings), 50 km, complex64"
WORKLOAD_C5 = "long-haul stress: 8192^2 grid, delta 0.75 mm, 20 SS screens (MVK, 2^10 rings), 100 km"
METRIC = "channel realizations/sec (2048^2, 5 screens)"
UNIT = "realizations/s"
# identical in the b200 and the reference arm (the driver compares the two `config` objects); per-arm details go to `run`
CONFIG = {"workload": WORKLOAD,
          "l2": "inputs larger than L2: one realization sweeps a 32 MiB field and five 16 MiB screens through six legs; the "
                "GPU arm works on chunks of 32 realizations (1.5 GiB against 126 MB of L2)"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_stamp():
    """sha1 over the FFT-pass sources: ties profiles/*traffic.json (an ncu capture) to the kernels it was taken from."""
    h = hashlib.sha1()
    csrc = os.path.join(ROOT, "pyatmosphere_b200", "csrc")
    for name in ("fft_core.cuh", "fft_passes.cuh", "fft_tma.cuh", "fft_split.cuh", "fft_inst.inc"):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


def ncu_traffic(kernel, batch):
    """dram bytes per launch from the committed ncu capture, only if it was taken from the kernels as they are now."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("kernel_source_stamp") != kernel_source_stamp():
            return None, f"profiles/r2_traffic.json is stale (captured at stamp {tj.get('kernel_source_stamp')})"
        return int(tj[kernel] * batch / tj["batch"]), "ncu --set full capture, profiles/r2_traffic.json"
    except Exception as e:          # noqa: BLE001
        return None, f"no capture ({e!r})"


def build_channel(pa, p):
    return pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]), source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                                          f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))


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
    tests/test_oracle_golden.py, at this very configuration too) with one realization per worker process per step, on
    all host cores (capped)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = cpu_cores()
    workers = max(1, min(cores, int(os.environ.get("PYATM_REF_WORKERS", "32"))))
    psd = cpu_psd()
    ctx = mp.get_context("fork")
    etas = []
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(psd,)) as pool:
        seed = 1                      # seeds 1, 2, 3, ...: the GPU arm replays seeds 1..3 on the same draws (stats_check)
        for _ in range(args.warmup):
            pool.map(_cpu_realization, range(seed, seed + workers))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            etas += pool.map(_cpu_realization, range(seed, seed + workers))
            seed += workers
        dt = time.perf_counter() - t0
    value = workers * args.steps / dt
    sample = f"{workers} realizations per step (one per worker process), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (numpy: c128 transfer-function product, as the reference)", "data": "synthetic",
        "config": CONFIG, "run": {"realizations_per_step": workers, "worker_processes": workers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "stats_check": {"seeds": [1, len(etas)], "eta_first3": etas[:3], "mean_eta": float(np.mean(etas)),
                        "sem_eta": float(np.std(etas, ddof=1) / np.sqrt(len(etas))) if len(etas) > 1 else None},
    }
    emit(line)


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
        busy = [r for r in self.rows if t0 is not None and t0 <= r[2] <= t1] or self.rows[-3:]      # samples inside the timed region
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in busy))
        return {"sm_mhz": statistics.median(r[0] for r in busy), "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(busy)}


class Dist:
    """Rank bookkeeping + the barrier / max-over-ranks helpers of the timing contract."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as td
            td.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.td = td

    def barrier(self):
        if self.world > 1:
            self.td.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.td.destroy_process_group()


def time_pass(lib, h, nat, torch, field, B, kind, turns, leg, wvl, stream, reps=20):
    for _ in range(3):
        nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), leg, wvl, stream))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), leg, wvl, stream))
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


# ---------------------------------------------------------------------------------------------------------------
# workload c3: the headline line
# ---------------------------------------------------------------------------------------------------------------
class HostDraws:
    """Coefficient sets of one step drawn on host threads in the reference's way (grids.py:98-107, phase_screens.py:98-103:
    one radius fraction per screen, M angles, 2M normals) straight into pinned buffers, `depth` steps ahead of the GPU."""

    def __init__(self, torch, S, B, M, base, psd, seed, depth=3, threads=3):
        self.S, self.B, self.M = S, B, M
        self.base = np.asarray(base, dtype=np.float32)
        self.inner = np.insert(self.base, 0, 0)[:-1]
        self.amp = np.sqrt(np.asarray(psd, dtype=np.float32))
        self.sets = [[torch.empty((S, B, M), dtype=torch.float32).pin_memory(), torch.empty((S, B, M), dtype=torch.float32).pin_memory(),
                      torch.empty((S, B, M, 2), dtype=torch.float32).pin_memory()] for _ in range(depth + 1)]
        self.rngs = [np.random.default_rng([seed, i]) for i in range(len(self.sets))]
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=threads)
        self.pending = []
        self.issued = 0
        self.draw_seconds = 0.0

    def _draw(self, k):
        t0 = time.perf_counter()
        rng = self.rngs[k]
        fx, fy, cf = (t.numpy() for t in self.sets[k])
        S, B, M = self.S, self.B, self.M
        u = rng.random((S, B, 1), dtype=np.float32)
        rho = np.sqrt(self.inner**2 + u * (self.base**2 - self.inner**2), dtype=np.float32)
        th = rng.random((S, B, M), dtype=np.float32)
        th *= np.float32(2 * np.pi)
        np.multiply(rho, np.cos(th), out=fx)
        np.multiply(rho, np.sin(th), out=fy)
        rng.standard_normal(dtype=np.float32, out=cf)
        cf *= self.amp[:, None]
        self.draw_seconds += time.perf_counter() - t0
        return k

    def submit(self):
        k = self.issued % len(self.sets)
        self.issued += 1
        self.pending.append(self.pool.submit(self._draw, k))

    def next(self):
        """Block until the oldest submitted set is drawn; returns its three pinned tensors."""
        k = self.pending.pop(0).result()
        return self.sets[k]

    def close(self):
        self.pool.shutdown(wait=True)


def run_c3(args):
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    d = Dist()
    world, rank, local = d.world, d.rank, d.local
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method=args.screen_method, theta_cut=None, rng="philox", seed=1234)
    p = C3
    ch = build_channel(pa, p)
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
    ps0 = ch.path.phase_screens[0]

    def step_device(i):
        nat.check(lib.pa_simulate_batch_device(h, desc.ref(), B, 1234, first + i * B, nat.ptr(edges_d), nat.ptr(psd_d),
                                               nat.ptr(pup_d), 1, nat.ptr(table_d[i]), stride, stream))

    edges = torch.linspace(0, 1, 201, dtype=torch.float64, device=dev)
    from pyatmosphere_b200.distributed import StatsComm
    comm = StatsComm() if world > 1 else None

    def reduce_stats(lo, hi):
        """The only collective of the path: PDT histogram + beam-statistics sums of the realizations of steps
        [lo, hi), all-reduced over the ranks (NCCL)."""
        tab = table_d[lo:hi].reshape(-1, stride)
        hist = torch.zeros(200, dtype=torch.int64, device=dev)
        nat.check(lib.pa_histogram(h, nat.ptr(tab[:, nat.MEASURE_HEAD:]), stride, tab.shape[0], nat.ptr(edges), 200, nat.ptr(hist), stream))
        sums = torch.stack([tab[:, 1].pow(2).sum(), tab[:, 1].pow(4).sum(), tab[:, 3].sum(), tab[:, 3].pow(2).sum(),
                            tab[:, nat.MEASURE_HEAD].sum(), tab[:, nat.MEASURE_HEAD].pow(2).sum()])
        if world > 1:
            comm.allreduce(hist, sums)          # pa_stats_allreduce: NCCL all-reduce behind the C ABI (include/pyatm_b200.h)
        return hist, sums, tab

    def rooflines():
        """Each pass / the screen synthesis timed ALONE with CUDA events on the launching stream, right after the warm-up
        (burst conditions, like the copy that MEASURED_PEAKS.json's hbm_gbs comes from; inside the long timed loops the
        board runs into its power cap -- see `clocks`)."""
        roof = roof_screen = None
        # ---- roofline of the FFT passes (algorithmic bytes: 4 N^2 8 B per launch = half a split-step stage) -----
        Br = 8
        field = ctx.empty_field(Br)
        field.zero_()
        turns = torch.rand((Br, n, n), dtype=torch.float32, device=dev) - 0.5
        leg = float(ch.path.leg_lengths()[1])
        times = {name: time_pass(lib, h, nat, torch, field, Br, kind, turns, leg, p["wvl"], stream)
                 for kind, name in ((0, "k_cols"), (1, "k_rows(ifft*screen*fft)"))}
        peak, peak_src = measured_peaks()
        alg = 4 * n * n * 8 * Br
        name = max(times, key=times.get)
        traffic, traffic_src = ncu_traffic(name, Br)
        roof = {"bound": "hbm", "kernel": name, "achieved": alg / times[name] / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg / times[name] / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "fields_per_launch": Br,
                "per_kernel_us": {k: v * 1e6 for k, v in times.items()},
                "per_kernel_frac": {k: alg / v / 1e9 / peak for k, v in times.items()},
                "stage_us_per_realization": sum(times.values()) * 1e6 / Br,
                "stage_frac_of_hbm_roofline": (8 * n * n * 8 * Br) / sum(times.values()) / 1e9 / peak}
        # ---- tensor-pipe roofline of the screen synthesis (pa_screen_ss, tcgen05 path): 32 screens per call -- what one path
        # position of a chunk of the step asks for -- preparation kernels included.  Executed MMA flops = 3 split-fp16 products
        # x 2 N^2 K2 (K2 = 2 x high rings, padded to 32); algorithmic flops = 4 N^2 M.  Peak = measured dense
        # bf16 cuBLAS throughput (same pipe, same rate).
        try:
            m_split, degree = ps0.low_ring_plan()
            method = eng.screen_method(n)
            if method == nat.PA_SCREEN_TC:
                Bs = 32
                turns_s = torch.empty((Bs, n, n), dtype=torch.float32, device=dev)
                fx_d = torch.empty((Bs, M), dtype=torch.float32, device=dev)
                fy_d = torch.empty_like(fx_d)
                cf_d = torch.empty((Bs, M, 2), dtype=torch.float32, device=dev)
                nat.check(lib.pa_rng_spectrum(h, 99, 0, Bs, 0, 1, M, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(fx_d), nat.ptr(fy_d),
                                              nat.ptr(cf_d), stream))
                bound = eng.coef_bound(ps0._ring_power(), m_split)

                def screens():
                    nat.check(lib.pa_screen_ss(h, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), M, m_split, degree, 0.0, 0.0, Bs,
                                               nat.ptr(turns_s), None, 0, method, bound, stream))
                for _ in range(3):
                    screens()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(10):
                    screens()
                b.record()
                torch.cuda.synchronize()
                t_scr = a.elapsed_time(b) / 10 * 1e-3
                del turns_s
                k2 = -(-2 * (M - m_split) // 32) * 32
                mma_flops = 3 * 2.0 * n * n * k2 * Bs
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                        tpeak, tsrc = float(json.load(f)["bf16_tflops"]), "measured cuBLAS bf16 burst (MEASURED_PEAKS.json)"
                except Exception:
                    tpeak, tsrc = 2250.0, "nominal dense bf16 (no MEASURED_PEAKS.json)"
                roof_screen = {"bound": "tensor", "kernel": "pa_screen_ss (k_factors_tc + polynomial nodes + k_screen_tc)",
                               "achieved": mma_flops / t_scr / 1e12, "peak": tpeak, "unit": "TFLOP/s", "frac": mma_flops / t_scr / 1e12 / tpeak,
                               "peak_source": tsrc, "screens_per_call": Bs, "us_per_screen": t_scr * 1e6 / Bs,
                               "executed_mma_flops_per_screen": mma_flops / Bs,
                               "algorithmic_flops_per_screen": 4.0 * n * n * M, "rings_in_contraction": int(M - m_split),
                               "rings_as_polynomial": int(m_split)}
        except Exception as e:          # noqa: BLE001  (an extra, never fatal for the bench line)
            roof_screen = {"error": repr(e)}
        del field, turns
        return roof, roof_screen

    # ---- device-resident throughput ------------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up and given time to come up
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PYATM_BENCH_NOSAMPLER"):
        sampler.start()
        time.sleep(0.2)
    # the pass rooflines first, on an idle board at its boost clocks: a kernel timed alone (7 ms of launches), like the copy
    # behind MEASURED_PEAKS.json; the long loops below run into the power cap (see `clocks`)
    time.sleep(0.3)
    roof, roof_screen = rooflines() if rank == 0 else (None, None)
    d.barrier()
    for i in range(args.warmup):
        step_device(i)
    reduce_stats(0, args.warmup)          # also warms up the lazily loaded torch / NCCL kernels of the reduction
    d.barrier()
    nat.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.barrier()
    t_begin = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step_device(args.warmup + i)
    hist, sums, tab = reduce_stats(args.warmup, steps_total)
    e1.record()
    d.barrier()
    launches = nat.launch_count()
    ms = d.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    value = world * args.steps * B / (ms * 1e-3)
    n_real = world * args.steps * B
    mean_eta = float(sums[4].item()) / n_real
    sem_eta = float(np.sqrt(max(float(sums[5].item()) / n_real - mean_eta**2, 0.0) / n_real))

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    draws = HostDraws(torch, S, B, M, ps0.f_grid.base, ps0._get_psd(), seed=1000 + rank)
    outs = [torch.zeros((B, stride), dtype=torch.float64).pin_memory() for _ in range(2)]
    pup_host = torch.from_numpy(pup).pin_memory()
    consumed = {"eta_sum": 0.0, "records": 0}

    def enqueue_e2e(bufs, out):
        fx, fy, cf = bufs
        nat.check(lib.pa_simulate_batch_async(h, desc.ref(), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), 0, 0, None, None, nat.ptr(pup_host),
                                              1, nat.ptr(out), stride, stream))

    def consume(out):
        """The step's records, read on the host (they are what a caller of the reference gets back per realization)."""
        consumed["eta_sum"] += float(out[:, nat.MEASURE_HEAD].sum())
        consumed["records"] += out.shape[0]

    def e2e_loop(steps, live):
        """Every step: coefficients in pinned host memory -> pa_simulate_batch_async (copies in, batch, table out) -> records
        read on the host.  The host enqueues step i+1 before it waits for step i (two result buffers), so the GPU stays busy
        across the call boundary.  `live`: every step's coefficients are drawn inside the timed region (host threads, up to
        3 steps ahead); otherwise four sets drawn beforehand are cycled (a caller with its own generator)."""
        if live:
            for _ in range(min(3, steps)):
                draws.submit()
        events = []
        d.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            bufs = draws.next() if live else draws.sets[i % len(draws.sets)]
            enqueue_e2e(bufs, outs[i % 2])
            ev = torch.cuda.Event()
            ev.record()
            events.append(ev)
            if i >= 1:
                events[i - 1].synchronize()
                consume(outs[(i - 1) % 2])
            if live and i + 3 < steps:
                draws.submit()           # into the set of step i-1, which has just been waited for
        nat.check(lib.pa_stream_synchronize(h, stream))
        consume(outs[(steps - 1) % 2])
        d.barrier()
        return d.max_over_ranks(time.perf_counter() - t0)

    for k in range(len(draws.sets)):
        draws._draw(k)
    for i in range(max(1, min(args.warmup, 3))):
        enqueue_e2e(draws.sets[i % len(draws.sets)], outs[i % 2])
    nat.check(lib.pa_stream_synchronize(h, stream))
    draws.draw_seconds = 0.0
    dt_live = e2e_loop(args.steps, live=True)
    draw_ms = 1e3 * draws.draw_seconds / args.steps
    dt_pre = e2e_loop(args.steps, live=False)
    draws.close()
    e2e_value = world * args.steps * B / dt_live
    h2d = 3 * S * B * M * 4 + S * B * M * 4 + pup.nbytes      # fx, fy (4 B) + coef (8 B) per ring + pupil table
    d2h = B * stride * 8

    cpu = None
    stats = {"hist_total": int(hist.sum().item()), "mean_eta": mean_eta, "sem_eta": sem_eta, "realizations": n_real}
    if rank == 0:
        # ---- CPU baseline: numpy port of the reference, one realization per seed on one core; the SAME seeds are then
        # replayed on the GPU from the same numpy draws and the transmittances compared (reference's complex64 floor)
        if not args.no_cpu:
            _cpu_init(cpu_psd())
            _cpu_realization(0) if args.cpu_warm else None
            seeds = list(range(1, 1 + args.cpu_samples))
            t0 = time.perf_counter()
            cpu_eta = [_cpu_realization(seed) for seed in seeds]
            dt = time.perf_counter() - t0
            cpu = {"value": args.cpu_samples / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_samples} realizations of the same workload (oracle/splitstep.py mode='ref': numpy "
                             "restatement of the reference, pinned by tests/test_oracle_golden.py), single process",
                   "host_cores_available": cpu_cores()}
            pa.gpu.config.update(rng="numpy")
            gpu_eta = []
            for seed in seeds:
                np.random.seed(seed)
                rec = eng.simulate_realizations(ch, 0, 1, np.arange(1), [p["pupil"]], [])
                gpu_eta.append(float(rec[0, len(nat.MEASURE_NAMES)]))
            pa.gpu.config.update(rng="philox")
            rel = float(np.max(np.abs(np.array(gpu_eta) - np.array(cpu_eta)) / np.array(cpu_eta)))
            stats.update({"seeds_replayed": seeds, "cpu_eta": cpu_eta, "gpu_eta_same_draws": gpu_eta, "max_rel_diff": rel,
                          "tolerance": 5e-3})
            if not rel < 5e-3:
                raise AssertionError(f"GPU and CPU-reference transmittances differ on the same draws: {stats}")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64",
            "data": "synthetic",
            "config": CONFIG,
            "run": {"realizations_per_step_per_gpu": B, "screen_method": args.screen_method,
                    "rng": "device Philox4x32-10 (value) / host-drawn coefficients in pinned memory (e2e)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "pa_simulate_batch_async + pa_stream_synchronize (C ABI, host buffers in, per-realization table out and "
                           "read on the host every step, step i+1 enqueued before step i is waited for), coefficients drawn on 3 "
                           "host threads inside the timed region",
                    "records_read_on_host": consumed["records"],
                    "host_draw_ms_per_step": draw_ms, "value_with_predrawn_coefficients": world * args.steps * B / dt_pre},
            "roofline": roof, "roofline_screen": roof_screen, "cpu_baseline": cpu, "stats_check": stats,
        }
        emit(line)
    if comm is not None:
        comm.close()
    d.close()


# ---------------------------------------------------------------------------------------------------------------
# workload c4: Simulation([BeamResult, PDTResult]).run(), 6000 realizations, sharded
# ---------------------------------------------------------------------------------------------------------------
def run_c4(args):
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import distributed as pdist, _native as nat
    d = Dist()
    total = args.total
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method=args.screen_method, theta_cut=None, rng="philox", seed=4242,
                         batch=64)
    ch = build_channel(pa, C3)

    def job(count):
        beam = pa.simulations.BeamResult(ch, max_size=count)
        pdt = pa.simulations.PDTResult(ch, max_size=count)
        sim = pa.simulations.Simulation([beam, pdt])
        sim.run()
        return beam, pdt, sim

    for _ in range(max(1, args.warmup)):
        job(64 * d.world)         # at least two of the library's chunks per rank: every workspace has its final size
    sampler = ClockSampler(d.local)
    if d.rank == 0:
        sampler.start()
        time.sleep(0.2)
    d.barrier()
    nat.launch_count(reset=True)
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        beam, pdt, sim = job(total)
    d.barrier()
    dt = d.max_over_ranks(time.perf_counter() - t_begin)
    launches = nat.launch_count()
    clocks = sampler.stop(t_begin, time.perf_counter()) if d.rank == 0 else None
    # statistics two ways: from the gathered records (what BeamResult / PDTResult report) and from this rank's share
    # reduced with ONE all-reduce of [200-bin histogram] + [n, sum, sum of squares] (north_star's final all-reduce)
    cols = sim.last_columns
    red = pdist.reduce_statistics(sim.last_local_table.cpu().numpy(), cols, eta_names=[("fixed", C3["pupil"])])
    hist = pdt.histogram()
    ok = bool(np.array_equal(red[("hist", ("fixed", C3["pupil"]))], hist)) and all(
        np.isclose(red[k][0], getattr(beam, k)[0], rtol=1e-12) and np.isclose(red[k][1], getattr(beam, k)[1], rtol=1e-9) for k in ("bw", "lt", "st"))
    records = np.array([m.data for m in beam.measures] + [pdt.measures[0].data])
    digest = hashlib.sha1(np.ascontiguousarray(records).tobytes()).hexdigest()
    if d.rank == 0:
        value = total * args.steps / dt
        emit({
            "metric": METRIC + " through Simulation([BeamResult, PDTResult]).run()", "value": value, "unit": UNIT, "n_gpus": d.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
            "config": {"workload": "config 4: " + WORKLOAD + f"; {total} realizations sharded over the ranks, device RNG keyed by the "
                       "global realization index, one all-gather of the per-sample table at the end", "timing": "wall clock, max over ranks",
                       "realizations_per_step": total},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(records.nbytes),
                    "api": "pyatmosphere.simulations.Simulation.run (device RNG: no per-step input to copy)"},
            "statistics": {"sigma_bw": beam.bw, "sigma_lt": beam.lt, "w_st": beam.st, "mean_eta": float(np.mean(pdt.measures[0].data)),
                           "hist_nonzero_bins": int((hist > 0).sum()), "allreduced_statistics_equal_gathered": ok,
                           "records_sha1": digest},
        })
    assert ok, "all-reduced statistics differ from the statistics of the gathered records"
    d.close()


# ---------------------------------------------------------------------------------------------------------------
# workload c5: 8192^2, 20 screens, 100 km, complex64 (tensor-core screens) and complex128, batched
# ---------------------------------------------------------------------------------------------------------------
def run_c5(args):
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    d = Dist()
    p = C5
    n, S, M = p["n"], p["count"], p["m"]
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    pup = np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32)
    peak, peak_src = measured_peaks()
    out = {}
    fields = {}
    clocks = None
    launches = 0
    for dtype, B in (("complex64", args.batch), ("complex128", max(1, args.batch // 2))):
        nat.clear_contexts()
        torch.cuda.empty_cache()
        pa.gpu.config.update(use_gpu=True, dtype=dtype, screen_method="auto", theta_cut=None, rng="philox", seed=77)
        ch = build_channel(pa, p)
        ch.path.init_phase_screens()
        ctx = eng.channel_context(ch)
        lib, h, dev = ctx.lib, ctx.handle, ctx.tdevice
        desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
        edges_d, psd_d = eng.ring_tables(ctx, ch.path.phase_screens[0])
        pup_d = torch.as_tensor(pup, device=dev)
        steps_total = args.warmup + args.steps
        table_d = torch.zeros((steps_total, B, stride), dtype=torch.float64, device=dev)
        stream = nat.stream_ptr()
        first = d.rank * steps_total * B

        def step(i):
            nat.check(lib.pa_simulate_batch_device(h, desc.ref(), B, 77, first + i * B, nat.ptr(edges_d), nat.ptr(psd_d),
                                                   nat.ptr(pup_d), 1, nat.ptr(table_d[i]), stride, stream))
        sampler = ClockSampler(d.local)
        if d.rank == 0 and dtype == "complex64":
            sampler.start()
            time.sleep(0.2)
        for i in range(args.warmup):
            step(i)
        d.barrier()
        nat.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            step(args.warmup + i)
        e1.record()
        d.barrier()
        ms = d.max_over_ranks(e0.elapsed_time(e1))
        if dtype == "complex64":
            launches = nat.launch_count()
            clocks = sampler.stop(t_begin, time.perf_counter()) if d.rank == 0 else None
        eta = table_d[args.warmup:, :, 0].mean().item()
        out[dtype] = {"value": d.world * args.steps * B / (ms * 1e-3), "ms_per_step": ms / args.steps, "realizations_per_step_per_gpu": B,
                      "screens": "tcgen05 split-fp16" if eng.screen_method(n) == nat.PA_SCREEN_TC else "float64 CUDA cores",
                      "mean_total_power": eta}
        if d.rank == 0:
            # one realization's field on the same coefficients in both precisions (tolerance study), and the pass rooflines
            fx = torch.empty((S, 1, M), dtype=torch.float32, device=dev)
            fy, cf = torch.empty_like(fx), torch.empty((S, 1, M, 2), dtype=torch.float32, device=dev)
            nat.check(lib.pa_rng_spectrum(h, 77, 10**6, 1, 0, S, M, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), stream))
            field = ctx.empty_field(1)
            nat.check(lib.pa_propagate(h, desc.ref(), nat.ptr(field), 1, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), stream))
            fields[dtype] = field[0].to(torch.complex128).cpu()
            rdt = torch.float32 if dtype == "complex64" else torch.float64
            turns = torch.rand((1, n, n), dtype=rdt, device=dev) - 0.5
            leg = float(ch.path.leg_lengths()[1])
            esz = 8 if dtype == "complex64" else 16
            alg = 4 * n * n * esz
            times = {name: time_pass(lib, h, nat, torch, field, 1, kind, turns, leg, p["wvl"], stream, reps=10)
                     for kind, name in ((0, "column pass (k_col_outer + k_cols_tma<256> + k_col_outer)"), (1, "k_rows(ifft*screen*fft)"))}
            out[dtype]["roofline"] = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                                      "per_kernel_us": {k: v * 1e6 for k, v in times.items()},
                                      "per_kernel_frac": {k: alg / v / 1e9 / peak for k, v in times.items()}}
            del field, turns, fx, fy, cf
        del table_d
    if d.rank == 0:
        a, b = fields["complex64"], fields["complex128"]
        rel = float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
        c64 = out["complex64"]
        rl = c64.pop("roofline")
        name = max(rl["per_kernel_us"], key=rl["per_kernel_us"].get)
        roof = dict(rl, kernel=name, achieved=rl["algorithmic_bytes_per_launch"] / rl["per_kernel_us"][name] / 1e3,
                    frac=rl["per_kernel_frac"][name], traffic=None)
        emit({
            "metric": "channel realizations/sec (8192^2, 20 screens, 100 km)", "value": c64["value"], "unit": UNIT, "n_gpus": d.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": c64["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
            "config": {"workload": "config 5: " + WORKLOAD_C5, "realizations_per_step_per_gpu": c64["realizations_per_step_per_gpu"],
                       "rng": "device Philox4x32-10", "l2": "one 8192^2 complex64 field is 512 MiB > 126 MB L2"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof,
            "complex64": c64, "complex128": out["complex128"],
            "complex64_vs_complex128_rel_l2": rel, "stated_tolerance_complex64_config5": 2.5e-5,
        })
    d.close()


_REAL_STDOUT = None


def keep_stdout_for_the_json_line():
    """Everything any library writes to file descriptor 1 during the run (NCCL prints its version there when NCCL_DEBUG is
    set) goes to stderr; emit() puts the one JSON line on the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5"])
    ap.add_argument("--total", type=int, default=6000, help="c4: realizations per Simulation.run()")
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
    elif args.workload == "c4":
        args.steps = args.steps if args.steps is not None else 1
        args.warmup = args.warmup if args.warmup is not None else 3
        run_c4(args)
    elif args.workload == "c5":
        args.steps = args.steps if args.steps is not None else 4
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        args.batch = args.batch or 4
        run_c5(args)
    else:
        args.steps = args.steps if args.steps is not None else 24
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        args.batch = args.batch or 160
        run_c3(args)


if __name__ == "__main__":
    main()
