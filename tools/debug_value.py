"""Why is the device-RNG loop slower than the host-coefficient loop?  Times the pieces with CUDA events."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat
from bench import C3
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import build_channel

pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="auto", theta_cut=None, rng="philox", seed=1234)
ch = build_channel(pa, C3)
ch.path.init_phase_screens()
ctx = eng.channel_context(ch)
lib, h = ctx.lib, ctx.handle
B, S, M, n = 8, 5, 1024, 2048
dev = ctx.tdevice
desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
edges_d, psd_d = eng.ring_tables(ctx, ch.path.phase_screens[0])
stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
pup_d = torch.as_tensor(np.array([[np.float32(0.04), 0, 0]], dtype=np.float32), device=dev)
table = torch.zeros((B, stride), dtype=torch.float64, device=dev)
field = ctx.empty_field(B)
fx = torch.empty((S, B, M), dtype=torch.float32, device=dev); fy = torch.empty_like(fx)
cf = torch.empty((S, B, M, 2), dtype=torch.float32, device=dev)
st = nat.stream_ptr()

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

cnt = [0]
def full():
    cnt[0] += 1
    nat.check(lib.pa_simulate_batch_device(h, desc.ref(), B, 1234, cnt[0] * B, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(pup_d), 1, nat.ptr(table), stride, st))
def rng():
    cnt[0] += 1
    nat.check(lib.pa_rng_spectrum(h, 1234, cnt[0] * B, B, 0, S, M, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), st))
def prop():
    nat.check(lib.pa_propagate(h, desc.ref(), nat.ptr(field), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), st))
def meas():
    nat.check(lib.pa_measure(h, nat.ptr(field), B, nat.ptr(pup_d), 1, 0, nat.ptr(table), stride, st))
print("full step (device rng)  ms:", timeit(full))
print("rng only                ms:", timeit(rng))
print("propagate (device coefs) ms:", timeit(prop))
print("measure                 ms:", timeit(meas))
# host-drawn coefficients for comparison
np.random.seed(0)
hfx, hfy, hcf = eng.draw_spectra_numpy(ch.path, B)
fx.copy_(torch.as_tensor(np.ascontiguousarray(hfx.transpose(1, 0, 2)))); fy.copy_(torch.as_tensor(np.ascontiguousarray(hfy.transpose(1, 0, 2))))
cf.copy_(torch.as_tensor(np.ascontiguousarray(hcf.transpose(1, 0, 2)).view(np.float32).reshape(S, B, M, 2)))
print("propagate (numpy coefs)  ms:", timeit(prop))
rng()
print("fx stats device:", float(fx.abs().max()), float(fx.std()), "coef absmax", float(cf.abs().max()), "nan?", bool(torch.isnan(cf).any()))
print("numpy coef absmax", np.abs(hcf).max(), "fx absmax", np.abs(hfx).max())

# ---- how much does NVML polling perturb the GPU? ----
import threading, pynvml
pynvml.nvmlInit()
hd = pynvml.nvmlDeviceGetHandleByIndex(0)
def poll(kind, period, stop):
    while not stop[0]:
        if kind in ("clock", "both"): pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM)
        if kind in ("reasons", "both"): pynvml.nvmlDeviceGetCurrentClocksEventReasons(hd)
        if kind == "power": pynvml.nvmlDeviceGetPowerUsage(hd)
        time.sleep(period)
for kind, period in (("none", 1), ("clock", 0.05), ("reasons", 0.05), ("power", 0.05), ("both", 0.2), ("both", 0.05)):
    stop = [False]
    th = threading.Thread(target=poll, args=(kind, period, stop), daemon=True)
    if kind != "none": th.start()
    time.sleep(0.3)
    ms = timeit(full, reps=60)
    stop[0] = True
    if kind != "none": th.join()
    print(f"polling {kind:8s} every {period*1e3:.0f} ms: step {ms:.3f} ms")
