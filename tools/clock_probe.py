"""Sample SM clock / power while one kernel family runs in a loop (is a kernel power-capped?)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import build_channel
from bench import C3

pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
rows = []; stop = [False]
def poll():
    while not stop[0]:
        rows.append((pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3,
                     pynvml.nvmlDeviceGetCurrentClocksEventReasons(hd)))
        time.sleep(0.02)
pa.gpu.config.update(dtype="complex64", screen_method="auto", theta_cut=None)
ch = build_channel(pa, C3); ch.path.init_phase_screens(); ps = ch.path.phase_screens[0]
ctx = eng.channel_context(ch); dev = ctx.tdevice
NS = 40
np.random.seed(0); fx, fy, cf = eng.draw_spectra_numpy(ch.path, NS)
fx_d = torch.as_tensor(fx[:, 0].copy(), device=dev); fy_d = torch.as_tensor(fy[:, 0].copy(), device=dev)
cf_d = torch.as_tensor(cf[:, 0].copy().view(np.float32), device=dev)
m_split, degree = ps.low_ring_plan(); bound = eng.coef_bound(ps._get_psd(), m_split)
turns = torch.zeros((NS, 2048, 2048), dtype=torch.float32, device=dev)
field = ctx.empty_field(8); field.zero_()
def screens():
    nat.check(ctx.lib.pa_screen_ss(ctx.handle, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), 1024, m_split, degree, 0.0, 0.0, NS,
                                   nat.ptr(turns), None, 0, 1, bound, nat.stream_ptr()))
def ffts():
    for _ in range(10):
        nat.check(ctx.lib.pa_fft_pass(ctx.handle, nat.ptr(field), 8, 0, None, 1e4, 808e-9, nat.stream_ptr()))
        nat.check(ctx.lib.pa_fft_pass(ctx.handle, nat.ptr(field), 8, 1, nat.ptr(turns), 1e4, 808e-9, nat.stream_ptr()))
for name, fn in (("screens(tc)", screens), ("fft passes", ffts)):
    fn(); torch.cuda.synchronize()
    rows.clear(); stop[0] = False
    th = threading.Thread(target=poll, daemon=True); th.start()
    t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < 1.5:
        fn(); n += 1
        if n % 8 == 0: torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    sm = [r[0] for r in rows[5:]]; pw = [r[1] for r in rows[5:]]
    reasons = 0
    for r in rows[5:]: reasons |= r[2]
    print(f"{name:12s}: {a.elapsed_time(b)/n:.3f} ms/call, sm clock median {np.median(sm):.0f} min {min(sm)} MHz, power median {np.median(pw):.0f} max {max(pw):.0f} W, reasons 0x{reasons:x}")
