"""Launches one kind of kernel a few times for ncu: python tools/gpu/prof_pass.py cols|rows|screens|step [n] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pyatmosphere_b200 as pa  # noqa: E402
from pyatmosphere_b200 import _engine as eng, _native as nat  # noqa: E402
from bench import C3, build_channel  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="auto", theta_cut=None, rng="philox", seed=1)
p = dict(C3, n=n, delta=C3["delta"] * 2048 / n)
ch = build_channel(pa, p)
ch.path.init_phase_screens()
ctx = eng.channel_context(ch)
lib, h, stream = ctx.lib, ctx.handle, nat.stream_ptr()
field = ctx.empty_field(B)
field.zero_()
turns = torch.rand((B, n, n), dtype=torch.float32, device="cuda") - 0.5
if what in ("cols", "rows"):
    for _ in range(reps):
        nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, 0 if what == "cols" else 1, nat.ptr(turns), 1.0e4, 808e-9, stream))
elif what == "screens":
    ps0 = ch.path.phase_screens[0]
    M = p["m"]
    edges_d, psd_d = eng.ring_tables(ctx, ps0)
    fx = torch.empty((B, M), dtype=torch.float32, device="cuda")
    fy, cf = torch.empty_like(fx), torch.empty((B, M, 2), dtype=torch.float32, device="cuda")
    nat.check(lib.pa_rng_spectrum(h, 99, 0, B, 0, 1, M, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), stream))
    m_split, degree = ps0.low_ring_plan()
    for _ in range(reps):
        nat.check(lib.pa_screen_ss(h, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), M, m_split, degree, 0.0, 0.0, B, nat.ptr(turns), None, 0,
                                   eng.screen_method(n), eng.coef_bound(ps0._ring_power(), m_split), stream))
else:       # whole steps of B realizations
    desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
    edges_d, psd_d = eng.ring_tables(ctx, ch.path.phase_screens[0])
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    pup_d = torch.as_tensor(np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32), device="cuda")
    table = torch.zeros((B, stride), dtype=torch.float64, device="cuda")
    for i in range(reps):
        nat.check(lib.pa_simulate_batch_device(h, desc.ref(), B, 1, i * B, nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(pup_d), 1, nat.ptr(table),
                                               stride, stream))
torch.cuda.synchronize()
print("done", what, n, B)
