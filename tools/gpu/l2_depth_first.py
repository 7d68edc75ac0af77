"""Experiment: does a stage (column pass + row pass) get cheaper per field when only a few fields are in flight, so that
the field stays in the 126 MB L2 between the passes?  Times cols,rows,cols,rows,... on B fields in place, the screens read
from a rotating set of buffers (so they come from HBM as in the real step).   (run on the GPU box)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat, gpu

n = int(os.environ.get("N", 2048))
gpu.config.update(use_gpu=True, dtype="complex64")
grid = pa.RectGrid(n, 1.5e-3)
ctx = eng.grid_context(grid)
lib, h = ctx.lib, ctx.handle
stream = nat.stream_ptr()
L, wvl = 1.0e4, 808e-9
out = {}
for B in (1, 2, 3, 4, 6, 8, 16):
    field = ctx.empty_field(B)
    field.zero_()
    turns = [torch.rand((B, n, n), dtype=torch.float32, device="cuda") - 0.5 for _ in range(5)]
    def stages(k):
        for i in range(k):
            nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, 0, nat.ptr(turns[i % 5]), L, wvl, stream))
            nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, 1, nat.ptr(turns[i % 5]), L, wvl, stream))
    stages(5)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    reps = 40
    stages(reps)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3 / B
    out[B] = round(us, 2)
    print(f"B={B}: {us:.2f} us per field per stage ({8 * n * n * 8 / (us * 1e-6) / 6553.9e9:.3f} of the HBM roofline)", flush=True)
    del field, turns
    torch.cuda.empty_cache()
print("RESULT " + json.dumps(out))
