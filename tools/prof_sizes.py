"""Per-pass timing of the FFT pipeline at every grid size (profiling helper; prints µs and the roofline fraction).

    python tools/prof_sizes.py [sizes...]        # default 1024 2048 4096 8192
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat, gpu


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]
    peak = 6500.0
    try:
        with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        peak = float(mp.get("hbm_gbs") or peak)
    except Exception:
        pass
    out = {}
    for dtype in ("complex64", "complex128"):
        gpu.config["dtype"] = dtype
        for n in sizes:
            B = max(1, min(8, (2048 * 2048 * 8) // (n * n)))
            grid = pa.RectGrid(n, 0.0015)
            ctx = eng.grid_context(grid)
            lib, h = ctx.lib, ctx.handle
            field = ctx.empty_field(B)
            field.zero_()
            rdt = torch.float32 if dtype == "complex64" else torch.float64
            turns = torch.rand((B, n, n), dtype=rdt, device="cuda") - 0.5
            stream = torch.cuda.current_stream().cuda_stream
            esz = 8 if dtype == "complex64" else 16
            res = {}
            for kind, name in ((0, "cols"), (1, "rows")):
                for _ in range(3):
                    nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), 1000.0, 808e-9, stream))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                reps = 10
                for _ in range(reps):
                    nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), 1000.0, 808e-9, stream))
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) / reps * 1e3
                alg = 4 * n * n * esz * B
                res[name] = {"us": round(us, 1), "frac": round(alg / (us * 1e-6) / 1e9 / peak, 3)}
            out[f"{dtype}_{n}_b{B}"] = res
            print(dtype, n, "batch", B, res, flush=True)
            del field, turns
            torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/prof_sizes.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
