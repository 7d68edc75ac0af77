"""Times the two FFT passes (pa_fft_pass) for every library variant under pyatmosphere_b200/variants/ plus the default build,
each in a child process (PYATM_LIB / extra env), after checking one vacuum leg against torch.fft in complex128.

    python tools/gpu/fft_variants.py [--sizes 2048 ...] [--env NAME=VAL,NAME=VAL ...]     (run on the GPU box)
"""
import argparse
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import json, os, sys
sys.path.insert(0, os.environ["PYATM_ROOT"])
import numpy as np, torch
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat, gpu
sizes = json.loads(os.environ["PYATM_SIZES"])
dtypes = os.environ.get("PYATM_DTYPES", "complex64").split(",")
out = {}
for dtype in dtypes:
    cdt = torch.complex64 if dtype == "complex64" else torch.complex128
    rdt = torch.float32 if dtype == "complex64" else torch.float64
    esz = 8 if dtype == "complex64" else 16
    for n in sizes:
        gpu.config.update(use_gpu=True, dtype=dtype)
        delta, wvl, L = 1.5e-3, 808e-9, 1.0e4
        grid = pa.RectGrid(n, delta)
        ctx = eng.grid_context(grid)
        lib, h = ctx.lib, ctx.handle
        stream = nat.stream_ptr()
        B = max(1, min(8, (2048 * 2048 * 8) // (n * n * (esz // 8))))
        # correctness: one leg on a random field vs torch.fft (complex128)
        g = torch.Generator(device="cuda").manual_seed(n)
        u = torch.randn((1, n, n), dtype=torch.float32, device="cuda", generator=g) + 1j * torch.randn((1, n, n), dtype=torch.float32, device="cuda", generator=g)
        u = u.to(cdt).contiguous()
        f = torch.fft.fftfreq(n, d=delta, dtype=torch.float64, device="cuda")
        ph = -np.pi * L * wvl * (f[:, None] ** 2 + f[None, :] ** 2)
        H = torch.polar(torch.ones_like(ph), ph) * np.exp(1j * ((2 * np.pi / wvl * L) % (2 * np.pi)))
        want = torch.fft.ifft2(H * torch.fft.fft2(u[0].to(torch.complex128)))
        got = u.clone()
        nat.check(lib.pa_vacuum_leg(h, nat.ptr(got), 1, L, wvl, stream))
        torch.cuda.synchronize()
        err = float(torch.linalg.vector_norm(got[0].to(torch.complex128) - want) / torch.linalg.vector_norm(want))
        del want, H, ph, u, got
        field = ctx.empty_field(B); field.zero_()
        turns = torch.rand((B, n, n), dtype=rdt, device="cuda") - 0.5
        res = {"leg_rel_l2": err, "batch": B}
        for kind, name in ((0, "cols"), (1, "rows")):
            for _ in range(3):
                nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), L, wvl, stream))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            reps = 20
            for _ in range(reps):
                nat.check(lib.pa_fft_pass(h, nat.ptr(field), B, kind, nat.ptr(turns), L, wvl, stream))
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / reps * 1e3
            res[name + "_us"] = round(us, 1)
            res[name + "_frac"] = round(4 * n * n * esz * B / (us * 1e-6) / 6553.9e9, 3)
        out[dtype[7:] + "_" + str(n)] = res
        del field, turns
        nat.clear_contexts(); torch.cuda.empty_cache()
print("RESULT " + json.dumps(out))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="*", default=[2048])
    ap.add_argument("--env", nargs="*", default=[], help="extra runs of the default library with these env settings (A=1,B=2)")
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--dtypes", default="complex64")
    args = ap.parse_args()
    runs = [("default", None, {})]
    for spec in args.env:
        runs.append(("default+" + spec, None, dict(kv.split("=") for kv in spec.split(","))))
    for lib in sorted(glob.glob(os.path.join(ROOT, "pyatmosphere_b200", "variants", "libpyatm_*.so"))):
        runs.append((os.path.basename(lib)[9:-3], lib, {}))
    results = {}
    for name, lib, extra in runs:
        if args.only and name not in args.only:
            continue
        env = dict(os.environ, PYATM_ROOT=ROOT, PYATM_SIZES=json.dumps(args.sizes), PYATM_DTYPES=args.dtypes, **extra)
        if lib:
            env["PYATM_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        results[name] = json.loads(line[0][7:]) if line else {"error": (r.stdout + r.stderr)[-1500:]}
        print(name, json.dumps(results[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fft_variants.json"), "a") as f:
        f.write(json.dumps(results) + "\n")


if __name__ == "__main__":
    main()
