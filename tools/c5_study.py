"""Config 5 of BASELINE.json: long-haul stress -- 8192^2 grid, 20 phase screens over 100 km, complex128 vs complex64
tolerance study.  Runs the same seeded realization in both precisions (plus complex64 with exact screens) and prints
the relative L2 differences, the power budget and timings.  Writes profiles/r1_c5_study.json.

    python tools/c5_study.py [--n 8192] [--screens 20] [--length 100e3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pyatmosphere_b200 as pa  # noqa: E402


def channel(n, screens, length, rings):
    delta = 1.5e-3 * 2048 / n * 2            # 6.1 m aperture plane: twice the README extent for the longer path
    return pa.Channel(
        grid=pa.RectGrid(resolution=n, delta=delta), source=pa.GaussianSource(wvl=808e-9, w0=0.12, F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=5e-16, l0=6e-3, L0=1e3),
                                          f_grid=pa.RandLogPolarGrid(points=rings, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)),
            length=length, count=screens),
        pupil=pa.CirclePupil(radius=0.2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--screens", type=int, default=20)
    ap.add_argument("--length", type=float, default=100e3)
    ap.add_argument("--rings", type=int, default=2**10)
    args = ap.parse_args()
    fields, report = {}, {"n": args.n, "screens": args.screens, "length_m": args.length, "rings": args.rings, "runs": {}}
    for tag, cfg in (("complex128", dict(dtype="complex128", screen_method="exact", theta_cut=2.0)),
                     ("complex64_exact_screens", dict(dtype="complex64", screen_method="exact", theta_cut=2.0)),
                     ("complex64_auto", dict(dtype="complex64", screen_method="auto", theta_cut=None))):
        pa.gpu.config.update(use_gpu=True, rng="numpy", **cfg)
        ch = channel(args.n, args.screens, args.length, args.rings)
        for rep in range(2):                   # second run = warm timing
            np.random.seed(2026)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = ch.run(pupil=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        eta_tot = pa.measures.eta(ch, output=out)
        eta_ap = pa.measures.eta(ch, output=ch.pupil.output(out))
        fields[tag] = out.get()
        report["runs"][tag] = {"seconds_warm": dt, "total_power": eta_tot, "eta_aperture": eta_ap,
                               "rytov2": ch.get_rythov2()}
        print(tag, report["runs"][tag], flush=True)
        del out
        from pyatmosphere_b200 import _native as nat
        nat.clear_contexts()
        torch.cuda.empty_cache()
    ref = fields["complex128"]
    nrm = np.linalg.norm(ref)
    for tag in ("complex64_exact_screens", "complex64_auto"):
        report["runs"][tag]["rel_l2_vs_complex128"] = float(np.linalg.norm(fields[tag].astype(np.complex128) - ref) / nrm)
    print(json.dumps(report, indent=1))
    with open(os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles", "r1_c5_study.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
