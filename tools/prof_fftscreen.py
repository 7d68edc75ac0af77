"""Timing of pa_screen_fft (FFT phase screens, csrc/screen_fft.cu) with CUDA events (profiling helper).

    python tools/prof_fftscreen.py [n] [batch] [subharmonic terms]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    nterms = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    ctx = eng.grid_context(pa.RectGrid(n, 1.5e-3))
    spec = torch.randn((batch, n, n), dtype=torch.complex64, device="cuda")
    out = torch.empty((batch, n, n), dtype=torch.float32, device="cuda")
    # the reference's pattern: 3 x 3 patches (centre dropped) at spacing df / 3^(level+1) -> 3 distinct fx per level
    rng = np.random.default_rng(0)
    df = 1 / (n * 1.5e-3)
    rows = []
    for level in range((nterms + 7) // 8):
        d = df / 3 ** (level + 1)
        rows += [(a * d, b * d, *rng.standard_normal(2)) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
    terms = np.ascontiguousarray(np.tile(np.array(rows[:nterms], dtype=np.float64), (batch, 1, 1)))
    stream = nat.stream_ptr()

    def call():
        nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(spec), batch, nat.ptr(terms), nterms, None, nat.ptr(out), stream))

    for _ in range(3):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 20
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    traffic = batch * n * n * 8 * 8          # gather r+w, columns r+w, rows r+w, add r+w (complex64) -- minimum of this scheme
    print(f"pa_screen_fft n={n} batch={batch} terms={nterms}: {us:.1f} us per call, {us / batch:.1f} us per screen, "
          f"{traffic / us / 1e3:.0f} GB/s over the scheme's minimum traffic")


if __name__ == "__main__":
    main()
