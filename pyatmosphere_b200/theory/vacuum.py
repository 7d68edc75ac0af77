"""Free-space propagation of a sampled field (API of /root/reference/pyatmosphere/theory/vacuum.py:5-7).

The reference builds the leg from its centred fft2/ifft2 helpers; here it is the fused pass pair of the native library
(pa_vacuum_leg: FFT_x, FFT_y * H * IFFT_y, IFFT_x), which is what pathes.VacuumPath calls as well."""
from __future__ import annotations

import math

import numpy as np


def vacuum_propagation(input, length, k, delta, f2=None, f_delta=None):
    """IFFT2c( exp(ikL) exp(-i pi L lambda f^2) FFT2c(input) ) on an N x N grid of spacing `delta`.

    `f2` and `f_delta` (the squared frequency grid and its step, theory/vacuum.py:5) are implied by N and delta; they are
    accepted for signature compatibility and `f_delta` is checked against 1 / (N delta)."""
    from .. import _engine as eng
    from .. import _native as nat
    from .. import gpu
    from ..grids import RectGrid
    gpu.require_gpu()
    torch = nat.torch_mod()
    t = input.t if isinstance(input, gpu.DeviceArray) else torch.as_tensor(np.asarray(input), device="cuda")
    if t.ndim != 2 or t.shape[0] != t.shape[1]:
        raise ValueError("vacuum_propagation takes a square [N][N] field")
    n = int(t.shape[0])
    if f_delta is not None and not math.isclose(float(f_delta), 1.0 / (n * float(delta)), rel_tol=1e-6):
        raise ValueError("f_delta must be 1 / (N * delta) (the frequency grid of grids.RectGrid.get_f_grid)")
    ctx = eng.grid_context(RectGrid(n, float(delta)))
    field = t.to(ctx.cdtype).reshape(1, n, n).clone().contiguous()
    nat.check(ctx.lib.pa_vacuum_leg(ctx.handle, nat.ptr(field), 1, float(length), 2 * math.pi / float(k), nat.stream_ptr()))
    return gpu.DeviceArray(field[0])
