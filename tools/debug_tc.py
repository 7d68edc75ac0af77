"""GPU debug helper: tensor-core screen synthesis vs the float64 CUDA-core path (and timing)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng, _native as nat


def channel(n, delta):
    return pa.Channel(
        grid=pa.RectGrid(resolution=n, delta=delta), source=pa.GaussianSource(wvl=808e-9, w0=0.12, F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=5e-16, l0=6e-3, L0=1e3),
                                          f_grid=pa.RandLogPolarGrid(points=2**10, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)),
            length=50e3, count=5),
        pupil=pa.CirclePupil(radius=0.2))


def run(n, theta_cut, nscreens=1):
    pa.gpu.config.update(dtype="complex64", theta_cut=theta_cut)
    ch = channel(n, 1.5e-3 * 2048 / n)
    ch.path.init_phase_screens()
    ps = ch.path.phase_screens[0]
    ctx = eng.channel_context(ch)
    np.random.seed(0)
    fx, fy, cf = eng.draw_spectra_numpy(ch.path, nscreens)
    m = fx.shape[-1]
    dev = ctx.tdevice
    fx_d = torch.as_tensor(fx[:, 0].copy(), device=dev)
    fy_d = torch.as_tensor(fy[:, 0].copy(), device=dev)
    cf_d = torch.as_tensor(cf[:, 0].copy().view(np.float32), device=dev)
    m_split, degree = ps.low_ring_plan()
    bound = eng.coef_bound(ps._get_psd(), m_split)
    out = {}
    for method in (() if os.environ.get('NOPHI') else (0, 1)):
        phi = torch.zeros((nscreens, n, n), dtype=torch.float64, device=dev)
        turns = torch.zeros((nscreens, n, n), dtype=torch.float32, device=dev)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nat.check(ctx.lib.pa_screen_ss(ctx.handle, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), m, m_split, degree, 0.0, 0.0,
                                           nscreens, nat.ptr(turns), nat.ptr(phi), 1, method, bound, nat.stream_ptr()))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[method] = (phi.cpu().numpy(), turns.cpu().numpy(), dt)
    if os.environ.get('NOPHI'):
        turns = torch.zeros((nscreens, n, n), dtype=torch.float32, device=dev)
    # production configuration: turns only (no full-phase output), CUDA-event timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ev0.record()
    for _ in range(reps):
        nat.check(ctx.lib.pa_screen_ss(ctx.handle, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), m, m_split, degree, 0.0, 0.0,
                                       nscreens, nat.ptr(turns), None, 0, 1, bound, nat.stream_ptr()))
    ev1.record()
    torch.cuda.synchronize()
    print(f"   tc turns-only: {ev0.elapsed_time(ev1) / reps * 1e3 / nscreens:.1f} us per screen (all kernels)")
    if os.environ.get('NOPHI'):
        return
    e = out[1][0] - out[0][0]
    k = np.unravel_index(np.argmax(np.abs(e)), e.shape)
    print('   argmax err at (screen,row,col) =', k, 'row slice around:', np.array2string(e[k[0], k[1], max(0, k[2] - 10):k[2] + 10], precision=2))
    colmax = np.abs(e).max(axis=(0, 1)); print('   cols with err > 5e-5:', np.nonzero(colmax > 5e-5)[0][:40])
    rowmax = np.abs(e).max(axis=(0, 2)); print('   rows with err > 5e-5:', np.nonzero(rowmax > 5e-5)[0][:40])
    et = np.abs(np.exp(-2j * np.pi * out[1][1].astype(np.float64)) - np.exp(-2j * np.pi * out[0][1].astype(np.float64)))
    print(f"n={n} theta_cut={theta_cut} m_split={m_split} degree={degree} bound={bound:.2f} nscreens={nscreens}: "
          f"rms(phi_exact)={np.sqrt(np.mean(out[0][0]**2)):.3f} err rms={np.sqrt(np.mean(e**2)):.3e} max={np.abs(e).max():.3e} "
          f"bias={e.mean():.2e} | exp err max={et.max():.2e} | t_exact={out[0][2]*1e3:.2f} ms t_tc={out[1][2]*1e3:.3f} ms", flush=True)


if __name__ == "__main__":
    print("swap =", os.environ.get("PYATM_TC_SWAP", "0"))
    run(256, 2.0)
    run(2048, 2.0)
    run(2048, 6.0)
    run(2048, 6.0, nscreens=8)
