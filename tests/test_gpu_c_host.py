"""The drop-in boundary exercised from a host that is not Python: examples/c_host.c (plain C99, include/pyatm_b200.h only,
no CUDA headers) is compiled with gcc, fed one config-3 batch -- the reference's own seeds from tests/golden/c3_2048.npz --
and its records are compared with the float64 oracle's (oracle.moments, rtol 1e-5) and with the same call made through
ctypes by the Python host (same library, same inputs: equal to the last bits of the float64 accumulation)."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden
from oracle import splitstep as orc
from test_gpu_parity import build_channel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_host_runs_a_config3_batch(tmp_path):
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box")
    libdir = os.path.dirname(os.path.abspath(nat.LIB_PATH))
    exe = str(tmp_path / "c_host")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L", libdir, "-lpyatm_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)

    g = load_golden("c3_2048")
    p = g["params"]
    saved = dict(pa.gpu.config)
    try:
        pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="auto", theta_cut=None)
        ch = build_channel(pa, p)
        ch.path.init_phase_screens()
        ctx = eng.channel_context(ch)
        desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
        c = desc.c
        B, S, M = len(g["seeds"]), p["count"], p["m"]
        hfx = np.empty((S, B, M), dtype=np.float32)
        hfy = np.empty((S, B, M), dtype=np.float32)
        hcf = np.empty((S, B, M), dtype=np.complex64)
        for i in range(B):
            for s in range(S):
                fx, fy = orc.spectrum_to_fxy(g["rho"][i, s], g["theta"][i, s])
                hfx[s, i], hfy[s, i], hcf[s, i] = fx.ravel(), fy.ravel(), g["value"][i, s]
        pup = np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32)
        x, y = ctx.x, ctx.y                      # the float32 axes the Python host hands to pa_ctx_set_axes
        legs = np.ctypeslib.as_array(c.leg_lengths_host, shape=(S + 1,)).copy()
        scales = np.ctypeslib.as_array(c.screen_scale_host, shape=(S,)).copy()

        with open(tmp_path / "batch.bin", "wb") as f:
            np.array([p["n"], S, M, B, c.m_split, c.degree, c.screen_method, 1], dtype=np.int32).tofile(f)
            np.array([p["delta"], c.wvl, c.w0, c.F0, c.final_scale, c.coef_bound], dtype=np.float64).tofile(f)
            legs.tofile(f)
            scales.tofile(f)
            x.tofile(f)
            y.tofile(f)
            hfx.tofile(f)
            hfy.tofile(f)
            hcf.view(np.float32).tofile(f)
            pup.tofile(f)
        run = subprocess.run([exe, str(tmp_path / "batch.bin"), str(tmp_path / "records.bin")], capture_output=True, text=True,
                             timeout=300)
        sys.stdout.write(run.stdout)
        assert run.returncode == 0, run.stderr
        stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
        got = np.fromfile(tmp_path / "records.bin", dtype=np.float64).reshape(B, stride)

        # the same batch through ctypes from this process
        out = np.zeros((B, stride), dtype=np.float64)
        cfv = hcf.view(np.float32)
        nat.check(ctx.lib.pa_simulate_batch(ctx.handle, desc.ref(), B, nat.ptr(hfx), nat.ptr(hfy), nat.ptr(cfv), 0, 0, None,
                                            None, nat.ptr(pup), 1, nat.ptr(out), stride, nat.stream_ptr()))
        np.testing.assert_allclose(got, out, rtol=1e-12, atol=1e-18)
    finally:
        pa.gpu.config.clear()
        pa.gpu.config.update(saved)

    names = [str(k) for k in g["f64_names"]]
    for i in range(B):
        f64 = dict(zip(names, g["f64_measures"][i]))
        assert got[i, nat.MEASURE_HEAD] == pytest.approx(f64["eta_pupil"], rel=1e-5)
        assert got[i, 0] == pytest.approx(f64["eta"], rel=1e-5)
        assert got[i, 3] == pytest.approx(f64["mean_x2"], rel=1e-5)
        assert got[i, 5] == pytest.approx(f64["mean_y2"], rel=1e-5)
