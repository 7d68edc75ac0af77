"""Both implementations of the FFT passes (TMA-fed persistent kernels and direct-access kernels) must agree with
the oracle; the variant is chosen per context from the environment, so each runs in its own process."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
import pyatmosphere_b200 as pa
from pyatmosphere_b200.gpu import DeviceArray
from oracle import splitstep as orc
for dtype, tol in (("complex64", 2e-6), ("complex128", 1e-12)):
    pa.gpu.config.update(dtype=dtype)
    for n in (512, 2048):
        rng = np.random.default_rng(n)
        u = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(dtype)
        ch = pa.Channel(grid=pa.RectGrid(n, 2e-3), source=pa.GaussianSource(wvl=808e-9, w0=0.05, F0=np.inf),
                        path=pa.VacuumPath(length=1.5e3), pupil=pa.CirclePupil(radius=1.0))
        out = ch.path.output(DeviceArray(torch.as_tensor(u).cuda())).get()
        want = orc.vacuum_leg(u, 1.5e3, 808e-9, 2e-3, mode="f64")
        err = np.linalg.norm(out - want) / np.linalg.norm(want)
        assert err < tol, (dtype, n, err)
print("OK")
"""


@pytest.mark.parametrize("env", [{"PYATM_FFT_DIRECT": "1"}, {"PYATM_FFT_DIRECT": "0", "PYATM_FFT_ROWS_TMA": "1"},
                                 {"PYATM_FFT_DIRECT": "0", "PYATM_FFT_ROWS_TMA": "0"}])
def test_vacuum_leg_with_each_fft_variant(env):
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
