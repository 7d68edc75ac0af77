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


TURB_SCRIPT = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import pyatmosphere_b200 as pa
pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact")
n = 8192
ch = pa.Channel(
    grid=pa.RectGrid(n, 0.75e-3), source=pa.GaussianSource(wvl=808e-9, w0=0.12, F0=np.inf),
    path=pa.IdenticalPhaseScreensPath(
        phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=5e-16, l0=6e-3, L0=1e3),
                                      f_grid=pa.RandLogPolarGrid(points=2**7, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)),
        length=30e3, count=3),
    pupil=pa.CirclePupil(radius=0.2))
np.random.seed(11)
out = ch.run(pupil=False).get()
np.save(sys.argv[1], out[::8, ::8].copy())
crop = out[4096 - 64:4096 + 64, 4096 - 64:4096 + 64]
np.save(sys.argv[1] + ".crop.npy", crop.copy())
print("OK", float(np.sum(np.abs(out.astype(np.complex128)) ** 2) * 0.75e-3 ** 2))
"""


def test_turbulent_path_at_8192_agrees_between_the_two_row_kernels(tmp_path):
    '''8192^2, three screens: the TMA-fed row pass (two-slot ring, screen rows staged through shared memory -- the default
    there) and the direct-access row pass must produce the same field from the same numpy draws (the screens multiply the
    field inside the row pass, so a screen row applied to the wrong field row would show here; unitarity alone would not).'''
    import numpy as np
    outs = {}
    for name, env in (("tma", {"PYATM_FFT_ROWS_TMA": "1"}), ("direct", {"PYATM_FFT_ROWS_TMA": "0"})):
        path = str(tmp_path / (name + ".npy"))
        r = subprocess.run([sys.executable, "-c", TURB_SCRIPT % ROOT, path], env=dict(os.environ, **env), capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
        outs[name] = (np.load(path).astype(np.complex128), np.load(path + ".crop.npy").astype(np.complex128))
    for a, b in zip(outs["tma"], outs["direct"]):
        err = np.linalg.norm(a - b) / np.linalg.norm(b)
        assert err < 2e-6, err
    # and the screens did something: the turbulent field is not the vacuum beam
    assert np.abs(outs["tma"][1]).std() / np.abs(outs["tma"][1]).mean() > 0.05
