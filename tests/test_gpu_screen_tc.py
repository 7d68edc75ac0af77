"""Tensor-core (tcgen05, split-fp16) screen synthesis: accuracy against the float64 oracle / the float64 CUDA-core
path, and end-to-end field parity with it switched on (the default for complex64 grids that are multiples of 256).

Stated tolerances for the tensor-core method (README advanced channel, theta_cut = 10):
  phase error vs float64: rms <= 5e-6 rad, max <= 6e-5 rad per screen;
  complex64 field after 5 screens vs float64 oracle / complex128 path: relative L2 <= 1e-5 (the north-star tolerance;
  the direct comparison with the oracle at config 3 is tests/test_gpu_c3_parity.py).
"""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import splitstep as orc
from test_gpu_parity import build_channel, oracle_field

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cfg():
    import pyatmosphere_b200 as pa
    saved = dict(pa.gpu.config)
    yield
    pa.gpu.config.clear()
    pa.gpu.config.update(saved)


def test_auto_method_selection():
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    pa.gpu.config.update(dtype="complex64", screen_method="auto", theta_cut=None)
    assert eng.screen_method(2048) == nat.PA_SCREEN_TC and eng.screen_method(128) == nat.PA_SCREEN_EXACT
    assert eng.screen_method(1024) == nat.PA_SCREEN_TC and eng.screen_method(512) == nat.PA_SCREEN_EXACT
    assert eng.theta_cut(2048) == 10.0 and eng.theta_cut(128) == 2.0 and eng.theta_cut(1024) == pytest.approx(5.12)
    pa.gpu.config.update(dtype="complex128")
    assert eng.screen_method(2048) == nat.PA_SCREEN_EXACT
    pa.gpu.config.update(screen_method="tc")
    with pytest.raises(ValueError):
        eng.screen_method(2048)


@pytest.mark.parametrize("theta_cut", [0.5, 1.0, 10.0])
def test_tc_screen_vs_oracle_256(theta_cut):
    """quick256 fixture (M = 1024 rings, QuickChannel spectrum), tensor-core method forced on a 256^2 grid (the host
    clamps theta_cut to 1.28 there): phase vs float64 evaluation, tolerance relative to what goes through fp32."""
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200.utils import PolarDiscreteFunction
    pa.gpu.config.update(dtype="complex64", screen_method="tc", theta_cut=theta_cut)
    g = load_golden("quick256")
    p = g["params"]
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ps = ch.path.phase_screens[0]
    x, y = orc.rect_xy(p["n"], p["delta"])
    sp = PolarDiscreteFunction(g["rho"][0], g["theta"][0], g["value"][0])
    fx, fy = orc.spectrum_to_fxy(g["rho"][0], g["theta"][0])
    want = orc.ss_screen(x, y, fx, fy, g["value"][0], mode="f64")
    turns, phi = ps._synthesize(sp, (0, 0), want_turns=True, want_phi=True)
    m_split, _ = ps.low_ring_plan()
    hi = orc.ss_screen(x, y, fx[:, m_split:], fy[m_split:], g["value"][0][m_split:], mode="f64")
    err = np.exp(-2j * np.pi * turns.cpu().numpy().astype(np.float64)) - np.exp(-1j * want)
    # error scales with the magnitude that goes through the fp32 accumulator (rms of the high-ring part)
    scale = max(1.0, float(np.sqrt(np.mean(hi**2))))
    assert np.sqrt(np.mean(np.abs(err) ** 2)) < 3e-6 * scale
    assert np.max(np.abs(err)) < 3e-5 * scale
    assert np.max(np.abs(phi.cpu().numpy() - want)) < 1.5e-7 * np.max(np.abs(want)) + 3e-5 * scale


def test_tc_matches_exact_path_full_size():
    """2048^2 README channel: tensor-core vs float64 CUDA-core synthesis of the same 3 screens."""
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    from bench import C3
    pa.gpu.config.update(dtype="complex64", screen_method="auto", theta_cut=None)
    ch = build_channel(pa, C3)
    ch.path.init_phase_screens()
    ps = ch.path.phase_screens[0]
    ctx = eng.channel_context(ch)
    np.random.seed(0)
    fx, fy, cf = eng.draw_spectra_numpy(ch.path, 3)
    dev = ctx.tdevice
    fx_d, fy_d = torch.as_tensor(fx[:, 0].copy(), device=dev), torch.as_tensor(fy[:, 0].copy(), device=dev)
    cf_d = torch.as_tensor(cf[:, 0].copy().view(np.float32), device=dev)
    n, m = ctx.n, fx.shape[-1]
    out = {}
    for method, tc in ((nat.PA_SCREEN_EXACT, 2.0), (nat.PA_SCREEN_TC, 10.0)):
        pa.gpu.config.update(theta_cut=tc)
        m_split, degree = ps.low_ring_plan()
        phi = torch.zeros((3, n, n), dtype=torch.float64, device=dev)
        nat.check(ctx.lib.pa_screen_ss(ctx.handle, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), m, m_split, degree, 0.0, 0.0, 3,
                                       None, nat.ptr(phi), 1, method, eng.coef_bound(ps._get_psd(), m_split), nat.stream_ptr()))
        out[method] = phi.cpu().numpy()
    e = out[nat.PA_SCREEN_TC] - out[nat.PA_SCREEN_EXACT]
    assert np.sqrt(np.mean(e**2)) < 5e-6 and np.max(np.abs(e)) < 6e-5


@pytest.mark.parametrize("name", ["quick256"])
def test_channel_run_with_tc_screens(name):
    """The fixture's physical channel on a 1024^2 grid of the same extent (the tensor-core path starts at 1024):
    same seed -> same coefficients; checked against the float64 oracle evaluated on that grid."""
    import pyatmosphere_b200 as pa
    pa.gpu.config.update(dtype="complex64", screen_method="auto", theta_cut=None)
    g = load_golden(name)
    g = dict(g, params=dict(g["params"], n=1024, delta=g["params"]["delta"] / 4))
    ch = build_channel(pa, g["params"])
    np.random.seed(int(g["seed"]))
    out = ch.run(pupil=False).get()
    want, _ = oracle_field(g, "f64")
    err = rel_l2(out, want)
    print(f"quick channel on 1024^2, tensor-core screens vs float64 oracle: rel-L2 {err:.3e}")
    assert err < 1e-5


def test_full_size_tc_vs_complex128():
    """Config 3 end to end: complex64 + tensor-core screens vs the all-float64 path on the same seed."""
    import pyatmosphere_b200 as pa
    from bench import C3
    fields = {}
    for dtype in ("complex64", "complex128"):
        pa.gpu.config.update(dtype=dtype, screen_method="auto", theta_cut=None)
        ch = build_channel(pa, C3)
        np.random.seed(3)
        out = ch.run(pupil=False)
        assert pa.measures.eta(ch, output=out) == pytest.approx(1.0, abs=2e-5)
        fields[dtype] = out.get()
    err = rel_l2(fields["complex64"], fields["complex128"])
    print(f"config 3, tensor-core complex64 vs complex128: rel-L2 {err:.3e}")
    assert err < 1e-5


def test_single_cta_kernel_matches_oracle_in_a_subprocess():
    """The contraction runs as CTA pairs (cta_group::2) by default; the single-CTA variant (PYATM_TC_PAIR=0, read once per
    process) must stay correct too: the 256^2 oracle comparison and the full-size comparison with the exact path, in a
    child interpreter."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PYATM_TC_PAIR="0")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(here, "test_gpu_screen_tc.py"),
                        "-k", "tc_screen_vs_oracle_256 or tc_matches_exact_path_full_size"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
