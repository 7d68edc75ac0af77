"""Parity AT THE BENCHMARKED CONFIGURATION (config 3: 2048^2, 5 sparse-spectrum screens of 2^10 rings, 50 km, complex64)
against the float64 oracle run on the same box and against tests/golden/c3_2048.npz, which holds what the unmodified
reference produced for the same seeds (oracle/make_golden.py case_c3; pinned on CPU by
tests/test_oracle_golden.py::test_oracle_at_the_benchmarked_configuration_equals_reference).

Everything bench.py times is covered here with no self-referential hop:
  * Channel.run with the production screen method ('auto' -> tcgen05 split-fp16 contraction + float64 low-ring polynomial)
    and with the float64 CUDA-core screens, and the all-float64 path, each against oracle.propagate(mode='f64');
  * the fused Monte-Carlo route pa_simulate_batch (analytic first leg, tensor-core screens, reductions fused into the final
    row pass, no field written) against oracle.moments(mode='f64') for all seven scalars and the aperture transmittance.

Stated tolerances: complex64 field vs float64 oracle relative L2 <= 1e-5 (north_star), complex128 <= 1e-10; records
rtol 1e-5 (absolute floor: 1e-5 of the record's natural scale, the beam's second moment); vs the reference's own complex64
output 5e-3 (its screens sit ~3e-4 rad from their float64 evaluation, measured 5e-4..6e-4 on the field at this size).
"""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import splitstep as orc
from test_gpu_parity import build_channel

pytestmark = pytest.mark.gpu

KEYS = ("eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2", "mean_x2_r")


@pytest.fixture(autouse=True)
def _cfg():
    import pyatmosphere_b200 as pa
    saved = dict(pa.gpu.config)
    yield
    pa.gpu.config.clear()
    pa.gpu.config.update(saved)


@pytest.fixture(scope="module")
def c3():
    """Fixture + the float64 oracle fields of both seeds, computed once per module (about 5 s per seed)."""
    g = load_golden("c3_2048")
    p = g["params"]
    x, y = orc.rect_xy(p["n"], p["delta"])
    pos = orc.screen_positions(p["length"], p["count"])
    fields = []
    for i in range(len(g["seeds"])):
        screens = []
        for s in range(p["count"]):
            fx, fy = orc.spectrum_to_fxy(g["rho"][i, s], g["theta"][i, s])
            screens.append(orc.ss_screen(x, y, fx, fy, g["value"][i, s], mode="f64"))
        fields.append(orc.propagate(orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64"), screens, p["length"], pos,
                                    p["wvl"], p["delta"], mode="f64"))
        del screens
    g["oracle_fields"] = fields
    g["xy"] = (x, y)
    return g


def test_oracle_on_this_box_reproduces_the_committed_float64_records(c3):
    """The float64 oracle run here equals the one run next to the reference when the fixture was made."""
    p, (x, y) = c3["params"], c3["xy"]
    c0, c1 = (int(v) for v in c3["crop"])
    for i, want in enumerate(c3["oracle_fields"]):
        assert rel_l2(want[c0:c1, c0:c1], c3["f64_crop"][i]) < 1e-9
        m = orc.moments(want, x, y, p["delta"], pupils=[(p["pupil"], (0, 0))], mode="f64")
        assert np.allclose([m[k] for k in KEYS] + [m["eta_pupil"][0]], c3["f64_measures"][i], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("mode", ["auto", "exact", "complex128"])
def test_channel_run_at_config3_vs_float64_oracle(c3, mode):
    """Channel.run(pupil=False) on the reference's seeds: the draws equal the exported ones (same numpy stream), the output
    field matches oracle.propagate(mode='f64') to the north-star tolerance, and the reference's own crop to its floor."""
    import pyatmosphere_b200 as pa
    dtype = "complex128" if mode == "complex128" else "complex64"
    pa.gpu.config.update(use_gpu=True, dtype=dtype, screen_method="auto" if mode != "exact" else "exact",
                         theta_cut=None if mode != "exact" else 2.0)
    from pyatmosphere_b200 import _engine as eng, _native as nat
    p = c3["params"]
    ch = build_channel(pa, p)
    if mode == "auto":
        assert eng.screen_method(p["n"]) == nat.PA_SCREEN_TC          # the tensor-core contraction is what runs here
    c0, c1 = (int(v) for v in c3["crop"])
    tol = 1e-10 if dtype == "complex128" else 1e-5
    for i, seed in enumerate(c3["seeds"]):
        np.random.seed(int(seed))
        out = ch.run(pupil=False).get()
        err = rel_l2(out, c3["oracle_fields"][i])
        print(f"config 3 {mode} seed {int(seed)}: field rel-L2 vs float64 oracle {err:.3e}")
        assert err < tol, (mode, int(seed), err)
        assert rel_l2(out[c0:c1, c0:c1], c3["f64_crop"][i]) < 2 * tol
        assert rel_l2(out[c0:c1, c0:c1], c3["ref_crop"][i]) < 5e-3
        if mode != "auto":
            break                                                      # one seed suffices for the non-default modes


def test_simulate_batch_at_config3_vs_float64_oracle(c3):
    """The route bench.py times: pa_simulate_batch with host coefficient buffers (both seeds as one batch of 2) ->
    analytic first leg, tensor-core screens, fused reductions; all seven scalars and eta_pupil vs oracle.moments(f64)."""
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="auto", theta_cut=None)
    p, (x, y) = c3["params"], c3["xy"]
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ctx = eng.channel_context(ch)
    desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
    assert desc.c.screen_method == nat.PA_SCREEN_TC
    B, S, M = len(c3["seeds"]), p["count"], p["m"]
    hfx = np.empty((S, B, M), dtype=np.float32)
    hfy = np.empty((S, B, M), dtype=np.float32)
    hcf = np.empty((S, B, M), dtype=np.complex64)
    for i in range(B):
        for s in range(S):
            fx, fy = orc.spectrum_to_fxy(c3["rho"][i, s], c3["theta"][i, s])
            hfx[s, i], hfy[s, i], hcf[s, i] = fx.ravel(), fy.ravel(), c3["value"][i, s]
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    pup = np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32)
    out = np.zeros((B, stride), dtype=np.float64)
    cfv = hcf.view(np.float32)
    nat.launch_count(reset=True)
    nat.check(ctx.lib.pa_simulate_batch(ctx.handle, desc.ref(), B, nat.ptr(hfx), nat.ptr(hfy), nat.ptr(cfv), 0, 0, None, None,
                                        nat.ptr(pup), 1, nat.ptr(out), stride, nat.stream_ptr()))
    assert nat.launch_count() > 0
    scale = float(c3["f64_measures"][0][3])                 # <x^2>: natural magnitude of the moment records
    for i in range(B):
        want = orc.moments(c3["oracle_fields"][i], x, y, p["delta"], pupils=[(p["pupil"], (0, 0))], mode="f64")
        for j, k in enumerate(KEYS):
            floor = 1e-5 * (1.0 if k == "eta" else scale if "2" in k or "xy" in k else np.sqrt(scale))
            assert out[i, j] == pytest.approx(want[k], rel=1e-5, abs=floor), (i, k, out[i, j], want[k])
        assert out[i, nat.MEASURE_HEAD] == pytest.approx(want["eta_pupil"][0], rel=1e-5), (i, out[i, nat.MEASURE_HEAD])
        # and the reference's own complex64 records (its floor: 6e-4 on the field)
        assert out[i, nat.MEASURE_HEAD] == pytest.approx(float(c3["ref_measures"][i][-1]), rel=3e-3)
        assert out[i, 3] == pytest.approx(float(c3["ref_measures"][i][3]), rel=3e-3)


def test_simulation_run_numpy_rng_at_config3_reproduces_reference_records(c3):
    """The user-facing call, Simulation([BeamResult, PDTResult]).run(), with numpy-drawn spectra on the reference's seed:
    first record == the reference's record for that seed (its complex64 floor) and == the float64 oracle (1e-5)."""
    import pyatmosphere_b200 as pa
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="auto", theta_cut=None, rng="numpy", batch=2)
    p = c3["params"]
    ch = build_channel(pa, p)
    beam = pa.simulations.BeamResult(ch, max_size=2)
    pdt = pa.simulations.PDTResult(ch, max_size=2)
    np.random.seed(int(c3["seeds"][0]))
    pa.simulations.Simulation([beam, pdt]).run()
    f64 = dict(zip([str(k) for k in c3["f64_names"]], c3["f64_measures"][0]))
    assert pdt.measures[0].data[0] == pytest.approx(f64["eta_pupil"], rel=1e-5)
    assert beam.measures[2].data[0] == pytest.approx(f64["mean_x2"], rel=1e-5)
    assert beam.measures[0].data[0] == pytest.approx(f64["mean_x"], rel=1e-4, abs=1e-7)
    assert pdt.measures[0].data[0] == pytest.approx(float(c3["ref_measures"][0][-1]), rel=3e-3)
