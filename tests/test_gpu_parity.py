"""Parity of the CUDA path (through the C ABI, via the Python host mirror) against the CPU oracle and the golden
fixtures produced by the reference.  Run on the B200 box: python -m pytest tests -m gpu.

Stated tolerances (relative L2 of the complex field unless noted):
  complex64  path vs float64 oracle : 1e-5       complex128 path vs float64 oracle : 1e-10
  (complex64 with the tensor-core screen synthesis: see tests/test_gpu_screen_tc.py)
  any path vs the reference's own complex64 output: 5e-3 on turbulent fields -- the reference's complex64
  screens are themselves ~3e-4 rad away from their float64 evaluation, a floor we do not copy.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import splitstep as orc

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-5, "complex128": 1e-10}
TURB = ["turb128", "turb128_after_lossy", "turb64_before", "quick256"]


@pytest.fixture(autouse=True)
def _cfg():
    import pyatmosphere_b200 as pa
    saved = dict(pa.gpu.config)
    yield
    pa.gpu.config.clear()
    pa.gpu.config.update(saved)


def _pa(dtype="complex64", **kw):
    """Tests in this file pin the float64 CUDA-core screen path unless they ask for the tensor-core one."""
    import pyatmosphere_b200 as pa
    kw.setdefault("screen_method", "exact")
    kw.setdefault("theta_cut", 2.0)
    pa.gpu.config.update(use_gpu=True, dtype=dtype, **kw)
    return pa


def build_channel(pa, p):
    return pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=p.get("F0", np.inf)),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(
                model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"], position_in_slab=p.get("where", "middle"),
            losses_db=p.get("losses_db", 0)),
        pupil=pa.CirclePupil(radius=p["pupil"]))


def oracle_field(g, mode="f64", through_output=True):
    p = g["params"]
    x, y = orc.rect_xy(p["n"], p["delta"])
    u0 = orc.gaussian_source(x, y, p["w0"], p["wvl"], p.get("F0", np.inf), mode=mode)
    screens = []
    for s in range(p["count"]):
        fx, fy = orc.spectrum_to_fxy(g["rho"][s], g["theta"][s])
        screens.append(orc.ss_screen(x, y, fx, fy, g["value"][s], mode=mode))
    pos = orc.screen_positions(p["length"], p["count"], p.get("where", "middle"))
    return orc.propagate(u0, screens, p["length"], pos, p["wvl"], p["delta"], mode=mode,
                         losses_db=p.get("losses_db", 0), through_output=through_output), screens


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
def test_vacuum_leg_random_field_vs_oracle(dtype, n):
    """pa_vacuum_leg on a random complex field == oracle leg (float64): exercises every FFT size, the permuted
    spectrum order and the separable transfer function."""
    pa = _pa(dtype)
    from pyatmosphere_b200.gpu import DeviceArray
    import torch
    rng = np.random.default_rng(n)
    delta, wvl, length = 2e-3, 808e-9, 1.5e3
    u = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(dtype)
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=0.05, F0=np.inf),
                    path=pa.VacuumPath(length=length), pupil=pa.CirclePupil(radius=1.0))
    out = ch.path.output(DeviceArray(torch.as_tensor(u).cuda())).get()
    want = orc.vacuum_leg(u, length, wvl, delta, mode="f64")
    assert rel_l2(out, want) < (2e-6 if dtype == "complex64" else 1e-12)
    assert np.array_equal(ch.path.output(DeviceArray(torch.as_tensor(u).cuda()), length=0).get(), u)   # pathes.py:30


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_vacuum_notebook_asserts(dtype):
    """tests/itest_vacuum_propagation.ipynb cells 3-7 through the CUDA path."""
    pa = _pa(dtype)
    g = load_golden("vacuum256")
    n, delta, wvl, w0, length = int(g["n"]), float(g["delta"]), float(g["wvl"]), float(g["w0"]), float(g["length"])
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=w0, F0=np.inf),
                    path=pa.VacuumPath(length=length), pupil=pa.CirclePupil(radius=1.0))
    src = ch.source.output().get()
    assert rel_l2(src, g["source"]) < (2e-7 if dtype == "complex64" else 2e-7)   # reference source has a c64 exp
    out = ch.run(pupil=False)
    assert rel_l2(out.get(), g["field"]) < 1e-6
    m = pa.measures
    assert m.eta(ch, output=ch.pupil.output(out)) == pytest.approx(1.0, abs=1e-6)
    assert abs(out.get().astype(np.complex128).sum()) * delta**2 == pytest.approx(np.sqrt(2 * np.pi) * w0, abs=1e-7)
    assert m.eta(ch, output=out) == pytest.approx(1.0, abs=1e-5)
    w = np.sqrt(2 * (m.mean_x2(ch, output=out) + m.mean_y2(ch, output=out)))
    assert w == pytest.approx(ch.source.get_w(length), abs=1e-7)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("name", ["turb128", "turb64_before", "quick256"])
@pytest.mark.parametrize("theta_cut", [0.0, 2.0])
def test_screen_vs_oracle(dtype, name, theta_cut):
    """pa_screen_ss == float64 evaluation of the same harmonics; with and without the polynomial split."""
    pa = _pa(dtype, theta_cut=theta_cut)
    g = load_golden(name)
    p = g["params"]
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ps = ch.path.phase_screens[0]
    assert np.array_equal(ps._get_psd(), g["psd"])
    x, y = orc.rect_xy(p["n"], p["delta"])
    from pyatmosphere_b200.utils import PolarDiscreteFunction
    for s in range(min(2, p["count"])):
        sp = PolarDiscreteFunction(g["rho"][s], g["theta"][s], g["value"][s])
        fx, fy = orc.spectrum_to_fxy(g["rho"][s], g["theta"][s])
        want = orc.ss_screen(x, y, fx, fy, g["value"][s], mode="f64", complex_out=True)
        turns, phi = ps._synthesize(sp, (0, 0), want_turns=True, want_phi=True)
        phi, turns = phi.cpu().numpy().astype(np.float64), turns.cpu().numpy().astype(np.float64)
        scale = np.max(np.abs(want.real))
        if dtype == "complex128":
            assert np.max(np.abs(phi - want.real)) < 1e-10 * max(1.0, scale)
            assert np.max(np.abs(np.exp(-2j * np.pi * turns) - np.exp(-1j * want.real))) < 1e-9
        else:
            assert np.max(np.abs(phi - want.real)) < 1.5e-7 * scale + 1e-6       # float32 storage of the full phase
            assert np.max(np.abs(np.exp(-2j * np.pi * turns) - np.exp(-1j * want.real))) < 1e-6
        _, im = ps._synthesize(sp, (0, 0), want_turns=False, want_phi=True, imag_part=True)
        assert np.max(np.abs(im.cpu().numpy() - want.imag)) < 1.5e-7 * np.max(np.abs(want.imag)) + 1e-6
    if "screens" in g:
        # the reference's own complex64 screen: within its documented error of our float64-exact one
        sp = PolarDiscreteFunction(g["rho"][0], g["theta"][0], g["value"][0])
        _, phi = ps._synthesize(sp, (0, 0), want_turns=False, want_phi=True)
        assert np.max(np.abs(phi.cpu().numpy() - g["screens"][0])) < 5e-3


def test_screen_shift_and_wind():
    pa = _pa("complex128")
    g = load_golden("turb128")
    p = g["params"]
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ps = ch.path.phase_screens[0]
    x, y = orc.rect_xy(p["n"], p["delta"])
    np.random.seed(3)
    a = ps.generate(shift=(0.05, -0.02), wind=True).get()
    sp = ps._cached_spectrum
    fx, fy = orc.spectrum_to_fxy(sp.rho, sp.theta)
    want = orc.ss_screen(x, y, fx, fy, sp.value, shift=(0.05, -0.02), mode="f64")
    assert np.max(np.abs(a - want)) < 1e-9 * max(1.0, np.max(np.abs(want)))
    b = ps.generate(shift=(0.05, -0.02), wind=True).get()        # cached spectrum -> identical screen
    assert np.array_equal(a, b)
    ps.cache_clear()
    c = ps.generate(shift=(0.05, -0.02), wind=True).get()
    assert not np.array_equal(a, c)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("name", TURB)
def test_channel_run_vs_reference_and_oracle(dtype, name):
    """Channel.run with np.random.seed(seed): same coefficients as the reference run that made the fixture."""
    pa = _pa(dtype)
    g = load_golden(name)
    ch = build_channel(pa, g["params"])
    np.random.seed(int(g["seed"]))
    out = ch.run(pupil=False).get()
    want, _ = oracle_field(g, "f64")
    assert rel_l2(out, want) < TOL[dtype]
    assert rel_l2(out, g["field"]) < 5e-3
    # generator route (per-step building blocks) == fused route, and yields the per-leg fields / screens
    np.random.seed(int(g["seed"]))
    legs = list(ch.generator(pupil=False, store_output=True))
    want_gen, screens64 = oracle_field(g, "f64", through_output=False)
    assert rel_l2(ch.output.get(), want_gen) < TOL[dtype]
    if "legs" in g:
        for (u, phi), ref_u, ref_phi, s64 in zip(legs, g["legs"], g["screens"], screens64):
            assert rel_l2(u.get(), ref_u) < 5e-3
            assert np.max(np.abs(phi.get() - s64)) < 1.5e-7 * np.max(np.abs(s64)) + 1e-6
    # pupil=True applies the channel's aperture (channels.py:31-33)
    np.random.seed(int(g["seed"]))
    masked = ch.run().get()
    x, y = orc.rect_xy(g["params"]["n"], g["params"]["delta"])
    assert np.array_equal(masked != 0, (out != 0) & orc.circle_mask(x, y, g["params"]["pupil"]))


@pytest.mark.parametrize("name", TURB)
def test_measures_vs_oracle(name):
    pa = _pa("complex64")
    import torch
    from pyatmosphere_b200.gpu import DeviceArray
    g = load_golden(name)
    p = g["params"]
    ch = build_channel(pa, p)
    field = g["field"].astype(np.complex64)
    out = DeviceArray(torch.as_tensor(field).cuda())
    m = pa.measures
    x, y = orc.rect_xy(p["n"], p["delta"])
    want = orc.moments(field, x, y, p["delta"], pupils=[(p["pupil"], (0, 0)), (p["pupil"] / 2, (0.01, -0.02))], mode="f64")
    got = [m.eta(ch, output=out), m.mean_x(ch, output=out), m.mean_y(ch, output=out), m.mean_x2(ch, output=out),
           m.mean_xy(ch, output=out), m.mean_y2(ch, output=out)]
    ref = [want[k] for k in ("eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2")]
    assert np.allclose(got, ref, rtol=2e-6, atol=1e-9)
    assert np.allclose(got + [m.eta(ch, output=ch.pupil.output(out))], g["measures"], rtol=2e-5, atol=2e-7)
    allm = m.all_moments(ch, out, [(p["pupil"], (0, 0)), (p["pupil"] / 2, (0.01, -0.02))])
    assert np.allclose(allm["eta_pupil"][0], want["eta_pupil"], rtol=2e-6)
    assert allm["mean_x2_r"][0] == pytest.approx(want["mean_x2_r"], rel=2e-6)
    inten = m.I(ch, output=out).get()
    assert np.allclose(inten, np.abs(field) ** 2, rtol=1e-6)


def test_pupil_edge_membership_is_exact():
    """The float32 predicate of pupils.py:10 is reproduced bit for bit, including shifted apertures."""
    pa = _pa("complex64")
    import torch
    from pyatmosphere_b200.gpu import DeviceArray
    n, delta = 256, 1.5e-3
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=808e-9, w0=0.05, F0=np.inf),
                    path=pa.VacuumPath(length=10.0), pupil=pa.CirclePupil(radius=0.12))
    ones = DeviceArray(torch.ones((n, n), dtype=torch.complex64).cuda())
    x, y = orc.rect_xy(n, delta)
    for r, shift in [(0.12, (0, 0)), (0.0303, (0.0127, -0.0451)), (0.09, (-0.05, 0.0333)), (0.0015, (0, 0))]:
        pup = pa.CirclePupil(radius=r)
        pup.channel = ch
        got = pup.output(ones, shift=shift).get().real != 0
        assert np.array_equal(got, orc.circle_mask(x, y, r, shift))
        assert np.array_equal(got, pup.get_pupil(shift))


def test_histogram_numpy_semantics():
    pa = _pa("complex64")
    import torch
    from pyatmosphere_b200 import _native as nat
    ctx = nat.context(64, 1e-3, orc.rect_axis(64, 1e-3), orc.rect_axis(64, 1e-3), 0)
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.random(5000), np.linspace(0, 1, 201), [1.0, 0.0, -0.2, 1.3, 0.005, 0.9999999999]])
    edges = np.linspace(0, 1, 201)
    vd, ed = torch.as_tensor(v).cuda(), torch.as_tensor(edges).cuda()
    counts = torch.zeros(200, dtype=torch.int64).cuda()
    nat.check(ctx.lib.pa_histogram(ctx.handle, nat.ptr(vd), 1, len(v), nat.ptr(ed), 200, nat.ptr(counts), nat.stream_ptr()))
    assert np.array_equal(counts.cpu().numpy(), np.histogram(v, bins=200, range=(0, 1))[0])


def test_simulation_replay_matches_reference_table():
    """Simulation([BeamResult, PDTResult]).run() with the reference's seed reproduces the reference's records;
    batched route == generic route."""
    g = load_golden("simulation128")
    p = g["params"]
    tables = {}
    for batch in (4, 1):
        pa = _pa("complex64", batch=batch, rng="numpy")
        ch = build_channel(pa, p)
        beam = pa.simulations.BeamResult(ch, max_size=g["table"].shape[0])
        pdt = pa.simulations.PDTResult(ch, max_size=g["table"].shape[0])
        sim = pa.simulations.Simulation([beam, pdt])
        assert sim.batchable()
        np.random.seed(int(g["seed"]))
        sim.run()
        tables[batch] = np.array([m.data for m in beam.measures] + [pdt.measures[0].data]).T
        assert [m.name for m in beam.measures] + [pdt.measures[0].name] == list(g["names"])
        # vs the reference's complex64 run: limited by the reference's own screen error
        assert np.allclose(tables[batch], g["table"], rtol=5e-3, atol=2e-5)
        assert np.allclose([beam.bw, beam.lt, beam.st], g["stats"], rtol=5e-3)
    assert np.allclose(tables[4], tables[1], rtol=1e-6, atol=1e-9)


def test_tracked_pdt_matches_two_phase_oracle():
    pa = _pa("complex64", batch=3, rng="numpy")
    g = load_golden("turb128")
    p = g["params"]
    ch = build_channel(pa, p)
    res = pa.simulations.TrackedPDTResult(ch, max_size=3)
    sim = pa.simulations.Simulation([res])
    np.random.seed(11)
    sim.run()
    # oracle replay
    x, y = orc.rect_xy(p["n"], p["delta"])
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    pos = orc.screen_positions(p["length"], p["count"])
    np.random.seed(11)
    for r in range(3):
        u = orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64")
        scr = []
        for s in range(p["count"]):
            rho, theta, value = orc.draw_spectrum(base, g["psd"])
            fx, fy = orc.spectrum_to_fxy(rho, theta)
            scr.append(orc.ss_screen(x, y, fx, fy, value, mode="f64"))
        out = orc.propagate(u, scr, p["length"], pos, p["wvl"], p["delta"], mode="f64", through_output=False)
        m = orc.moments(out, x, y, p["delta"], mode="f64")
        shift = (np.float32(m["mean_x"]), np.float32(m["mean_y"]))
        eta = orc.moments(out, x, y, p["delta"], pupils=[(p["pupil"], shift)], mode="f64")["eta_pupil"][0]
        assert res.measures[0].data[r] == pytest.approx(m["mean_x"], rel=1e-4, abs=1e-8)
        assert res.measures[2].data[r] == pytest.approx(eta, rel=1e-3)


def test_device_rng_statistics_and_sharding_invariance():
    pa = _pa("complex64", rng="philox", seed=42, batch=4)
    import torch
    from pyatmosphere_b200 import _engine as eng, _native as nat
    p = load_golden("turb128")["params"]
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ctx = eng.channel_context(ch)
    ps = ch.path.phase_screens[0]
    edges_d, psd_d = eng.ring_tables(ctx, ps)
    S, M, B = 3, p["m"], 512

    def draw(first, batch):
        fx = torch.empty((S, batch, M), dtype=torch.float32).cuda()
        fy = torch.empty_like(fx)
        cf = torch.empty((S, batch, M, 2), dtype=torch.float32).cuda()
        nat.check(ctx.lib.pa_rng_spectrum(ctx.handle, 42, first, batch, 0, S, M, nat.ptr(edges_d), nat.ptr(psd_d),
                                          nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), nat.stream_ptr()))
        return fx.cpu().numpy(), fy.cpu().numpy(), cf.cpu().numpy()

    fx, fy, cf = draw(0, B)
    fx2, fy2, cf2 = draw(100, 50)
    assert np.array_equal(fx[:, 100:150], fx2) and np.array_equal(cf[:, 100:150], cf2)      # keyed by global index
    rho = np.hypot(fx, fy)
    base = ps.f_grid.base
    inner = np.insert(base, 0, 0)[:-1]
    assert np.all(rho <= base * (1 + 1e-6)) and np.all(rho >= inner * (1 - 1e-6))
    n0 = cf[..., 0] / np.sqrt(ps._get_psd())
    n1 = cf[..., 1] / np.sqrt(ps._get_psd())
    for v in (n0, n1):
        assert abs(v.mean()) < 5 / np.sqrt(v.size) and abs(v.var() - 1) < 0.02
    theta = np.arctan2(fy, fx)
    assert abs(np.mean(np.cos(theta))) < 0.01 and abs(np.mean(np.sin(theta))) < 0.01
    # Simulation in device-RNG mode: records do not depend on the batch size
    rows = {}
    for batch in (4, 3):
        pa.gpu.config.update(batch=batch)
        beam = pa.simulations.BeamResult(ch, max_size=6)
        sim = pa.simulations.Simulation([beam])
        sim.run()
        rows[batch] = np.array([m.data for m in beam.measures])
    assert np.array_equal(rows[4], rows[3])


def test_no_cpu_path():
    pa = _pa("complex64")
    pa.gpu.config["use_gpu"] = False
    ch = pa.QuickChannel(grid_resolution=64)
    with pytest.raises(pa.gpu.NoCpuPathError):
        ch.run()


# ---- full-size, size-independent properties --------------------------------------------------------------------
def test_full_size_vacuum_matches_analytic_gaussian():
    """Config 2: 2048^2, delta = 1.5 mm, w0 = 0.12, 50 km vacuum vs the closed-form Gaussian beam."""
    pa = _pa("complex64")
    n, delta, wvl, w0, length = 2048, 1.5e-3, 808e-9, 0.12, 50e3
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=w0, F0=np.inf),
                    path=pa.VacuumPath(length=length), pupil=pa.CirclePupil(radius=0.2))
    out = ch.run(pupil=False)
    x, y = orc.rect_xy(n, delta)
    assert rel_l2(out.get(), orc.analytic_gaussian_field(x, y, w0, wvl, length)) < 2e-6
    m = pa.measures
    w = np.sqrt(2 * (m.mean_x2(ch, output=out) + m.mean_y2(ch, output=out)))
    assert w == pytest.approx(ch.source.get_w(length), abs=2e-7)
    assert m.eta(ch, output=out) == pytest.approx(1.0, abs=1e-5)


def test_full_size_turbulent_energy_and_determinism():
    """Config 3 (README advanced channel): the split-step operator is unitary on the periodic grid, so the total
    power stays 1; the same seed gives the same field; complex64 and complex128 agree to the complex64 tolerance."""
    fields = {}
    for dtype in ("complex64", "complex128"):
        pa = _pa(dtype)
        ch = pa.Channel(
            grid=pa.RectGrid(resolution=2048, delta=0.0015), source=pa.GaussianSource(wvl=808e-9, w0=0.12, F0=np.inf),
            path=pa.IdenticalPhaseScreensPath(
                phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=5e-16, l0=6e-3, L0=1e3),
                                              f_grid=pa.RandLogPolarGrid(points=2**10, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)),
                length=50e3, count=5),
            pupil=pa.CirclePupil(radius=0.2))
        assert ch.get_rythov2() == pytest.approx(27.7, abs=0.05)          # main.ipynb:184
        np.random.seed(1)
        out = ch.run(pupil=False)
        assert pa.measures.eta(ch, output=out) == pytest.approx(1.0, abs=2e-5)
        fields[dtype] = out.get()
        if dtype == "complex64":
            np.random.seed(1)
            assert np.array_equal(ch.run(pupil=False).get(), fields[dtype])
    assert rel_l2(fields["complex64"], fields["complex128"]) < 1e-5


@pytest.mark.parametrize("apertures", [[], [0.12], [0.12, 0.05], [0.12, 0.05, 0.2, 0.01], [0.12, 0.05, 0.2, 0.01, 0.3]])
def test_fused_statistics_route_equals_literal_route(monkeypatch, apertures):
    """The Monte-Carlo route (analytic first leg, reductions fused into the final row pass, no field written) gives
    the same records as the literal route (source pass, six full legs, separate measure sweep), for every number of
    apertures the fused pass is compiled for (0..4) and beyond (5: separate sweep again)."""
    from pyatmosphere_b200 import _engine as eng, _native as nat
    g = load_golden("quick256")
    p = dict(g["params"], n=512, delta=2e-3)       # 512: smallest size with the fused final pass
    tables = {}
    for tag, env in (("fused", {}), ("literal", {"PYATM_NO_FUSED_MEASURE": "1", "PYATM_NO_ANALYTIC_LEG": "1"})):
        for k in ("PYATM_NO_FUSED_MEASURE", "PYATM_NO_ANALYTIC_LEG"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        nat.clear_contexts()
        pa = _pa("complex64", rng="philox", seed=7, batch=4)
        ch = build_channel(pa, p)
        cols = eng.table_columns(sorted(apertures), [])
        tables[tag] = eng.simulate_realizations(ch, 0, 4, np.arange(4), sorted(apertures), [])
        assert tables[tag].shape == (4, len(cols))
    nat.clear_contexts()
    assert np.allclose(tables["fused"], tables["literal"], rtol=2e-5, atol=1e-8)
    assert np.all(tables["fused"][:, 0] == pytest.approx(1.0, abs=1e-4))


@pytest.mark.parametrize("n,dtype", [(4096, "complex64"), (8192, "complex64"), (4096, "complex128")])
def test_long_haul_grid_sizes(n, dtype):
    """Config 5 grid sizes (up to 8192^2): vacuum leg vs the analytic Gaussian beam and the energy / width
    invariants, plus one turbulent step (screen + leg) checked for unitarity."""
    pa = _pa(dtype)
    delta, wvl, w0, length = 1.5e-3 * 2048 / n * 2, 808e-9, 0.12, 20e3
    ch = pa.Channel(
        grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=w0, F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=5e-16, l0=6e-3, L0=1e3),
                                          f_grid=pa.RandLogPolarGrid(points=2**8, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)),
            length=length, count=2),
        pupil=pa.CirclePupil(radius=0.2))
    vac = pa.VacuumPath(length=length)
    vac.channel = ch
    out = vac.output(ch.source.output())
    x, y = orc.rect_xy(n, delta)
    ana = orc.analytic_gaussian_field(x, y, w0, wvl, length)
    assert rel_l2(out.get(), ana) < 3e-6          # limited by the discretisation (the reference itself reaches ~1e-7)
    if dtype == "complex128":                     # float64 path: bit-level agreement with the float64 oracle
        want = orc.vacuum_leg(orc.gaussian_source(x, y, w0, wvl, mode="f64"), length, wvl, delta, mode="f64")
        assert rel_l2(out.get(), want) < 1e-10
        del want
    m = pa.measures
    assert m.eta(ch, output=out) == pytest.approx(1.0, abs=1e-5)
    w = np.sqrt(2 * (m.mean_x2(ch, output=out) + m.mean_y2(ch, output=out)))
    assert w == pytest.approx(ch.source.get_w(length), abs=5e-7)
    del out, ana
    np.random.seed(4)
    turb = ch.run(pupil=False)
    assert m.eta(ch, output=turb) == pytest.approx(1.0, abs=3e-5)


def test_time_series_and_si_records_match_reference():
    """§8f n2/n3: TimeBWcorrSimulation + TimeCoherenceResult (frozen flow, shift=(0,t), wind=True) and SIResult through
    the generic Simulation route reproduce the reference's records for the same seed."""
    g = load_golden("timeseries128")
    p = g["params"]
    pa = _pa("complex64", rng="numpy")
    times = tuple(float(t) for t in g["times"])
    count = g["mean_x"].shape[0]
    ch = build_channel(pa, p)
    bw = pa.simulations.TimeBWcorrSimulation(ch, times, max_size=count)
    tc = pa.simulations.TimeCoherenceResult(ch, times, max_size=count)
    sim = pa.simulations.Simulation([bw, tc])
    assert not sim.batchable()
    np.random.seed(int(g["seed"]))
    sim.run()
    # vs the reference's complex64 run: limited by the reference's own screen error (5e-3 rel on the field)
    assert np.allclose(np.asarray(bw.measures[0]), g["mean_x"], rtol=2e-2, atol=3e-5)
    assert np.allclose(np.asarray(bw.measures[1]), g["mean_y"], rtol=2e-2, atol=3e-5)
    assert np.allclose(np.asarray(tc.measures[0]), g["eta"], rtol=2e-3)
    assert len(bw.xx) == len(times) and len(tc.tc) == len(times) and tc.tc[0] == pytest.approx(1.0)
    ch2 = build_channel(pa, p)
    si = pa.simulations.SIResult(ch2, max_size=count)
    np.random.seed(int(g["seed"]))
    pa.simulations.Simulation([si]).run()
    assert np.allclose(si.intensities_at_center, g["i0"], rtol=2e-2)
    assert si.si.shape == (p["count"] + 1,) and np.allclose(si.positions[-1], p["length"])


def test_simulation_sharded_over_two_gpus_matches_single_process(tmp_path):
    """§8e: realizations sharded over ranks (NCCL), records gathered on every rank == single-process records."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from conftest import ROOT
    import os
    outs = {}
    for world in (1, 2):
        out = str(tmp_path / f"rec{world}.npy")
        env = dict(os.environ, PYATM_NCCL_OUT=out)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(29520 + world),
                            os.path.join(ROOT, "tests", "_nccl_sim_worker.py")], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
        outs[world] = np.load(out)
    assert np.array_equal(outs[1], outs[2])


def test_error_paths_and_many_apertures():
    """Status codes surface as exceptions with the library's message; more apertures than one sweep holds are chunked."""
    pa = _pa("complex64")
    from pyatmosphere_b200 import _native as nat
    import torch
    from pyatmosphere_b200.gpu import DeviceArray
    bad = pa.Channel(grid=pa.RectGrid(100, 1e-3), source=pa.GaussianSource(wvl=808e-9, w0=0.02, F0=np.inf),
                     path=pa.VacuumPath(length=10.0), pupil=pa.CirclePupil(radius=0.01))
    with pytest.raises(nat.NativeError, match="power of two"):
        bad.run()
    with pytest.raises(ValueError):
        pa.Channel(grid=pa.RectGrid((128, 256), 1e-3), source=pa.GaussianSource(wvl=808e-9, w0=0.02, F0=np.inf),
                   path=pa.VacuumPath(length=10.0), pupil=pa.CirclePupil(radius=0.01)).run()
    g = load_golden("turb128")
    p = g["params"]
    ch = build_channel(pa, p)
    out = DeviceArray(torch.as_tensor(g["field"].astype(np.complex64)).cuda())
    pupils = [(0.01 * (i + 1), (0.002 * i, -0.001 * i)) for i in range(11)]
    got = pa.measures.all_moments(ch, out, pupils)["eta_pupil"][0]
    x, y = orc.rect_xy(p["n"], p["delta"])
    want = orc.moments(g["field"].astype(np.complex64), x, y, p["delta"], pupils=pupils, mode="f64")["eta_pupil"]
    assert got.shape == (11,) and np.allclose(got, want, rtol=2e-6)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_split_column_pass_random_field_8192(dtype):
    """8192^2 grids use the split column pass (outer radix-32 stage + 256-point tiles, fft_split.cuh): a random
    field through one leg against the float64 oracle, which exercises every frequency of the composite order."""
    pa = _pa(dtype)
    import torch
    from pyatmosphere_b200.gpu import DeviceArray
    n, delta, length, wvl = 8192, 1e-3, 1.2e3, 808e-9
    rng = np.random.default_rng(8192)
    u = (rng.standard_normal((n, n), dtype=np.float32) + 1j * rng.standard_normal((n, n), dtype=np.float32)).astype(dtype)
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=0.05, F0=np.inf),
                    path=pa.VacuumPath(length=length), pupil=pa.CirclePupil(radius=1.0))
    out = ch.path.output(DeviceArray(torch.as_tensor(u).cuda())).get()
    want = orc.vacuum_leg(u, length, wvl, delta, mode="f64")
    assert rel_l2(out, want) < (2e-6 if dtype == "complex64" else 1e-12)


def test_monte_carlo_statistics_match_reference_notebook():
    """The reference's recorded Monte-Carlo run (main.ipynb cells 15-16: README QuickChannel, 2000 samples):
        sigma_BW_x = 3.9e-02 +- 6.2e-04,  sigma_LT_x = 1.8e-01 +- 6.8e-04,  W_ST = 1.6e-01 +- 4.3e-04
    reproduced within Monte-Carlo error by 2000 device-RNG realizations (tensor-core screens, fused statistics).
    Tolerance: 4 combined standard errors plus half a unit of the last printed digit.  The transmittance sample of
    the device RNG is also compared with a host-RNG sample (reference draw order) by a two-sample KS test."""
    from scipy import stats
    pa = _pa("complex64", screen_method="auto", theta_cut=None, rng="philox", seed=2021, batch=16)
    ch = pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12)
    beam = pa.simulations.BeamResult(ch, max_size=2000)
    pdt = pa.simulations.PDTResult(ch, max_size=2000)
    pa.simulations.Simulation([beam, pdt]).run()
    assert len(beam.measures[0]) == 2000 and len(pdt.measures[0]) == 2000
    recorded = {"bw": (3.9e-2, 6.2e-4, 0.05e-2), "lt": (1.8e-1, 6.8e-4, 0.05e-1), "st": (1.6e-1, 4.3e-4, 0.05e-1)}
    for key, (ref, ref_err, half_digit) in recorded.items():
        val, err = getattr(beam, key)
        assert abs(val - ref) < 4 * np.hypot(err, ref_err) + half_digit, (key, val, err)
        assert 0.5 * ref_err < err < 2 * ref_err, (key, err)              # same sample size -> same error bar
    eta_dev = np.asarray(pdt.measures[0].data)
    assert np.all((eta_dev > 0) & (eta_dev < 1))
    pa.gpu.config.update(rng="numpy", batch=16)
    ch2 = pa.QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12)
    pdt2 = pa.simulations.PDTResult(ch2, max_size=320)
    np.random.seed(99)
    pa.simulations.Simulation([pdt2]).run()
    eta_host = np.asarray(pdt2.measures[0].data)
    assert stats.ks_2samp(eta_dev, eta_host).pvalue > 1e-3
