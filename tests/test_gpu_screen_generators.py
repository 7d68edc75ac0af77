"""Parity of the other screen generators through the CUDA path: SUPhaseScreen
(phase_screens.py:154-179, pa_screen_ss with sampled-spectrum coefficients) and FFTPhaseScreen (phase_screens.py:37-67,
pa_screen_fft: gather into spectrum storage order, inverse column + row passes, subharmonic sum, mean removal).

Tolerances (stated per test): screens against the float64 oracle on the reference's own seeded draws -- 1e-5 rad rms
per radian of screen amplitude for complex64, 1e-11 for complex128; fields 1e-5 / 1e-10 relative L2 as in
tests/test_gpu_parity.py; against the reference's own output 5e-3 (its complex64 harmonic sums are the floor) for
SU and 2e-6 for FFT screens (the reference computes those in double precision under numpy >= 2)."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import splitstep as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cfg():
    import pyatmosphere_b200 as pa
    saved = dict(pa.gpu.config)
    yield
    pa.gpu.config.clear()
    pa.gpu.config.update(saved)


def _pa(dtype="complex64", **kw):
    import pyatmosphere_b200 as pa
    kw.setdefault("screen_method", "exact")
    kw.setdefault("theta_cut", 2.0)
    pa.gpu.config.update(use_gpu=True, dtype=dtype, **kw)
    return pa


def _su_channel(pa, p, model=None):
    return pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SUPhaseScreen(
                model=model or pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))


def _fft_channel(pa, p):
    return pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.FFTPhaseScreen(p["subharmonics"], model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))


def _su_oracle(g, psd_n=orc.mvk_psd_n):
    p = g["params"] if "params" in g else g
    x, y = orc.rect_xy(p["n"], p["delta"])
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    screens = []
    for _ in range(p["count"]):
        rho, theta, value = orc.draw_su_spectrum(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"], psd_n)
        fx, fy = orc.spectrum_to_fxy(rho, theta)
        screens.append(orc.ss_screen(x, y, fx, fy, value, mode="f64"))
    u0 = orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64")
    pos = orc.screen_positions(p["length"], p["count"])
    out, legs = orc.propagate(u0, screens, p["length"], pos, p["wvl"], p["delta"], mode="f64", keep_legs=True)
    return screens, legs, out


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_su_channel_vs_reference_and_oracle(dtype):
    pa = _pa(dtype)
    g = load_golden("su128")
    p = g["params"]
    ch = _su_channel(pa, p)
    np.random.seed(int(g["seed"]))
    out = ch.run(pupil=False).get()                       # fused propagator (SU screens share one log-polar grid)
    np.random.seed(int(g["seed"]))
    screens64, legs64, want = _su_oracle(g)
    tol = 1e-5 if dtype == "complex64" else 1e-10
    assert rel_l2(out, want) < tol
    assert rel_l2(out, g["field"]) < 5e-3
    # step-by-step generator: every screen and every intermediate field
    np.random.seed(int(g["seed"]))
    steps = list(ch.generator(pupil=False, store_output=True))
    assert len(steps) == p["count"]
    for s, (u, phi) in enumerate(steps):
        err = np.sqrt(np.mean((phi.get().astype(np.float64) - screens64[s]) ** 2))
        assert err < (1e-5 if dtype == "complex64" else 1e-11) * max(1.0, np.abs(screens64[s]).max())
        assert rel_l2(u.get(), legs64[s]) < tol
        assert np.max(np.abs(phi.get() - g["screens"][s])) < 5e-3
    assert rel_l2(ch.output.get(), want) < tol
    # the screen object on its own: complex screen = (real, imaginary) harmonic sums of the same draw
    ps = ch.path.phase_screens[0]
    np.random.seed(5)
    full = ps.generate(complex=True).get()
    np.random.seed(5)
    x, y = orc.rect_xy(p["n"], p["delta"])
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    rho, theta, value = orc.draw_su_spectrum(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"])
    fx, fy = orc.spectrum_to_fxy(rho, theta)
    want_full = orc.ss_screen(x, y, fx, fy, value, mode="f64", complex_out=True)
    assert rel_l2(full, want_full) < (2e-6 if dtype == "complex64" else 1e-12)


def test_su_andrews_model_and_simulation_records():
    """AndrewsModel spectrum (theory/models.py:94-101) through the SU screen, and a Simulation over an SU channel:
    numpy draws in the reference's order, records equal to the oracle's replay."""
    pa = _pa("complex64")
    p = dict(n=128, delta=4e-3, wvl=808e-9, w0=0.06, Cn2=2e-15, l0=6e-3, L0=1e2, m=64, f_min=1 / 1e2 / 15, f_max=1 / 8e-3,
             length=6e3, count=2, pupil=0.1)
    ch = _su_channel(pa, p, model=pa.AndrewsModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]))
    np.random.seed(11)
    out = ch.run(pupil=False).get()
    np.random.seed(11)
    _, _, want = _su_oracle(p, psd_n=orc.andrews_psd_n)
    assert rel_l2(out, want) < 1e-5
    # Monte-Carlo records (batched route, host-drawn spectra)
    pa.gpu.config.update(rng="numpy", batch=4)
    ch2 = _su_channel(pa, p)
    beam = pa.simulations.BeamResult(ch2, max_size=4)
    pdt = pa.simulations.PDTResult(ch2, max_size=4)
    np.random.seed(21)
    pa.simulations.Simulation([beam, pdt]).run()
    np.random.seed(21)
    x, y = orc.rect_xy(p["n"], p["delta"])
    for r in range(4):
        _, _, field = _su_oracle(p)
        m = orc.moments(field, x, y, p["delta"], pupils=[(p["pupil"], (0, 0))], mode="f64")
        assert beam.measures[0].data[r] == pytest.approx(m["mean_x"], rel=1e-4, abs=1e-7)
        assert beam.measures[2].data[r] == pytest.approx(m["mean_x2"], rel=1e-4)
        assert pdt.measures[0].data[r] == pytest.approx(m["eta_pupil"][0], rel=1e-4)
    # the device RNG draws sparse-spectrum coefficients only
    pa.gpu.config.update(rng="philox")
    with pytest.raises(ValueError):
        pa.simulations.Simulation([pa.simulations.PDTResult(_su_channel(pa, p), max_size=2)]).run()


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_fft_channel_vs_reference_and_oracle(dtype):
    pa = _pa(dtype)
    g = load_golden("fft128")
    p = g["params"]
    x, y = orc.rect_xy(p["n"], p["delta"])
    ch = _fft_channel(pa, p)
    ch.path.init_phase_screens()
    # complex screen of the first draw against the reference's own
    np.random.seed(int(g["seed"]))
    full0 = ch.path.phase_screens[0].generate(complex=True).get()
    assert rel_l2(full0, g["screen0_complex"]) < (2e-6 if dtype == "complex64" else 1e-12)
    assert abs(full0.mean()) < 1e-5 * np.abs(full0).max()
    # per-leg screens and fields, output field
    np.random.seed(int(g["seed"]))
    steps = list(ch.generator(pupil=False, store_output=True))
    for s, (u, phi) in enumerate(steps):
        assert rel_l2(phi.get(), g["screens"][s]) < (2e-6 if dtype == "complex64" else 1e-12)
        assert rel_l2(u.get(), g["legs"][s]) < (2e-5 if dtype == "complex64" else 1e-6)     # reference legs are complex64-rounded
    np.random.seed(int(g["seed"]))
    out = ch.run(pupil=False).get()
    assert rel_l2(out, g["field"]) < (2e-5 if dtype == "complex64" else 1e-6)
    # float64 oracle on the same draws
    np.random.seed(int(g["seed"]))
    screens = []
    for _ in range(p["count"]):
        cn, terms = orc.draw_fft_screen(p["n"], p["delta"], p["subharmonics"], p["Cn2"], p["l0"], p["L0"], p["wvl"],
                                        p["length"] / p["count"])
        screens.append(orc.fft_screen(cn, terms, x, y, mode="f64").real)
    u0 = orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64")
    want = orc.propagate(u0, screens, p["length"], orc.screen_positions(p["length"], p["count"]), p["wvl"], p["delta"], mode="f64")
    assert rel_l2(out, want) < (2e-5 if dtype == "complex64" else 1e-10)
    with pytest.raises(TypeError):                          # like the reference: this generator takes no shift / wind
        next(ch.generator(pupil=False, shift=(0, 0.1), wind=True))


@pytest.mark.parametrize("n,dtype", [(64, "complex64"), (256, "complex128"), (512, "complex64"), (1024, "complex64"),
                                     (2048, "complex64")])
def test_pa_screen_fft_random_spectrum_all_sizes(n, dtype):
    """pa_screen_fft on a random spectrum == centred inverse DFT + terms - mean (float64 oracle) for every plan family;
    two screens per call (batch index), real-only and complex outputs, linearity at the full size."""
    pa = _pa(dtype)
    import torch
    from pyatmosphere_b200 import _engine as eng, _native as nat
    grid = pa.RectGrid(n, 2e-3)
    ctx = eng.grid_context(grid)
    rng = np.random.default_rng(n)
    cn = (rng.standard_normal((2, n, n)) + 1j * rng.standard_normal((2, n, n))).astype(dtype)
    terms = np.ascontiguousarray(rng.standard_normal((2, 5, 4)) * np.array([0.3, 0.3, 40.0, 40.0]))
    spec = torch.as_tensor(cn).cuda()
    out_c = torch.empty((2, n, n), dtype=ctx.cdtype, device="cuda")
    out_r = torch.empty((2, n, n), dtype=ctx.rdtype, device="cuda")
    nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(spec), 2, nat.ptr(terms), 5, nat.ptr(out_c), nat.ptr(out_r), nat.stream_ptr()))
    torch.cuda.synchronize()
    got = out_c.cpu().numpy()
    assert np.array_equal(out_r.cpu().numpy(), got.real)
    x, y = grid.get_xy()
    tol = 3e-6 if dtype == "complex64" else 1e-12
    for b in range(2 if n <= 1024 else 1):
        tl = [(terms[b, t, 0], terms[b, t, 1], complex(terms[b, t, 2], terms[b, t, 3])) for t in range(5)]
        want = orc.fft_screen(cn[b], tl, x, y, mode="f64")
        assert rel_l2(got[b], want) < tol
    # linearity without terms: F(a + 2 b) = F(a) + 2 F(b)
    mix = torch.as_tensor((cn[0] + 2 * cn[1]).astype(dtype)[None]).cuda()
    o0 = torch.empty((1, n, n), dtype=ctx.cdtype, device="cuda")
    nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(mix), 1, None, 0, nat.ptr(o0), None, nat.stream_ptr()))
    o2 = torch.empty((2, n, n), dtype=ctx.cdtype, device="cuda")
    nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(spec), 2, None, 0, nat.ptr(o2), None, nat.stream_ptr()))
    torch.cuda.synchronize()
    a = o0.cpu().numpy()[0]
    b2 = o2.cpu().numpy()
    assert rel_l2(a, b2[0] + 2 * b2[1]) < (3e-6 if dtype == "complex64" else 1e-12)
    with pytest.raises(nat.NativeError):
        nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(spec), 2, None, 0, None, None, nat.stream_ptr()))


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_wind_su_series_vs_reference_and_oracle(dtype):
    """Three successive runs of a WindSUPhaseScreen channel (frozen flow, speed per call) against the reference's
    outputs (5e-3: its complex64 harmonic factors) and against the float64 oracle fed the coefficients as the kernels
    receive them (complex64): 1e-5 / 1e-10."""
    pa = _pa(dtype)
    g = load_golden("windsu128")
    p = g["params"]
    speed = float(g["speed"])
    ch = pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]), source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.WindSUPhaseScreen(pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"]), speed,
                                              model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))
    np.random.seed(int(g["seed"]))
    outs = [ch.run(pupil=False).get() for _ in range(g["fields"].shape[0])]
    x, y = orc.rect_xy(p["n"], p["delta"])
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    np.random.seed(int(g["seed"]))
    spectra = [orc.draw_wind_su_spectrum(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"]) for _ in range(p["count"])]
    u0 = orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64")
    pos = orc.screen_positions(p["length"], p["count"])
    for k, out in enumerate(outs):
        screens = []
        for rho, theta, value in spectra:
            fx, fy = orc.spectrum_to_fxy(rho, theta)
            screens.append(orc.ss_screen(x, y, fx, fy, value.astype(np.complex64), shift=(k * speed, 0), mode="f64"))
        want = orc.propagate(u0, screens, p["length"], pos, p["wvl"], p["delta"], mode="f64")
        assert rel_l2(out, want) < (1e-5 if dtype == "complex64" else 1e-10)
        assert rel_l2(out, g["fields"][k]) < 5e-3
    assert rel_l2(outs[1], outs[0]) > 1e-3                 # the screens did move
    scr = next(ch.path.phase_screens[0].generator())
    assert scr.shape == (p["n"], p["n"])
