"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference imports matplotlib at import time; `oracle/mpl_stub` satisfies the import, nothing plots.
Every fixture stores the inputs (parameters, seed, exported random coefficients) next to the reference's
outputs so that the tests can (a) pin `oracle/splitstep.py` against the reference and (b) feed the same
coefficients to the CUDA path.  Versions are recorded in each file (`versions`).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYATM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "mpl_stub"))
    warnings.simplefilter("ignore", SyntaxWarning)
    import pyatmosphere  # noqa: F401  (the reference)
    return pyatmosphere


def _versions():
    import scipy
    return np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}"])


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
        pupil=pa.CirclePupil(radius=p["pupil"]),
    )


def export_realization(pa, ch, seed):
    """Replays the reference's draw order screen by screen, then re-seeds and runs the reference itself."""
    ch.path.init_phase_screens()
    np.random.seed(seed)
    rho, theta, value = [], [], []
    for ps in ch.path.phase_screens:
        sp = ps._get_spectrum(False)
        rho.append(np.asarray(sp.rho)); theta.append(np.asarray(sp.theta)); value.append(np.asarray(sp.value))
    psd = np.asarray(ch.path.phase_screens[0]._get_psd())
    np.random.seed(seed)
    legs = []
    gen = ch.generator(pupil=False, store_output=True)
    screens = []
    for u, phi in gen:
        legs.append(np.asarray(u)); screens.append(np.asarray(phi))
    out = np.asarray(ch.output)
    np.random.seed(seed)
    out_run = np.asarray(ch.run(pupil=False))
    # Channel.generator stores the generator's return value (part losses only, pathes.py:71-75) while
    # Channel.run goes through AbstractPath.output, which applies the full dB loss once more (pathes.py:23-24).
    return dict(rho=np.stack(rho), theta=np.stack(theta), value=np.stack(value), psd=psd,
                screens=np.stack(screens), legs=np.stack(legs), field=out_run, field_generator=out)


def measures_of(pa, ch, out):
    m = pa.measures
    names = ["eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2"]
    vals = [getattr(m, k)(ch, output=out) for k in names]
    eta_pupil = m.eta(ch, output=ch.pupil.output(out))
    return np.array(vals + [eta_pupil], dtype=np.float64)


def params_arrays(p):
    keys = sorted(p)
    return dict(param_keys=np.array(keys), param_vals=np.array([str(p[k]) for k in keys]))


def case_vacuum(pa):
    """Restated tests/itest_vacuum_propagation.ipynb cells 3-7 on the current reference tree."""
    n, delta, wvl, w0, length = 256, 2e-3, 809e-9, 2e-2, 4e3
    ch = pa.Channel(grid=pa.RectGrid(resolution=n, delta=delta),
                    source=pa.GaussianSource(wvl=wvl, w0=w0, F0=np.inf),
                    path=pa.pathes.VacuumPath(length=length), pupil=pa.CirclePupil(radius=1.0))
    src = np.asarray(ch.source.output())
    out = np.asarray(ch.run(pupil=False))
    m = measures_of(pa, ch, out)
    np.savez_compressed(os.path.join(OUT, "vacuum256.npz"), n=n, delta=delta, wvl=wvl, w0=w0, length=length,
                        source=src.astype(np.complex128), field=out, measures=m,
                        abs_sum=np.abs(out.sum()) * delta**2, versions=_versions())


def case_turbulent(pa, name, p, seed, with_legs=True):
    ch = build_channel(pa, p)
    r = export_realization(pa, ch, seed)
    m = measures_of(pa, ch, r["field"])
    extra = {}
    if with_legs:
        extra = dict(screens=r["screens"], legs=r["legs"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, rho=r["rho"], theta=r["theta"],
                        value=r["value"], psd=r["psd"], field=r["field"], field_generator=r["field_generator"],
                        measures=m,
                        positions=np.asarray(ch.path.positions), versions=_versions(), **params_arrays(p),
                        **extra)


def case_psd(pa):
    """Ring PSDs of the two README channels (independent of the spatial grid)."""
    c1 = dict(n=64, delta=1e-3, wvl=808e-9, w0=0.09, Cn2=1e-15, l0=3e-3, L0=1e3, m=2**10,
              f_min=1 / 1e3 / 15, f_max=1 / 3e-3 * 2, length=10e3, count=5, pupil=0.12)
    c3 = dict(n=64, delta=1.5e-3, wvl=808e-9, w0=0.12, Cn2=5e-16, l0=6e-3, L0=1e3, m=2**10,
              f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2, length=50e3, count=5, pupil=0.2)
    out = {}
    for tag, p in (("c1", c1), ("c3", c3)):
        ch = build_channel(pa, p)
        ch.path.init_phase_screens()
        ps = ch.path.phase_screens[0]
        out[tag + "_psd"] = np.asarray(ps._get_psd())
        out[tag + "_base"] = np.asarray(ps.f_grid.base)
        out[tag + "_rytov2"] = ch.get_rythov2()
    np.savez_compressed(os.path.join(OUT, "psd_readme.npz"), versions=_versions(), **out)


def case_simulation(pa, p, seed, count):
    """Per-realization scalar tables of BeamResult + PDTResult from the reference's Simulation loop."""
    ch = build_channel(pa, p)
    beam = pa.simulations.BeamResult(ch, max_size=count)
    pdt = pa.simulations.PDTResult(ch, max_size=count)
    sim = pa.simulations.Simulation([beam, pdt])
    np.random.seed(seed)
    sim.run()
    table = np.array([mm.data for mm in beam.measures] + [pdt.measures[0].data], dtype=np.float64).T
    names = np.array([mm.name for mm in beam.measures] + [pdt.measures[0].name])
    stats = np.array([beam.bw, beam.lt, beam.st], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "simulation128.npz"), seed=seed, table=table, names=names,
                        stats=stats, versions=_versions(), **params_arrays(p))


def case_time_series(pa, p, seed, count, times):
    """Frozen-flow records: TimeBWcorrSimulation (mean_x, mean_y per time lag), TimeCoherenceResult (eta behind the
    channel's pupil per time lag) and SIResult-style on-axis intensity per leg, from the reference's Simulation loop."""
    ch = build_channel(pa, p)
    bw = pa.simulations.TimeBWcorrSimulation(ch, times, max_size=count)
    tc = pa.simulations.TimeCoherenceResult(ch, times, max_size=count)
    sim = pa.simulations.Simulation([bw, tc])
    np.random.seed(seed)
    sim.run()
    # on-axis intensity per leg (SIResult's record) in a simulation of its own: mixing it with time-lag records runs into
    # the reference's broken cached-screen branch (phase_screens.py:117-121)
    ch2 = build_channel(pa, p)
    si = pa.simulations.Measure(ch2, "propagation", pa.simulations.si.intensity_at_center, name="i0", max_size=count)
    sim2 = pa.simulations.Simulation([], [si])
    np.random.seed(seed)
    sim2.run()
    np.savez_compressed(os.path.join(OUT, "timeseries128.npz"), seed=seed, times=np.array(times),
                        mean_x=np.array(bw.measures[0].data, dtype=np.float64), mean_y=np.array(bw.measures[1].data, dtype=np.float64),
                        eta=np.array(tc.measures[0].data, dtype=np.float64), i0=np.array(si.data, dtype=np.float64),
                        versions=_versions(), **params_arrays(p))


def _run_and_record(ch, seed):
    """Channel.generator drained (per-leg fields and screens), then Channel.run re-seeded, as export_realization."""
    np.random.seed(seed)
    legs, screens = [], []
    for u, phi in ch.generator(pupil=False, store_output=True):
        legs.append(np.asarray(u)); screens.append(np.asarray(phi))
    np.random.seed(seed)
    out_run = np.asarray(ch.run(pupil=False))
    return np.stack(screens), np.stack(legs), out_run


def case_su(pa, p, seed):
    """SUPhaseScreen channel (phase_screens.py:154-179): exported draws, per-leg screens and fields, output field."""
    ch = pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SUPhaseScreen(
                model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))
    screens, legs, out = _run_and_record(ch, seed)
    np.savez_compressed(os.path.join(OUT, "su128.npz"), seed=seed, screens=screens, legs=legs, field=out,
                        measures=measures_of(pa, ch, out), versions=_versions(), **params_arrays(p))


def case_wind_su(pa, p, seed, speed, calls):
    """WindSUPhaseScreen channel (phase_screens.py:182-215): `calls` successive Channel.run outputs of one channel
    (the screens translate by `speed` per call) and the screens of the first call."""
    ch = pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.WindSUPhaseScreen(
                pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"]), speed,
                model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))
    np.random.seed(seed)
    screens0 = []
    fields = []
    for u, phi in ch.generator(pupil=False, store_output=True):
        screens0.append(np.asarray(phi))
    fields.append(np.asarray(ch.output))
    for _ in range(calls - 1):
        fields.append(np.asarray(ch.run(pupil=False)))
    cn0 = np.asarray(ch.path.phase_screens[0].cnp)
    np.savez_compressed(os.path.join(OUT, "windsu128.npz"), seed=seed, speed=speed, screens0=np.stack(screens0),
                        fields=np.stack(fields), cnp0=cn0, versions=_versions(), **params_arrays(p))


def case_fft(pa, p, seed):
    """FFTPhaseScreen channel (phase_screens.py:37-67) with subharmonic levels.  The coefficient arrays handed to
    ifft2 are captured by wrapping the name the reference module calls (the reference code itself is untouched)."""
    ch = pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.FFTPhaseScreen(p["subharmonics"], model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])),
            length=p["length"], count=p["count"]),
        pupil=pa.CirclePupil(radius=p["pupil"]))
    mod = pa.phase_screens
    captured, plain = [], mod.ifft2

    def spy(x, delta):
        captured.append(np.array(x))
        return plain(x, delta)

    mod.ifft2 = spy
    try:
        screens, legs, out = _run_and_record(ch, seed)
    finally:
        mod.ifft2 = plain
    cn0 = captured[0]
    # one complex screen straight from the generator's first draw (real and imaginary parts both checked)
    ch.path.init_phase_screens()
    np.random.seed(seed)
    full0 = np.asarray(ch.path.phase_screens[0].generate(complex=True))
    np.savez_compressed(os.path.join(OUT, "fft128.npz"), seed=seed, screens=screens, legs=legs, field=out,
                        cn0_checksum=np.array([cn0.sum(), np.abs(cn0).sum()]), cn0_corner=cn0[:8, :8], cn0_dtype=str(cn0.dtype),
                        screen0_complex=full0, measures=measures_of(pa, ch, out), versions=_versions(), **params_arrays(p))


def case_c3(pa, seeds=(3, 4)):
    """Config 3 -- the configuration the bench metric is quoted on (README advanced channel: 2048^2, 5 SS screens,
    50 km) -- run by the UNMODIFIED reference for each seed.  Stored per seed: the exported draws (rho, theta, value),
    the reference's own measures (measures.py:7-38 on its complex64 output, eta behind the 0.2 m pupil), and a 64 x 64
    centre crop of its output field.  Next to them the float64 restatement (oracle/splitstep.py mode='f64') on the same
    draws: its seven moments, eta_pupil and the same crop, so that the GPU box can check the fused Monte-Carlo route
    without re-running 2048^2 float64 GEMMs for every test (tests that do want the full field run the oracle there)."""
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import splitstep as orc
    p = dict(n=2048, delta=1.5e-3, wvl=808e-9, w0=0.12, Cn2=5e-16, l0=6e-3, L0=1e3, m=2**10,
             f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2, length=50e3, count=5, pupil=0.2)
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    x, y = orc.rect_xy(p["n"], p["delta"])
    c0, c1 = p["n"] // 2 - 32, p["n"] // 2 + 32
    keys = ("eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2", "mean_x2_r")
    rec = {k: [] for k in ("rho", "theta", "value", "ref_measures", "ref_crop", "f64_measures", "f64_crop", "ref_vs_f64")}
    for seed in seeds:
        np.random.seed(seed)
        rho, theta, value = [], [], []
        for ps in ch.path.phase_screens:
            sp = ps._get_spectrum(False)
            rho.append(np.asarray(sp.rho)); theta.append(np.asarray(sp.theta)); value.append(np.asarray(sp.value))
        np.random.seed(seed)
        out = np.asarray(ch.run(pupil=False))
        rec["rho"].append(np.stack(rho)); rec["theta"].append(np.stack(theta)); rec["value"].append(np.stack(value))
        rec["ref_measures"].append(measures_of(pa, ch, out))
        rec["ref_crop"].append(out[c0:c1, c0:c1].copy())
        screens = []
        for s in range(p["count"]):
            fx, fy = orc.spectrum_to_fxy(rho[s], theta[s])
            screens.append(orc.ss_screen(x, y, fx, fy, value[s], mode="f64"))
        want = orc.propagate(orc.gaussian_source(x, y, p["w0"], p["wvl"], mode="f64"), screens, p["length"],
                             orc.screen_positions(p["length"], p["count"]), p["wvl"], p["delta"], mode="f64")
        m = orc.moments(want, x, y, p["delta"], pupils=[(p["pupil"], (0, 0))], mode="f64")
        rec["f64_measures"].append(np.array([m[k] for k in keys] + [m["eta_pupil"][0]], dtype=np.float64))
        rec["f64_crop"].append(want[c0:c1, c0:c1].copy())
        rec["ref_vs_f64"].append(np.linalg.norm(out - want) / np.linalg.norm(want))
        print("c3 seed", seed, "reference vs float64 restatement rel-L2", rec["ref_vs_f64"][-1], "eta_pupil", m["eta_pupil"][0])
    np.savez_compressed(os.path.join(OUT, "c3_2048.npz"), seeds=np.array(seeds), psd=np.asarray(ch.path.phase_screens[0]._get_psd()),
                        f64_names=np.array(list(keys) + ["eta_pupil"]),
                        ref_names=np.array(["eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2", "eta_pupil"]),
                        crop=np.array([c0, c1]), versions=_versions(), **params_arrays(p),
                        **{k: np.stack(v) for k, v in rec.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    pa = _import_reference()
    case_vacuum(pa)
    case_psd(pa)
    small = dict(n=128, delta=4e-3, wvl=808e-9, w0=0.06, Cn2=2e-15, l0=6e-3, L0=1e2, m=96,
                 f_min=1 / 1e2 / 15, f_max=1 / 8e-3, length=6e3, count=3, pupil=0.1)
    case_turbulent(pa, "turb128", small, seed=1234)
    lossy = dict(small, where="after", losses_db=1.5, count=2, F0=4e3, m=64)
    case_turbulent(pa, "turb128_after_lossy", lossy, seed=7)
    before = dict(small, where="before", count=2, m=64, n=64, delta=6e-3)
    case_turbulent(pa, "turb64_before", before, seed=99)
    quick = dict(n=256, delta=1e-3 * 4, wvl=808e-9, w0=0.09, Cn2=1e-15, l0=3e-3, L0=1e3, m=2**10,
                 f_min=1 / 1e3 / 15, f_max=1 / 3e-3 * 2, length=10e3, count=5, pupil=0.12)
    case_turbulent(pa, "quick256", quick, seed=5, with_legs=False)
    case_simulation(pa, small, seed=2024, count=6)
    case_time_series(pa, small, seed=77, count=3, times=(0.0, 0.012, 0.05))
    case_su(pa, dict(small, count=2), seed=31)
    case_wind_su(pa, dict(small, count=2), seed=17, speed=0.011, calls=3)
    case_fft(pa, dict(n=128, delta=4e-3, wvl=808e-9, w0=0.06, Cn2=2e-15, l0=6e-3, L0=1e2, length=6e3, count=2,
                      pupil=0.1, subharmonics=2), seed=13)
    case_c3(pa)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--only-n4":      # add the SU / FFT fixtures without touching the others
        os.makedirs(OUT, exist_ok=True)
        _pa = _import_reference()
        _small = dict(n=128, delta=4e-3, wvl=808e-9, w0=0.06, Cn2=2e-15, l0=6e-3, L0=1e2, m=96,
                      f_min=1 / 1e2 / 15, f_max=1 / 8e-3, length=6e3, count=2, pupil=0.1)
        case_su(_pa, _small, seed=31)
        case_wind_su(_pa, _small, seed=17, speed=0.011, calls=3)
        case_fft(_pa, dict(n=128, delta=4e-3, wvl=808e-9, w0=0.06, Cn2=2e-15, l0=6e-3, L0=1e2, length=6e3, count=2,
                           pupil=0.1, subharmonics=2), seed=13)
    elif len(sys.argv) > 1 and sys.argv[1] == "--only-c3":    # add the config-3 fixture without touching the others
        os.makedirs(OUT, exist_ok=True)
        case_c3(_import_reference())
    else:
        main()
