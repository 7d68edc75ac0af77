"""CPU-only tests: host mirror of the reference API, the C-ABI surface, the planning of the ring split and the
multi-rank reduction (gloo, world_size 2).  No compute call is made here (there is no GPU)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import splitstep as orc

import pyatmosphere_b200 as pa
from pyatmosphere_b200 import _engine as eng
from pyatmosphere_b200 import _native as nat
from pyatmosphere_b200 import distributed as dist


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pyatm_b200.h")).read()
    declared = set(re.findall(r"PA_API[^;(]*?\b(pa_\w+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    lib = nat.load()                       # raises if the .so was not built: there is no fallback
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pa_version() == 100


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ch = pa.QuickChannel(grid_resolution=64)
    with pytest.raises(nat.NativeError):
        ch.run()
    pa.gpu.config["use_gpu"] = False
    try:
        with pytest.raises(pa.gpu.NoCpuPathError):
            ch.run()
    finally:
        pa.gpu.config["use_gpu"] = True


def test_public_names_match_reference_surface():
    for name in ["Channel", "QuickChannel", "RectGrid", "RandLogPolarGrid", "GaussianSource", "IdenticalPhaseScreensPath",
                 "PhaseScreensPath", "VacuumPath", "SSPhaseScreen", "CirclePupil", "MVKModel", "measures", "simulations", "gpu"]:
        assert hasattr(pa, name), name
    for name in ["I", "eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2"]:
        assert callable(getattr(pa.measures, name))
    for name in ["Simulation", "Measure", "Result", "BeamResult", "PDTResult", "TrackedPDTResult", "SIResult", "WindResult",
                 "TimeCoherenceResult", "TimeBWcorrSimulation"]:
        assert hasattr(pa.simulations, name)
    assert "use_gpu" in pa.gpu.config and callable(pa.gpu.get_array) and callable(pa.gpu.get_xp)


def test_grids_match_oracle_bitwise():
    for n, d in [(128, 4e-3), (2048, 1.5e-3), (1024, 1e-3), (7, 0.3)]:
        g = pa.RectGrid(n, d)
        x, y = orc.rect_xy(n, d)
        assert np.array_equal(g.get_x(), x) and np.array_equal(g.get_y(), y)
        assert g.get_x().dtype == np.float32 and g.get_x().shape == (1, n) and g.get_y().shape == (n, 1)
        assert g.origin_index == (n // 2, n // 2)
        assert g.get_f_grid().delta == orc.f_grid_delta(n, d)
        assert np.array_equal(g.get_rho2(), x**2 + y**2)
    assert list(pa.RectGrid(4, 0.5).extent) == [-1.0, 1.0, -1.0, 1.0]
    lp = pa.RandLogPolarGrid(points=2**10, f_min=1 / 1e3 / 15, f_max=1 / 6e-3 * 2)
    assert np.array_equal(lp.base, orc.logpolar_base(2**10, 1 / 1e3 / 15, 1 / 6e-3 * 2))


def _channel(p):
    return pa.Channel(
        grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]),
        source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=p.get("F0", np.inf)),
        path=pa.IdenticalPhaseScreensPath(
            phase_screen=pa.SSPhaseScreen(model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                                          f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])),
            length=p["length"], count=p["count"], position_in_slab=p.get("where", "middle"), losses_db=p.get("losses_db", 0)),
        pupil=pa.CirclePupil(radius=p["pupil"]))


@pytest.mark.parametrize("name", ["turb128", "turb128_after_lossy", "turb64_before", "quick256"])
def test_spectra_drawn_in_reference_order(name):
    """np.random.seed(s) + the host mirror reproduces the coefficients the reference exported."""
    g = load_golden(name)
    ch = _channel(g["params"])
    ch.path.init_phase_screens()
    assert np.array_equal(ch.path.phase_screens[0]._get_psd(), g["psd"])
    assert np.array_equal(np.asarray(ch.path.positions), g["positions"])
    np.random.seed(int(g["seed"]))
    for s, ps in enumerate(ch.path.phase_screens):
        sp = ps._get_spectrum(False)
        assert np.array_equal(sp.rho, g["rho"][s]) and np.array_equal(sp.theta, g["theta"][s])
        assert np.array_equal(sp.value, g["value"][s])
    np.random.seed(int(g["seed"]))
    fx, fy, cf = eng.draw_spectra_numpy(ch.path, 1)
    ofx, ofy = orc.spectrum_to_fxy(g["rho"][1], g["theta"][1])
    assert np.array_equal(fx[0, 1], ofx.ravel()) and np.array_equal(fy[0, 1], ofy.ravel()) and np.array_equal(cf[0, 1], g["value"][1])


def test_path_geometry_and_losses():
    p = load_golden("turb128_after_lossy")["params"]
    ch = _channel(p)
    legs = ch.path.leg_lengths()
    assert legs == orc.leg_lengths(p["length"], orc.screen_positions(p["length"], p["count"], "after"))
    assert legs[-1] == 0.0
    sc = eng.path_losses(ch.path, legs)
    assert np.allclose(sc, [10 ** (-(p["losses_db"] * l / p["length"]) / 20) for l in legs[:-1]])
    before = _channel(dict(p, where="before"))
    lb = before.path.leg_lengths()
    assert lb[0] == 0.0
    # zero share falls back to the full loss (pathes.py:20)
    assert eng.path_losses(before.path, lb)[0] == pytest.approx(10 ** (-p["losses_db"] / 20))
    with pytest.raises(ValueError):
        pa.IdenticalPhaseScreensPath(length=1.0, count=2, phase_screen=ch.path.phase_screen, position_in_slab="centre")
    assert ch.get_rythov2() == pytest.approx(orc.rytov2(p["Cn2"], 2 * np.pi / p["wvl"], p["length"]))


def test_low_ring_plan_bounds_truncation_error():
    """The (m_split, degree) the host picks keeps the Taylor remainder of every low ring under the tolerance."""
    g = load_golden("psd_readme")
    base, psd = g["c3_base"], g["c3_psd"]
    ext = 1024 * 1.5e-3
    for theta_cut, tol in [(2.0, 1e-7), (2.0, 1e-12), (0.5, 1e-12), (6.0, 1e-7)]:
        ms, deg = eng.plan_low_rings(base, psd, ext, ext, theta_cut, tol)
        assert 0 < ms < len(base) and 0 < deg <= eng.MAX_DEGREE
        t = 2 * np.pi * base[:ms].astype(np.float64) * np.hypot(ext, ext)
        assert t.max() <= theta_cut
        from math import factorial
        rem = np.sum(6 * np.sqrt(psd[:ms].astype(np.float64)) * t ** (deg + 1) / factorial(deg + 1))
        assert rem <= tol
    assert eng.plan_low_rings(base, psd, ext, ext, 0.0, 1e-7) == (0, -1)
    assert eng.plan_low_rings(base[::-1], psd, ext, ext, 2.0, 1e-7) == (0, -1)     # unsorted rings: no split
    # numerically: polynomial + remaining harmonics == all harmonics (float64 numpy model of the kernel's split)
    rng = np.random.default_rng(0)
    m = 64
    fx = (rng.standard_normal(m) * np.logspace(-4, 0, m)).astype(np.float32)
    fy = (rng.standard_normal(m) * np.logspace(-4, 0, m)).astype(np.float32)
    c = (rng.standard_normal(m) + 1j * rng.standard_normal(m)) * np.logspace(3, -1, m)
    xs = np.linspace(-1.5, 1.5, 33)
    full = np.real(np.sum(c[:, None, None] * np.exp(2j * np.pi * (fy[:, None, None] * xs[None, :, None] + fx[:, None, None] * xs[None, None, :])), axis=0))
    ms, D = 40, 30
    from math import factorial
    poly = np.zeros_like(full)
    for pp in range(D + 1):
        for q in range(D + 1 - pp):
            S = np.sum(c[:ms] * (2 * np.pi * fx[:ms].astype(np.float64)) ** pp * (2 * np.pi * fy[:ms].astype(np.float64)) ** q)
            T = np.real(1j ** (pp + q) * S) / (factorial(pp) * factorial(q))
            poly += T * xs[None, :] ** pp * xs[:, None] ** q
    hi = np.real(np.sum(c[ms:, None, None] * np.exp(2j * np.pi * (fy[ms:, None, None] * xs[None, :, None] + fx[ms:, None, None] * xs[None, None, :])), axis=0))
    assert np.max(np.abs(poly + hi - full)) < 1e-9 * np.max(np.abs(full))


def test_simulation_tree_and_batchability():
    p = load_golden("turb128")["params"]
    ch = _channel(p)
    beam = pa.simulations.BeamResult(ch, max_size=5)
    pdt = pa.simulations.PDTResult(ch, max_size=9)
    sim = pa.simulations.Simulation([beam, pdt])
    assert len(list(sim.flattened_measures())) == 7
    assert sim.batchable() and sim.remaining() == 9 and not sim.is_measures_done()
    assert [m.name for m in beam.measures] == ["mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2", "mean_x2_r"]
    assert pdt.measures[0].name == str(p["pupil"])
    custom = pa.simulations.Measure(ch, "atmosphere", lambda channel, output: 0.0, name="custom", max_size=2)
    assert not pa.simulations.Simulation([beam], [custom]).batchable()
    cols = eng.table_columns([0.1], [0.2])
    assert cols["mean_x"] == 1 and cols[("fixed", 0.1)] == 7 and cols[("tracked", 0.2)] == 8
    for m in beam.measures:
        m.data = [0.1, 0.2, 0.3, 0.4, 0.5]
    assert beam.measures[0].is_done and sim.remaining() == 9
    bw = beam.bw
    want = orc.beam_statistics(beam.measures[0].data, beam.measures[2].data)
    assert bw == pytest.approx(want["bw"]) and beam.lt == pytest.approx(want["lt"]) and beam.st == pytest.approx(want["st"])


def test_result_csv_roundtrip(tmp_path):
    """Same on-disk format as the reference: header = measure names, '%.3e' floats; resumed on construction."""
    p = load_golden("turb128")["params"]
    ch = _channel(p)
    path = str(tmp_path / "beam.csv")
    beam = pa.simulations.BeamResult(ch, max_size=4, save_path=path)
    for i, m in enumerate(beam.measures):
        m.data = [0.123456 * (i + 1), -1.5e-7 * (i + 1)]
    beam.save_output()
    lines = open(path).read().strip().splitlines()
    assert lines[0] == "mean_x,mean_y,mean_x2,mean_xy,mean_y2,mean_x2_r"
    assert lines[1].split(",")[0] == "1.235e-01" and lines[2].split(",")[1] == "-3.000e-07"
    again = pa.simulations.BeamResult(ch, max_size=4, save_path=path)
    assert again.measures[0].data == [0.1235, -1.5e-07] and len(again.measures[5]) == 2


def test_pdt_histogram_binning():
    p = load_golden("turb128")["params"]
    ch = _channel(p)
    pdt = pa.simulations.PDTResult(ch, max_size=10)
    pdt.measures[0].data = [0.0, 0.004999, 0.005, 1.0, 0.9999, 0.5]
    assert np.array_equal(pdt.histogram(), orc.pdt_histogram(pdt.measures[0].data, 200))


def test_shard_indices_partition():
    for first, count, world in [(0, 8, 2), (5, 7, 4), (0, 3, 8), (16, 64, 8)]:
        parts = [dist.shard_indices(first, count, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), np.arange(first, first + count))
        assert max(len(q) for q in parts) - min(len(q) for q in parts) <= 1
        for q in parts:
            assert len(q) == 0 or np.array_equal(q, np.arange(q[0], q[0] + len(q)))


def test_two_rank_gloo_reduction():
    """world_size 2 over gloo: gathered sample table and reduced statistics equal the single-process result."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o, o


# ---- the other screen generators (host side only: draws, bounds, routing) ------------------------------------------
def _n4_channel(screen, n=128, count=2):
    return pa.Channel(grid=pa.RectGrid(resolution=n, delta=4e-3), source=pa.GaussianSource(wvl=808e-9, w0=0.06, F0=np.inf),
                      path=pa.IdenticalPhaseScreensPath(phase_screen=screen, length=6e3, count=count),
                      pupil=pa.CirclePupil(radius=0.1))


def test_su_screen_draws_match_oracle_and_reference():
    g = load_golden("su128")
    p = g["params"]
    ch = _n4_channel(pa.SUPhaseScreen(model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]),
                                      f_grid=pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])))
    ch.path.init_phase_screens()
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    np.random.seed(int(g["seed"]))
    mine = [ps._get_spectrum(False) for ps in ch.path.phase_screens]
    np.random.seed(int(g["seed"]))
    for sp in mine:
        rho, theta, value = orc.draw_su_spectrum(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"])
        assert np.array_equal(sp.rho, rho) and np.array_equal(sp.theta, theta)
        assert sp.value.dtype == np.complex64 and np.array_equal(sp.value, value)
    # the per-ring power handed to the planner bounds E|c|^2/2 at every admissible radius, here at the drawn ones
    ps = ch.path.phase_screens[0]
    power = ps._ring_power()
    dens = orc.psd_phi_f(mine[0].rho.astype(np.float64), p["Cn2"], p["l0"], p["L0"], 2 * np.pi / p["wvl"], ps.thickness)
    assert np.all(power >= dens * np.pi * orc.su_delta_k(base).astype(np.float64))
    assert np.array_equal(ps.delta_k_base, orc.su_delta_k(base))
    # routing: SU screens over one shared log-polar grid go through the fused propagator, but not the device RNG
    assert ch.path._fusable() and not ps.device_rng and pa.SSPhaseScreen.device_rng
    m_split, degree = ps.low_ring_plan()
    assert 0 <= m_split < p["m"] and degree <= eng.MAX_DEGREE


def test_fft_screen_draws_match_oracle():
    g = load_golden("fft128")
    p = g["params"]
    ch = _n4_channel(pa.FFTPhaseScreen(p["subharmonics"], model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])), n=p["n"])
    ch.path.init_phase_screens()
    assert not ch.path._fusable()
    np.random.seed(int(g["seed"]))
    cn, terms = ch.path.phase_screens[0]._draw()
    np.random.seed(int(g["seed"]))
    cn_o, terms_o = orc.draw_fft_screen(p["n"], p["delta"], p["subharmonics"], p["Cn2"], p["l0"], p["L0"], p["wvl"],
                                        p["length"] / p["count"])
    assert cn.dtype == np.complex128 and np.array_equal(cn, cn_o)
    assert np.allclose(cn[:8, :8], g["cn0_corner"], rtol=1e-14, atol=0)
    live = [(fx, fy, c) for fx, fy, c in terms_o if c != 0]
    assert terms.shape == (8 * p["subharmonics"], 4) and len(live) == len(terms)
    for row, (fx, fy, c) in zip(terms, live):
        assert row[0] == fx and row[1] == fy and complex(row[2], row[3]) == c


def test_andrews_model_matches_oracle():
    m = pa.AndrewsModel(Cn2=3e-15, l0=4e-3, L0=50.0)
    kappa = np.geomspace(1e-3, 5e3, 64)
    assert np.array_equal(m.psd_n(kappa), orc.andrews_psd_n(kappa, 3e-15, 4e-3, 50.0))
    assert np.array_equal(m.psd_phi_f(kappa, 7.7e6, 300.0), orc.psd_phi_f(kappa, 3e-15, 4e-3, 50.0, 7.7e6, 300.0, orc.andrews_psd_n))


def test_wind_su_screen_state_and_routing():
    g = load_golden("windsu128")
    p = g["params"]
    screen = pa.WindSUPhaseScreen(pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"]), float(g["speed"]),
                                  model=pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"]))
    ch = _n4_channel(screen)
    ch.path.init_phase_screens()
    assert not ch.path._fusable()                    # carries its own translation state -> step-by-step path
    first = ch.path.phase_screens[0]
    np.random.seed(int(g["seed"]))
    sp = first._get_spectrum()
    assert np.array_equal(first.cnp, g["cnp0"]) and sp.value.dtype == np.complex128
    base = orc.logpolar_base(p["m"], p["f_min"], p["f_max"])
    np.random.seed(int(g["seed"]))
    rho, theta, value = orc.draw_wind_su_spectrum(base, p["Cn2"], p["l0"], p["L0"], p["wvl"], p["length"] / p["count"])
    assert np.array_equal(sp.rho, rho) and np.array_equal(sp.theta, theta) and np.array_equal(sp.value, value)
    assert [first._next_offset() for _ in range(3)] == [(0.0, 0), (float(g["speed"]), 0), (2 * float(g["speed"]), 0)]
    assert first._get_spectrum() is not sp and np.array_equal(first._get_spectrum().value, value)     # drawn once
    with pytest.raises(TypeError):
        first._screen_for_path(shift=(0, 0.1))


def test_header_cites_the_reference_for_every_compute_entry_point():
    """include/pyatm_b200.h: every exported function that replaces reference code carries a file:line citation of what it
    replaces in the comment block in front of it (utility entry points -- version, error text, counters, profiling -- excepted)."""
    hdr = open(os.path.join(ROOT, "include", "pyatm_b200.h")).read()
    utilities = {"pa_version", "pa_last_error", "pa_device_count", "pa_launch_count", "pa_ctx_destroy", "pa_ctx_permutation",
                 "pa_ctx_fft_geometry", "pa_fft_pass", "pa_phase_to_turns", "pa_simulate_batch_device", "pa_comm_unique_id",
                 "pa_comm_create", "pa_comm_destroy", "pa_stream_synchronize"}
    cite = re.compile(r"[\w/]+\.py:\d+")
    last_comment = ""
    checked = 0
    for m in re.finditer(r"/\*.*?\*/|PA_API[^;(]*?\b(pa_\w+)\s*\(", hdr, flags=re.S):
        if m.group(1) is None:
            last_comment = m.group(0)
        elif m.group(1) not in utilities:
            assert cite.search(last_comment), f"{m.group(1)}: no reference citation in the preceding comment"
            checked += 1
    assert checked >= 13


def test_c_abi_rejects_null_arguments_with_a_status_code():
    """Every entry point that takes a context validates its arguments before touching CUDA: a null context / null pointers
    give status PA_ERR_ARG (1) and a message through pa_last_error() -- no exception, no crash, no GPU needed."""
    import ctypes as C
    lib = nat.load()
    skip = {"pa_version", "pa_last_error", "pa_launch_count", "pa_device_count", "pa_ctx_create", "pa_ctx_destroy"}
    for name, (_, args) in nat.SIGNATURES.items():
        if name in skip:
            continue
        call = [0.0 if a is C.c_double else (None if a in (C.c_void_p, C.POINTER(nat.PaPath), C.POINTER(C.c_void_p)) else 0) for a in args]
        assert getattr(lib, name)(*call) == 1, name
        assert b"bad arguments" in lib.pa_last_error(), name
    assert lib.pa_ctx_destroy(None) == 0                      # destroying nothing is not an error
    h = C.c_void_p()
    assert lib.pa_ctx_create(C.byref(h), 0, 100, 0) != 0      # not a power of two in [64, 8192]
    assert b"unsupported" in lib.pa_last_error()
    with pytest.raises(nat.NativeError):
        nat.check(1)


# ---- round-2 host logic -----------------------------------------------------------------------------------------------
def test_pyatmosphere_import_name_is_a_true_alias():
    """README.md:26-28,100 / main.ipynb import lines resolve against this tree: the `pyatmosphere` package re-exports the
    modules of pyatmosphere_b200 under the reference's module paths (same objects, no second copy of any state)."""
    import importlib
    import pyatmosphere
    from pyatmosphere import gpu, QuickChannel, simulations, measures          # noqa: F401
    from pyatmosphere import (Channel, RectGrid, RandLogPolarGrid, GaussianSource, IdenticalPhaseScreensPath,    # noqa: F401
                              SSPhaseScreen, CirclePupil, MVKModel)
    from pyatmosphere.simulations import BeamResult, PDTResult, Simulation      # noqa: F401
    assert gpu.config is pa.gpu.config and simulations is pa.simulations
    for name in ("channels", "grids", "pathes", "phase_screens", "pupils", "sources", "utils", "measures", "gpu",
                 "simulations.beam", "simulations.pdt", "simulations.simulation", "simulations.result", "simulations.measure",
                 "simulations.si", "simulations.wind", "theory.models", "theory.sources", "theory.atmosphere"):
        assert importlib.import_module("pyatmosphere." + name) is importlib.import_module("pyatmosphere_b200." + name), name
    ns = {}
    exec("from pyatmosphere import *", ns)
    assert sorted(k for k in ns if not k.startswith("__")) == ["Channel", "QuickChannel"]      # __init__.py:14-17
    assert pyatmosphere.PlaneSource is pa.PlaneSource
    ch = QuickChannel(Cn2=1e-15, length=10000, count_ps=5, beam_w0=0.09, beam_wvl=8.08e-07, aperture_radius=0.12)
    assert ch.grid.resolution[0] == 1024 and len(ch.path.phase_screens) == 5


def test_resume_continues_the_device_rng_counter(tmp_path):
    """A Result resumed from its CSV checkpoint holds L records drawn with device-RNG indices 0..L-1: the Simulation must
    continue at index L (it restarted at 0 and stored duplicate samples)."""
    p = load_golden("turb128")["params"]
    ch = _channel(p)
    path = str(tmp_path / "beam.csv")
    beam = pa.simulations.BeamResult(ch, max_size=10, save_path=path)
    for i, m in enumerate(beam.measures):
        m.data = [0.1 * (i + 1)] * 4
    beam.save_output()
    again = pa.simulations.BeamResult(ch, max_size=10, save_path=path)
    pdt = pa.simulations.PDTResult(ch, max_size=10)
    sim = pa.simulations.Simulation([again, pdt])
    assert sim.realizations_done == 4
    assert pa.simulations.Simulation([pdt]).realizations_done == 0


def test_non_square_grid_is_rejected_before_any_native_call():
    from pyatmosphere_b200 import gpu
    assert gpu.config["use_gpu"]
    with pytest.raises(ValueError, match="square"):
        eng.grid_context(pa.RectGrid((128, 256), 1e-3))


def test_reference_style_screen_subclass_still_generates():
    """A user subclass written against the reference's contract (generate_phase_screen returns the complex screen and
    knows nothing about real_only, phase_screens.py:21-28) keeps working through PhaseScreen.generate."""
    class Ramp(pa.phase_screens.PhaseScreen):
        def generate_phase_screen(self, gain=1.0):
            return gain * (np.arange(6).reshape(2, 3) + 1j * np.ones((2, 3)))

    scr = Ramp(model=None)
    assert np.array_equal(scr.generate(), np.arange(6).reshape(2, 3))
    assert np.array_equal(scr.generate(gain=2.0), 2.0 * np.arange(6).reshape(2, 3))
    assert np.iscomplexobj(scr.generate(complex=True))
    gen = scr.generator()
    assert np.array_equal(next(gen), np.arange(6).reshape(2, 3)) and np.array_equal(next(gen), np.ones((2, 3)))


def test_uniform_ring_powers_detects_unequal_slabs():
    """The device RNG of pa_simulate_batch draws every screen from ONE ring-power table: only valid when all screens of
    the path carry the same powers."""
    p = load_golden("turb128")["params"]
    ch = _channel(p)
    ch.path.init_phase_screens()
    assert eng.uniform_ring_powers(ch.path.phase_screens)
    model = pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])
    fg = pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])
    thin, thick = pa.SSPhaseScreen(model=model, f_grid=fg, thickness=1e3), pa.SSPhaseScreen(model=model, f_grid=fg, thickness=3e3)
    ch2 = pa.Channel(grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]), source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
                     path=pa.PhaseScreensPath(length=4e3, phase_screens=[thin, thick], positions=[5e2, 2.5e3]),
                     pupil=pa.CirclePupil(radius=p["pupil"]))
    ch2.path.init_phase_screens()
    assert ch2.path._fusable() and not eng.uniform_ring_powers(ch2.path.phase_screens)
    assert np.allclose(thick._get_psd(), 3 * thin._get_psd(), rtol=1e-5)


def test_header_is_plain_c(tmp_path):
    """include/pyatm_b200.h is a C header (the boundary the reference's ctypes / cffi / cgo-style binding would consume): it
    compiles as C99 with no C++ or CUDA types, and a C translation unit can take the address of every declared function."""
    import shutil
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = open(os.path.join(ROOT, "include", "pyatm_b200.h")).read()
    names = sorted(set(re.findall(r"PA_API[^;(]*?\b(pa_\w+)\s*\(", hdr)))
    assert len(names) == len(nat.SIGNATURES) == 32
    src = tmp_path / "use.c"
    src.write_text('#include "pyatm_b200.h"\n' + "void* table[] = {" + ", ".join(f"(void*){n}" for n in names) + "};\n" +
                   "int main(void) { pa_path p; p.n_screens = 0; return (int)sizeof(table) * 0 + p.n_screens; }\n")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "use.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plain_c_host_example_links_against_the_library(tmp_path):
    """examples/c_host.c (a host with no Python, torch or CUDA headers) compiles as C99 and links against the shipped
    library; without a batch file it prints its usage and exits 2 (the GPU run is tests/test_gpu_c_host.py)."""
    import shutil
    import subprocess
    from pyatmosphere_b200 import _native as nat
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(os.path.abspath(nat.LIB_PATH))
    exe = str(tmp_path / "c_host")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "c_host.c"), "-o", exe, "-L", libdir, "-lpyatm_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 2 and "usage" in run.stderr


def test_every_symbol_of_the_scope_table_is_importable_under_the_reference_names():
    """SURVEY.md section 8(a) rows a1-a14 and (f) n1-n4, symbol by symbol, under the reference's module paths."""
    import importlib
    surface = {
        "pyatmosphere.grids": ["RectGrid.get_x", "RectGrid.get_y", "RectGrid.get_xy", "RectGrid.get_rho2", "RectGrid.get_f_grid",
                               "RectGrid.origin_index", "RectGrid.extent", "RectGrid._left_bound", "RectGrid._right_bound",
                               "RectGrid._top_bound", "RectGrid._bottom_bound", "RandLogPolarGrid.base", "RandLogPolarGrid.get_rho",
                               "RandLogPolarGrid.get_theta", "RandLogPolarGrid.get_xy"],
        "pyatmosphere.sources": ["GaussianSource.output", "PlaneSource"],
        "pyatmosphere.theory.sources": ["GaussianBeam.amplitude", "GaussianBeam.get_w", "GaussianBeam.get_theta0",
                                        "GaussianBeam.get_Lambda0", "GaussianBeam.get_theta", "GaussianBeam.get_Lambda"],
        "pyatmosphere.theory.vacuum": ["vacuum_propagation"],
        "pyatmosphere.theory.models": ["MVKModel.psd_n", "Model.psd_phi_f"],
        "pyatmosphere.utils": ["fft2", "ifft2", "CrossRef", "Default", "PolarDiscreteFunction"],
        "pyatmosphere.pathes": ["VacuumPath.lossless_output", "PhaseScreensPath.generator", "PhaseScreensPath.lossless_output",
                                "AbstractPath.output", "AbstractPath.append_losses", "IdenticalPhaseScreensPath"],
        "pyatmosphere.phase_screens": ["SSPhaseScreen.generate_phase_screen", "SSPhaseScreen.generate", "SSPhaseScreen._get_spectrum",
                                       "SSPhaseScreen._get_psd", "SSPhaseScreen.cache_clear", "SUPhaseScreen", "FFTPhaseScreen",
                                       "WindSUPhaseScreen"],
        "pyatmosphere.pupils": ["CirclePupil.get_pupil", "CirclePupil.output"],
        "pyatmosphere.measures": ["I", "eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2"],
        "pyatmosphere.channels": ["Channel.run", "Channel.generator", "Channel.get_rythov2", "QuickChannel"],
        "pyatmosphere.gpu": ["config", "get_xp", "get_array"],
        "pyatmosphere.simulations.simulation": ["Simulation.run", "Simulation.iter"],
        "pyatmosphere.simulations.measure": ["Measure"],
        "pyatmosphere.simulations.result": ["Result.save_output", "Result.load_output"],
        "pyatmosphere.simulations.beam": ["BeamResult"],
        "pyatmosphere.simulations.pdt": ["PDTResult", "TrackedPDTResult"],
        "pyatmosphere.simulations.si": ["SIResult"],
        "pyatmosphere.simulations.wind": ["TimeCoherenceResult", "TimeBWcorrSimulation"],
    }
    missing = []
    for module, names in surface.items():
        mod = importlib.import_module(module)
        for dotted in names:
            node = mod
            for part in dotted.split("."):
                if not hasattr(node, part):
                    missing.append(f"{module}:{dotted}")
                    break
                node = getattr(node, part)
    assert not missing, missing
    g = importlib.import_module("pyatmosphere.grids").RectGrid((5, 8), 0.5)
    assert (g._left_bound, g._right_bound, g._top_bound, g._bottom_bound) == (-2, 3, -4, 4)       # grids.py:36-50
    assert list(g.extent) == [-1.0, 1.5, -2.0, 2.0]


def test_bench_reference_arm_prints_exactly_one_json_line(tmp_path):
    """bench.py --impl reference (the CPU arm, runnable here): stdout carries ONE line and it parses; whatever else lands on
    file descriptor 1 during the run (NCCL prints its version there when NCCL_DEBUG is set) is moved to stderr."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-samples", "1"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, run.stdout[:500]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "realizations/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"

    import bench
    # the mechanism itself: a write to fd 1 after keep_stdout_for_the_json_line() does not reach the real stdout
    code = ("import os, bench; bench.keep_stdout_for_the_json_line(); os.write(1, b'NCCL version x\\n'); "
            "bench.emit({'ok': 1})")
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    assert run.stdout == '{"ok": 1}\n' and "NCCL version x" in run.stderr
