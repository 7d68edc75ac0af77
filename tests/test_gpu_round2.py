"""Round-2 GPU tests of the host routes around the hot path: the reference's README run verbatim through the
`pyatmosphere` import name, the block route of Simulation.run (device-side accumulation, one gather per block), resume of
a device-RNG run, and paths whose screens carry different ring powers."""
import sys
import types

import numpy as np
import pytest

from conftest import load_golden, rel_l2
from test_gpu_parity import build_channel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cfg():
    import pyatmosphere_b200 as pa
    saved = dict(pa.gpu.config)
    yield
    pa.gpu.config.clear()
    pa.gpu.config.update(saved)


class _Axis:
    def __init__(self, log):
        self._log = log

    def __getattr__(self, name):
        def call(*args, **kwargs):
            self._log.append((name, args, kwargs))
        return call


def _fake_matplotlib(monkeypatch):
    """matplotlib is not installed on the box: a recording stand-in (plotting itself is out of scope)."""
    log = []
    plt = types.ModuleType("matplotlib.pyplot")

    def subplots(rows=1, cols=1, squeeze=True, **kwargs):
        axes = np.empty((rows, cols), dtype=object)
        for i in range(rows):
            for j in range(cols):
                axes[i, j] = _Axis(log)
        log.append(("subplots", (rows, cols), kwargs))
        return object(), (axes[0, 0] if squeeze and rows * cols == 1 else axes)

    plt.subplots = subplots
    for name in ("imshow", "hist", "plot", "legend", "show", "xlabel", "ylabel", "title", "figure", "colorbar"):
        setattr(plt, name, getattr(_Axis(log), name))
    root = types.ModuleType("matplotlib")
    root.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", root)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    return log


README_GPU = """
from pyatmosphere import gpu
gpu.config['use_gpu'] = True
"""
README_QUICK = """
from pyatmosphere import QuickChannel

quick_channel = QuickChannel(
    Cn2=1e-15,
    length=10000,
    count_ps=5,
    beam_w0=0.09,
    beam_wvl=8.08e-07,
    aperture_radius=0.12
    )

quick_channel.plot()
"""
README_ADVANCED = """
import numpy as np
from pyatmosphere import (
    Channel,
    RectGrid,
    RandLogPolarGrid,
    GaussianSource,
    IdenticalPhaseScreensPath,
    SSPhaseScreen,
    CirclePupil,
    MVKModel,
    measures
    )

channel = Channel(
    grid=RectGrid(
        resolution=2048,
        delta=0.0015
    ),
    source=GaussianSource(
        wvl=808e-9,
        w0=0.12,
        F0=np.inf
    ),
    path=IdenticalPhaseScreensPath(
        phase_screen=SSPhaseScreen(
            model=MVKModel(
                Cn2=5e-16,
                l0=6e-3,
                L0=1e3,
            ),
            f_grid=RandLogPolarGrid(
                points=2**10,
                f_min=1 / 1e3 / 15,
                f_max=1 / 6e-3 * 2
            )
        ),
        length=50e3,
        count=5
    ),
    pupil=CirclePupil(
        radius=0.2
    ),
)

channel_output = channel.run(pupil=False)
intensity = measures.I(channel, output=channel_output)
mean_x = measures.mean_x(channel, output=channel_output)
"""
README_SIM = """
from pyatmosphere import simulations

beam_result = simulations.BeamResult(quick_channel, max_size=2000)
pdt_result = simulations.PDTResult(quick_channel, max_size=6000)
sim = simulations.Simulation([beam_result, pdt_result])
sim.run(plot_step=1000)
"""


def test_reference_readme_runs_verbatim(monkeypatch):
    """The four usage snippets of the reference's README.md (:26-29, :33-46, :50-97, :100-106), executed as written
    against the `pyatmosphere` import name of this tree.  Config 3's seed-3 records tie the advanced-channel snippet to
    the reference's own output; the simulation snippet reproduces the statistics the reference's notebook recorded."""
    log = _fake_matplotlib(monkeypatch)
    ns = {}
    exec(README_GPU, ns)
    exec(README_QUICK, ns)
    shown = [e for e in log if e[0] == "imshow"]
    assert len(shown) == 1 and shown[0][1][0].shape == (1024, 1024)
    assert np.array_equal(shown[0][2]["extent"], ns["quick_channel"].grid.extent)
    g = load_golden("c3_2048")
    np.random.seed(int(g["seeds"][0]))
    exec(README_ADVANCED, ns)
    assert ns["channel_output"].shape == (2048, 2048) and ns["intensity"].shape == (2048, 2048)
    ref = dict(zip([str(k) for k in g["ref_names"]], g["ref_measures"][0]))
    f64 = dict(zip([str(k) for k in g["f64_names"]], g["f64_measures"][0]))
    assert ns["mean_x"] == pytest.approx(f64["mean_x"], rel=1e-4, abs=1e-7)
    assert ns["mean_x"] == pytest.approx(ref["mean_x"], rel=2e-2, abs=2e-5)
    assert float(ns["intensity"].sum().item()) * 0.0015**2 == pytest.approx(1.0, abs=2e-5)
    np.random.seed(2023)
    exec(README_SIM, ns)
    beam, pdt = ns["beam_result"], ns["pdt_result"]
    assert len(beam.measures[0]) == 2000 and len(pdt.measures[0]) == 6000
    assert sum(1 for e in log if e[0] == "stairs") == 7          # plot_step=1000 -> 6 plots + the closing one
    for key, (ref_v, ref_err, half_digit) in {"bw": (3.9e-2, 6.2e-4, 0.05e-2), "lt": (1.8e-1, 6.8e-4, 0.05e-1),
                                              "st": (1.6e-1, 4.3e-4, 0.05e-1)}.items():      # main.ipynb:252-254
        val, err = getattr(beam, key)
        assert abs(val - ref_v) < 4 * np.hypot(err, ref_err) + half_digit, (key, val, err)
    assert pdt.histogram().sum() == 6000


def test_block_route_equals_batch_route_and_is_sync_free():
    """Simulation.run in device-RNG mode accumulates on the device and gathers once; its records equal the per-batch
    engine route (simulate_realizations) for the same global indices, for fixed and for tracked apertures."""
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact", theta_cut=2.0, rng="philox", seed=5, batch=4)
    p = load_golden("turb128")["params"]
    ch = build_channel(pa, p)
    beam = pa.simulations.BeamResult(ch, max_size=9)
    pdt = pa.simulations.PDTResult(ch, pupils=[pa.CirclePupil(0.1), pa.CirclePupil(0.03)], max_size=11)
    trk = pa.simulations.TrackedPDTResult(ch, pupils=[pa.CirclePupil(0.05)], max_size=11)
    sim = pa.simulations.Simulation([beam, pdt, trk])
    calls = []
    real = pa.distributed.all_gather_blocks
    pa.distributed.all_gather_blocks = lambda local, counts: calls.append(tuple(counts)) or real(local, counts)
    try:
        sim.run()
    finally:
        pa.distributed.all_gather_blocks = real
    assert calls == [(11,)]                                     # one gather for the whole run
    assert len(beam.measures[0]) == 9 and len(pdt.measures[1]) == 11 and sim.realizations_done == 11
    cols = eng.table_columns([0.03, 0.1], [0.05])
    want = eng.simulate_realizations(ch, 0, 11, np.arange(11), [0.03, 0.1], [0.05])
    assert np.array_equal(np.asarray(beam.measures[0].data), want[:9, cols["mean_x"]])
    assert np.array_equal(np.asarray(pdt.measures[0].data), want[:, cols[("fixed", 0.1)]])
    assert np.array_equal(np.asarray(pdt.measures[1].data), want[:, cols[("fixed", 0.03)]])
    assert np.array_equal(np.asarray(trk.measures[2].data), want[:, cols[("tracked", 0.05)]])
    # save_step cuts the run into blocks; the records do not change
    pdt2 = pa.simulations.PDTResult(ch, pupils=[pa.CirclePupil(0.1)], max_size=11)
    pa.simulations.Simulation([pdt2]).run(save_step=4)
    assert np.array_equal(np.asarray(pdt2.measures[0].data), want[:, cols[("fixed", 0.1)]])


def test_resumed_device_rng_run_draws_new_realizations(tmp_path):
    """resuming from a CSV checkpoint in 'philox' mode must not replay indices 0..L-1."""
    import pyatmosphere_b200 as pa
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact", theta_cut=2.0, rng="philox", seed=8, batch=4)
    p = load_golden("turb128")["params"]
    ch = build_channel(pa, p)
    full = pa.simulations.PDTResult(ch, max_size=10)
    pa.simulations.Simulation([full]).run()
    path = str(tmp_path / "pdt.csv")
    part = pa.simulations.PDTResult(ch, max_size=6, save_path=path)
    part.save_float_format = lambda v: "%.17e" % v                        # keep the checkpoint lossless for the comparison
    pa.simulations.Simulation([part]).run(save_step=3)
    again = pa.simulations.PDTResult(ch, max_size=10, save_path=path)
    assert len(again.measures[0]) == 6
    sim = pa.simulations.Simulation([again])
    assert sim.realizations_done == 6
    sim.run()
    data = np.asarray(again.measures[0].data)
    assert len(set(data.tolist())) == 10                                   # no duplicate samples
    want = np.asarray(full.measures[0].data)
    assert np.array_equal(data[6:], want[6:])                              # new draws == those of the uninterrupted run
    assert np.allclose(data[:6], want[:6], rtol=1e-13, atol=0)             # (pandas' CSV float parser is not round-trip exact)


def test_device_rng_uses_each_screens_own_ring_powers():
    """a PhaseScreensPath of unequal slabs in 'philox' mode -- every screen is drawn from its own table."""
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact", theta_cut=2.0, rng="philox", seed=3, batch=4)
    p = load_golden("turb128")["params"]
    model = pa.MVKModel(Cn2=p["Cn2"], l0=p["l0"], L0=p["L0"])
    fg = pa.RandLogPolarGrid(points=p["m"], f_min=p["f_min"], f_max=p["f_max"])
    thin, thick = pa.SSPhaseScreen(model=model, f_grid=fg, thickness=1e3), pa.SSPhaseScreen(model=model, f_grid=fg, thickness=3e3)
    ch = pa.Channel(grid=pa.RectGrid(resolution=p["n"], delta=p["delta"]), source=pa.GaussianSource(wvl=p["wvl"], w0=p["w0"], F0=np.inf),
                    path=pa.PhaseScreensPath(length=4e3, phase_screens=[thin, thick], positions=[5e2, 2.5e3]),
                    pupil=pa.CirclePupil(radius=p["pupil"]))
    pdt = pa.simulations.PDTResult(ch, max_size=4)
    sim = pa.simulations.Simulation([pdt])
    sim.run()
    assert not sim._runner.one_call
    # restated by hand: per-screen draws from per-screen tables, literal propagation, separate measure sweep
    ctx = eng.channel_context(ch)
    S, M, B = 2, p["m"], 4
    fx = torch.empty((S, B, M), dtype=torch.float32, device="cuda")
    fy, cf = torch.empty_like(fx), torch.empty((S, B, M, 2), dtype=torch.float32, device="cuda")
    for s, ps in enumerate((thin, thick)):
        e_d, p_d = eng.ring_tables(ctx, ps)
        nat.check(ctx.lib.pa_rng_spectrum(ctx.handle, 3, 0, B, s, 1, M, nat.ptr(e_d), nat.ptr(p_d), nat.ptr(fx[s]), nat.ptr(fy[s]),
                                          nat.ptr(cf[s]), nat.stream_ptr()))
    amp = (cf[..., 0] ** 2 + cf[..., 1] ** 2).mean(dim=1).cpu().numpy()       # E|c|^2 = 2 psd per ring
    assert np.median(amp[1] / amp[0]) == pytest.approx(3.0, rel=0.35)        # thick slab: three times the ring powers
    field = ctx.empty_field(B)
    desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
    nat.check(ctx.lib.pa_propagate(ctx.handle, desc.ref(), nat.ptr(field), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), nat.stream_ptr()))
    etas = [pa.measures.eta(ch, output=ch.pupil.output(pa.gpu.DeviceArray(field[i]))) for i in range(B)]
    assert np.allclose(pdt.measures[0].data, etas, rtol=1e-6)


def test_stats_allreduce_behind_the_c_abi_single_rank():
    """pa_comm_* / pa_stats_allreduce (NCCL bound with dlopen inside libpyatm_b200.so): a one-rank communicator reduces in
    place to the same values; either buffer may be absent.  (Two ranks: tests/_nccl_sim_worker.py on a two-GPU box.)"""
    import torch
    from pyatmosphere_b200.distributed import StatsComm
    comm = StatsComm(rank=0, world=1)
    hist = torch.arange(200, dtype=torch.int64, device="cuda")
    sums = torch.tensor([0.25, -3.0], dtype=torch.float64, device="cuda")
    comm.allreduce(hist, sums)
    comm.allreduce(hist, None)
    comm.allreduce(None, sums)
    torch.cuda.synchronize()
    assert torch.equal(hist.cpu(), torch.arange(200, dtype=torch.int64)) and sums.cpu().tolist() == [0.25, -3.0]
    assert len(comm.unique_id) == 128
    comm.close()


def test_plane_source_through_a_vacuum_path():
    """PlaneSource (sources.py:16-18; unusable with any path in the reference) is the field of ones here: a
    plane wave stays a plane wave in vacuum and only picks up the carrier phase e^{ikL}."""
    import pyatmosphere_b200 as pa
    pa.gpu.config.update(use_gpu=True, dtype="complex128")
    n, delta, wvl, length = 256, 2e-3, 808e-9, 1234.5
    ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.PlaneSource(wvl=wvl), path=pa.VacuumPath(length=length),
                    pupil=pa.CirclePupil(radius=0.1))
    out = ch.run(pupil=False).get()
    import math
    ang = 2 * math.pi / wvl * length                   # 9.6e9 rad: libm reduces it exactly (a float64 `% (2 pi)` would not)
    want = complex(math.cos(ang), math.sin(ang))
    assert out.shape == (n, n) and np.allclose(out, want, rtol=0, atol=1e-9)


def test_async_batches_equal_the_synchronous_call():
    """pa_simulate_batch_async: two batches enqueued back to back (pinned host buffers, separate result tables), one
    pa_stream_synchronize; records equal those of two synchronous pa_simulate_batch calls.  The batch (40) spans several
    internal chunks at this size."""
    import torch
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import _engine as eng, _native as nat
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact", theta_cut=2.0)
    p = dict(load_golden("turb128")["params"], n=2048, delta=2.5e-4)        # 2048^2: chunks of 32 realizations
    ch = build_channel(pa, p)
    ch.path.init_phase_screens()
    ctx = eng.channel_context(ch)
    desc = ch.path._descriptor((0, 0), through_output=False, from_field=False)
    S, M, B = p["count"], p["m"], 40
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    rng = np.random.default_rng(5)
    pup = torch.from_numpy(np.array([[np.float32(p["pupil"] ** 2), 0, 0]], dtype=np.float32)).pin_memory()
    sets = []
    for _ in range(2):
        np.random.seed(int(rng.integers(1 << 30)))
        fx, fy, cf = eng.draw_spectra_numpy(ch.path, B)
        sets.append([torch.from_numpy(np.ascontiguousarray(a.transpose(1, 0, 2))).pin_memory() for a in (fx, fy)] +
                    [torch.from_numpy(np.ascontiguousarray(cf.transpose(1, 0, 2)).view(np.float32).copy()).pin_memory()])
    sync_out = [np.zeros((B, stride)) for _ in range(2)]
    for k in range(2):
        fx, fy, cf = sets[k]
        nat.check(ctx.lib.pa_simulate_batch(ctx.handle, desc.ref(), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), 0, 0, None, None, nat.ptr(pup),
                                            1, nat.ptr(sync_out[k]), stride, nat.stream_ptr()))
    async_out = [torch.zeros((B, stride), dtype=torch.float64).pin_memory() for _ in range(2)]
    for k in range(2):
        fx, fy, cf = sets[k]
        nat.check(ctx.lib.pa_simulate_batch_async(ctx.handle, desc.ref(), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), 0, 0, None, None,
                                                  nat.ptr(pup), 1, nat.ptr(async_out[k]), stride, nat.stream_ptr()))
    nat.check(ctx.lib.pa_stream_synchronize(ctx.handle, nat.stream_ptr()))
    for k in range(2):
        assert np.array_equal(async_out[k].numpy(), sync_out[k])
        assert np.all(np.abs(sync_out[k][:, 0] - 1.0) < 1e-4)              # total power of every realization
    assert not np.array_equal(sync_out[0], sync_out[1])


def test_block_route_with_tiny_and_unequal_record_counts():
    """One realization, and results that want different numbers of records (README: BeamResult 2000, PDTResult 6000): every
    Measure stops at its own max_size, the records are prefixes of one another's run."""
    import pyatmosphere_b200 as pa
    pa.gpu.config.update(use_gpu=True, dtype="complex64", screen_method="exact", theta_cut=2.0, rng="philox", seed=21, batch=4)
    p = load_golden("turb128")["params"]
    ch = build_channel(pa, p)
    one = pa.simulations.PDTResult(ch, max_size=1)
    pa.simulations.Simulation([one]).run()
    assert len(one.measures[0]) == 1
    beam = pa.simulations.BeamResult(ch, max_size=3)
    pdt = pa.simulations.PDTResult(ch, max_size=7)
    sim = pa.simulations.Simulation([beam, pdt])
    sim.run()
    assert [len(m) for m in beam.measures] == [3] * 6 and len(pdt.measures[0]) == 7 and sim.realizations_done == 7
    assert pdt.measures[0].data[0] == one.measures[0].data[0]
    sim.run()                                           # nothing left to do: no new records, no error
    assert len(pdt.measures[0]) == 7


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [64, 256, 2048])
def test_centred_transform_pair_of_utils(dtype, n):
    """utils.fft2 / utils.ifft2 (utils.py:42-50) through pa_fft2c against numpy's definition in complex128, their round
    trip, and theory.vacuum.vacuum_propagation (theory/vacuum.py:5-7) against the oracle's leg."""
    import pyatmosphere_b200 as pa
    from pyatmosphere_b200 import utils
    from pyatmosphere_b200.theory.vacuum import vacuum_propagation
    from oracle import splitstep as orc
    saved = dict(pa.gpu.config)
    try:
        pa.gpu.config.update(use_gpu=True, dtype=dtype)
        rng = np.random.default_rng(n)
        delta = 1.5e-3
        df = 1.0 / (n * delta)
        u = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(dtype)
        tol = 2e-6 if dtype == "complex64" else 1e-12
        want_f = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(u.astype(np.complex128)))) * delta**2
        got_f = utils.fft2(u, delta)
        assert got_f.shape == (n, n) and got_f.dtype == np.dtype(dtype)
        assert rel_l2(got_f.get(), want_f) < tol
        want_i = np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(u.astype(np.complex128)))) * (n * df) ** 2
        assert rel_l2(utils.ifft2(u, df).get(), want_i) < tol
        assert rel_l2(utils.ifft2(got_f, df).get(), u) < 2 * tol                      # device array in, round trip
        batch = np.stack([u, 2j * u])
        assert rel_l2(utils.fft2(batch, delta).get()[1], 2j * want_f) < tol             # [batch][N][N]
        # the reference's free function for one leg
        wvl, length = 808e-9, 3.0e3
        x, y = orc.rect_xy(n, delta)
        beam = orc.gaussian_source(x, y, 0.02 * n * delta * 10, wvl, mode="f64") * np.exp(1j * rng.standard_normal((n, n)) * 0.1)
        want = orc.vacuum_leg(beam, length, wvl, delta, mode="f64")
        f2 = None
        got = vacuum_propagation(beam.astype(dtype), length, 2 * np.pi / wvl, delta, f2, df)
        assert rel_l2(got.get(), want) < (1e-5 if dtype == "complex64" else 1e-10)
        with pytest.raises(ValueError):
            vacuum_propagation(beam, length, 2 * np.pi / wvl, delta, f2, 2 * df)
    finally:
        pa.gpu.config.clear()
        pa.gpu.config.update(saved)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_gaussian_beam_amplitude_on_given_radii(dtype):
    """theory/sources.py:16-18 through pa_gaussian_amplitude; equals GaussianSource.output() on the channel grid's rho^2."""
    import pyatmosphere_b200 as pa
    from oracle import splitstep as orc
    saved = dict(pa.gpu.config)
    try:
        pa.gpu.config.update(use_gpu=True, dtype=dtype)
        n, delta, wvl, w0, F0 = 256, 2e-3, 808e-9, 0.03, 4.0e3
        ch = pa.Channel(grid=pa.RectGrid(n, delta), source=pa.GaussianSource(wvl=wvl, w0=w0, F0=F0),
                        path=pa.VacuumPath(length=1e3), pupil=pa.CirclePupil(radius=1.0))
        rho2 = ch.grid.get_rho2()
        got = ch.source.amplitude(rho2)
        assert got.shape == (n, n) and got.dtype == np.dtype(dtype)
        x, y = orc.rect_xy(n, delta)
        want = orc.gaussian_source(x, y, w0, wvl, F0, mode="f64")
        # rho^2 arrives in float32 as the reference forms it (grids.py:71-73): the curvature phase k rho^2 / (2 F0) ~ 1e2 rad
        # carries its rounding, 6e-8 * 1e2 rad, in complex64 and complex128 alike
        assert rel_l2(got.get(), want) < 2e-5
        assert rel_l2(got.get(), ch.source.output().get()) < 2e-5
        g64 = pa.sources.GaussianSource(wvl=wvl, w0=w0, F0=np.inf).amplitude((x**2 + y**2).astype(np.float64))   # float32 rho^2, promoted: the oracle's
        assert rel_l2(g64.get(), orc.gaussian_source(x, y, w0, wvl, np.inf, mode="f64")) < (2e-7 if dtype == "complex64" else 1e-14)
    finally:
        pa.gpu.config.clear()
        pa.gpu.config.update(saved)
