"""Host-side glue between the reference-shaped objects and the native library: context lookup, the split of
the ring sum into polynomial + contraction, path descriptors, and the batched Monte-Carlo driver."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native as nat
from . import gpu

SCREEN_METHODS = {"exact": nat.PA_SCREEN_EXACT, "tc": nat.PA_SCREEN_TC}
MAX_DEGREE = 63


TC_MIN_GRID = 1024          # 'auto' uses the tensor-core screens from this grid size on
TC_NODE_SPACING = 16        # columns between exact polynomial evaluations in the tensor-core epilogue (screen_tc.cu)


def screen_method(n: int) -> int:
    """Resolve gpu.config['screen_method'] for a grid of size n."""
    name = gpu.config["screen_method"]
    if name == "auto":
        name = "tc" if (gpu.precision() == 0 and n % 256 == 0 and n >= TC_MIN_GRID) else "exact"
    if name == "tc" and (gpu.precision() != 0 or n % 256 != 0):
        raise ValueError("screen_method 'tc' needs dtype complex64 and a grid size that is a multiple of 256")
    return SCREEN_METHODS[name]


def theta_cut(n: int) -> float:
    """Largest phase argument (rad, over the whole grid) of the rings that are summed as a polynomial.
    The tensor-core epilogue interpolates that polynomial between nodes TC_NODE_SPACING columns apart, which needs
    the phase of its fastest harmonic to advance by <= 0.16 rad per node: theta_cut <= 0.16 * n / (2 * spacing)."""
    tc = gpu.config["theta_cut"]
    if screen_method(n) == nat.PA_SCREEN_TC:
        limit = 0.16 * n / (2 * TC_NODE_SPACING)
        return min(10.0 if tc is None else float(tc), limit)
    return 2.0 if tc is None else float(tc)


def channel_context(channel) -> nat.Context:
    gpu.require_gpu()
    grid = channel.grid
    n = grid.resolution[0]
    if grid.resolution[0] != grid.resolution[1]:
        raise ValueError("only square grids are supported (the reference's get_f_grid / ifft2 assume it too)")
    return nat.context(n, grid.delta, grid.get_x(), grid.get_y(), gpu.precision())


def grid_context(grid) -> nat.Context:
    gpu.require_gpu()
    if grid.resolution[0] != grid.resolution[1]:
        raise ValueError("only square grids are supported (the reference's get_f_grid / ifft2 assume it too)")
    return nat.context(grid.resolution[0], grid.delta, grid.get_x(), grid.get_y(), gpu.precision())


def plan_low_rings(ring_edges, ring_psd, x_extent, y_extent, theta_cut, tol):
    """Choose (m_split, degree): rings [0, m_split) have phase arguments <= theta_cut everywhere on the grid
    and are summed as a Taylor polynomial of total degree `degree` whose truncation error is below `tol` rad.

    ring_edges: outer radii (monotone); ring_psd: phase variances.  |c_m| is bounded by 6 sqrt(psd_m)
    (Rayleigh tail 1.5e-8).  The bound on the argument of ring m over the grid is
    2 pi f_m sqrt(x_extent^2 + y_extent^2)."""
    edges = np.asarray(ring_edges, dtype=np.float64)
    psd = np.asarray(ring_psd, dtype=np.float64)
    if theta_cut is None or theta_cut <= 0 or np.any(np.diff(edges) < 0):
        return 0, -1
    tmax = 2 * np.pi * edges * math.hypot(x_extent, y_extent)
    m_split = int(np.searchsorted(tmax, theta_cut, side="right"))
    amp = 6 * np.sqrt(np.maximum(psd, 0))
    while m_split > 0:
        t = tmax[:m_split]
        a = amp[:m_split]
        for deg in range(1, MAX_DEGREE + 1):
            # log of sum_m a_m t_m^(deg+1) / (deg+1)!
            err = np.sum(a * np.exp((deg + 1) * np.log(np.maximum(t, 1e-300)) - math.lgamma(deg + 2)))
            if err <= tol:
                return m_split, deg
        m_split -= max(1, m_split // 16)
    return 0, -1


class PathDescriptor:
    """Owns the ctypes pa_path and the host arrays it points to."""

    def __init__(self, legs, screen_scale, final_scale, wvl, w0, F0, m, m_split, degree, shift, method, from_field,
                 coef_bound=0.0):
        self.legs = np.ascontiguousarray(legs, dtype=np.float64)
        self.scales = np.ascontiguousarray(screen_scale if len(screen_scale) else [1.0], dtype=np.float64)
        n_screens = len(self.legs) - 1
        self.c = nat.PaPath(
            n_screens=n_screens,
            leg_lengths_host=self.legs.ctypes.data_as(C.POINTER(C.c_double)),
            screen_scale_host=self.scales.ctypes.data_as(C.POINTER(C.c_double)),
            final_scale=float(final_scale), wvl=float(wvl), w0=float(w0), F0=float(F0), m=int(m), m_split=int(m_split),
            degree=int(degree), shift_x=float(shift[0]), shift_y=float(shift[1]), screen_method=int(method),
            coef_bound=float(coef_bound), from_field=int(bool(from_field)))

    def ref(self):
        return C.byref(self.c)


def loss_amplitude(db):
    return 10 ** (-db / 20) if db else 1.0


def path_losses(path, legs):
    """Amplitude factors exactly as the reference applies dB losses (pathes.py:19-24,71-73): after screen i the
    share losses*leg_i/length -- a share of exactly 0 falls back to the FULL loss because of
    `losses_db or self.losses_db` -- and nothing for the closing leg."""
    scales = []
    for i in range(len(legs) - 1):
        share = path.losses_db * legs[i] / path.length
        scales.append(loss_amplitude(share or path.losses_db))
    return scales


def screen_tolerance():
    return 1e-7 if gpu.precision() == 0 else 1e-12


def coef_bound(ring_psd, m_split):
    """Upper bound of |c_m| over the rings that go through the contraction: |n0 + i n1| < 6.5 has probability
    1 - 7e-10 per ring (Rayleigh), and c_m = (n0 + i n1) sqrt(psd_m)."""
    hi = np.asarray(ring_psd, dtype=np.float64)[m_split:]
    return float(6.5 * np.sqrt(hi.max())) if hi.size else 1.0


# ---- batched Monte-Carlo driver -----------------------------------------------------------------------------
def table_columns(pupils_fixed, pupils_tracked):
    """Column of every record in the table returned by simulate_realizations."""
    cols = {name: i for i, name in enumerate(nat.MEASURE_NAMES)}
    k = len(nat.MEASURE_NAMES)
    for r in pupils_fixed:
        cols[("fixed", r)] = k
        k += 1
    for r in pupils_tracked:
        cols[("tracked", r)] = k
        k += 1
    return cols


def draw_spectra_numpy(path, count):
    """`count` realizations x S screens drawn from numpy's global RNG in the reference's order (realization-major,
    then screen; per screen random(1), random(M), normal(2,M)).  Returns fx, fy [count][S][M] float32 and
    coef [count][S][M] complex64.

    Paths made of plain SSPhaseScreens over one frequency grid take the raw numbers screen by screen (three RNG calls
    each, nothing else inside the loop) and turn them into (fx, fy, coef) with whole-array operations afterwards -- the
    same elementwise float32 / complex64 arithmetic as grids.py:98-119 and phase_screens.py:98-103, so the results are
    bit-identical to the per-screen route below (tests/test_host_logic.py) at a fifth of the host time."""
    screens = path.phase_screens
    S, M = len(screens), screens[0].f_grid.points
    from .phase_screens import SSPhaseScreen
    if all(type(ps) is SSPhaseScreen and ps.f_grid is screens[0].f_grid for ps in screens):
        shared = np.empty((count, S, 1), dtype=np.float32)
        angle = np.empty((count, S, M), dtype=np.float32)
        normal = np.empty((count, S, 2, M), dtype=np.float64)
        rnd, nrm = np.random.random, np.random.normal
        for r in range(count):
            for s in range(S):
                shared[r, s] = rnd(size=(1,))
                angle[r, s] = rnd(size=(M,))
                normal[r, s] = nrm(size=(2, M))
        for ps in screens:
            ps.cache_clear()
        outer = screens[0].f_grid.base                                   # float32
        inner = np.insert(outer, 0, 0)[:-1]
        rho = np.sqrt(inner**2 + shared * (outer**2 - inner**2))         # grids.py:98-102, float32 throughout
        theta = 2 * np.pi * angle                                        # grids.py:104-107
        fx = rho * np.cos(theta)
        fy = rho * np.sin(theta)
        amp = np.stack([np.sqrt(ps._get_psd()) for ps in screens])       # [S][M] float32
        cf = (normal[:, :, 0] + 1j * normal[:, :, 1]).astype(np.complex64) * amp      # phase_screens.py:101-102
        return fx, fy, cf
    fx = np.empty((count, S, M), dtype=np.float32)
    fy = np.empty((count, S, M), dtype=np.float32)
    cf = np.empty((count, S, M), dtype=np.complex64)
    for r in range(count):
        for s, ps in enumerate(screens):
            ps.cache_clear()
            sp = ps._get_spectrum(use_cached_spectrum=False)
            fx[r, s] = ps.f_grid.get_x(sp.rho, sp.theta)
            fy[r, s] = ps.f_grid.get_y(sp.rho, sp.theta)
            cf[r, s] = sp.value
    return fx, fy, cf


def _buffers(ctx, batch):
    torch = nat.torch_mod()
    buf = getattr(ctx, "_mc_buffers", None)
    if buf is None or buf["field"].shape[0] < batch:
        buf = {"field": ctx.empty_field(batch),
               "table": torch.empty((batch, nat.MEASURE_HEAD + nat.MAX_PUPILS), dtype=torch.float64, device=ctx.tdevice),
               "table2": torch.empty((batch, nat.MEASURE_HEAD + nat.MAX_PUPILS), dtype=torch.float64, device=ctx.tdevice)}
        ctx._mc_buffers = buf
    return buf


def uniform_ring_powers(screens) -> bool:
    """True when every screen of the path has the same ring powers (IdenticalPhaseScreensPath, or a PhaseScreensPath of
    equal slabs): one table then serves the whole-path device RNG of pa_simulate_batch."""
    first = screens[0]._get_psd()
    return all(q._get_psd() is first or np.array_equal(q._get_psd(), first) for q in screens[1:])


def ring_tables(ctx, screen):
    """Device copies of the ring edges and ring powers of a screen (inputs of the device RNG)."""
    torch = nat.torch_mod()
    key = id(screen.f_grid), screen._get_psd().ctypes.data
    cache = getattr(ctx, "_ring_tables", {})
    if key not in cache:
        cache[key] = (torch.as_tensor(np.ascontiguousarray(screen.f_grid.base, dtype=np.float32), device=ctx.tdevice),
                      torch.as_tensor(np.ascontiguousarray(screen._get_psd(), dtype=np.float32), device=ctx.tdevice))
        ctx._ring_tables = cache
    return cache[key]


def simulate_realizations(channel, first, count, mine, pupils_fixed, pupils_tracked):
    """Propagate the realizations with global indices `mine` (a contiguous block of [first, first+count)) and
    reduce them.  Returns a float64 table [len(mine)][len(table_columns(...))]."""
    ctx = channel_context(channel)
    torch = nat.torch_mod()
    path = channel.path
    path.init_phase_screens()
    S, M = len(path.phase_screens), path.phase_screens[0].f_grid.points
    if len(pupils_fixed) > nat.MAX_PUPILS or len(pupils_tracked) > nat.MAX_PUPILS:
        raise ValueError(f"at most {nat.MAX_PUPILS} fixed and {nat.MAX_PUPILS} tracked apertures per simulation")
    cols = table_columns(pupils_fixed, pupils_tracked)
    # numpy mode: every rank draws the whole batch so that the global RNG stream stays identical on all ranks
    device_rng = gpu.config["rng"] != "numpy"
    if device_rng and not all(getattr(ps, "device_rng", False) for ps in path.phase_screens):
        raise ValueError("gpu.config['rng'] = 'philox' draws sparse-spectrum coefficients with fixed ring powers "
                         "(SSPhaseScreen); use rng = 'numpy' for the other screen generators")
    host = None if device_rng else draw_spectra_numpy(path, count)
    B = len(mine)
    if B == 0:
        return np.zeros((0, len(cols)), dtype=np.float64)
    dev = ctx.tdevice
    buf = _buffers(ctx, B)
    field, table, table2 = buf["field"], buf["table"], buf["table2"]
    desc = path._descriptor((0, 0), through_output=False, from_field=False)
    stream = nat.stream_ptr()
    stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
    nm = len(nat.MEASURE_NAMES)
    tab = np.array([[np.float32(r**2), 0, 0] for r in pupils_fixed], dtype=np.float32).reshape(-1, 3)
    # the one-call route draws every screen from ONE ring-power table; screens of different thickness / model are drawn
    # screen by screen below, each from its own table
    one_table = host is not None or uniform_ring_powers(path.phase_screens)
    if not pupils_tracked and len(pupils_fixed) <= 4 and one_table:
        # statistics only: one C-ABI call per batch; the library reduces inside the final row pass and never
        # writes the output fields (simulations/simulation.py:89-114 for Beam/PDT records)
        out_host = np.empty((B, stride), dtype=np.float64)
        if host is not None:
            sel = slice(int(mine[0] - first), int(mine[0] - first) + B)
            hfx = np.ascontiguousarray(host[0][sel].transpose(1, 0, 2))                               # [S][B][M]
            hfy = np.ascontiguousarray(host[1][sel].transpose(1, 0, 2))
            hcf = np.ascontiguousarray(host[2][sel].transpose(1, 0, 2)).view(np.float32)
            nat.check(ctx.lib.pa_simulate_batch(ctx.handle, desc.ref(), B, nat.ptr(hfx), nat.ptr(hfy), nat.ptr(hcf), 0, 0, None, None,
                                                nat.ptr(tab) if len(tab) else None, len(pupils_fixed), nat.ptr(out_host), stride, stream))
        else:
            edges_d, psd_d = ring_tables(ctx, path.phase_screens[0])
            nat.check(ctx.lib.pa_simulate_batch(ctx.handle, desc.ref(), B, None, None, None, int(gpu.config["seed"]), int(mine[0]),
                                                nat.ptr(edges_d), nat.ptr(psd_d), nat.ptr(tab) if len(tab) else None,
                                                len(pupils_fixed), nat.ptr(out_host), stride, stream))
        out = np.empty((B, len(cols)), dtype=np.float64)
        out[:, :nm] = out_host[:, :nm]
        out[:, nm:] = out_host[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + len(pupils_fixed)]
        return out
    if host is not None:
        sel = slice(int(mine[0] - first), int(mine[0] - first) + B)
        fx_d = torch.as_tensor(np.ascontiguousarray(host[0][sel].transpose(1, 0, 2)), device=dev)     # [S][B][M]
        fy_d = torch.as_tensor(np.ascontiguousarray(host[1][sel].transpose(1, 0, 2)), device=dev)
        cf_d = torch.as_tensor(np.ascontiguousarray(host[2][sel].transpose(1, 0, 2)).view(np.float32), device=dev)
    else:
        fx_d = torch.empty((S, B, M), dtype=torch.float32, device=dev)
        fy_d = torch.empty((S, B, M), dtype=torch.float32, device=dev)
        cf_d = torch.empty((S, B, M, 2), dtype=torch.float32, device=dev)
        for s, ps in enumerate(path.phase_screens):       # the records are keyed (seed; realization, screen, ring)
            edges_d, psd_d = ring_tables(ctx, ps)
            nat.check(ctx.lib.pa_rng_spectrum(ctx.handle, int(gpu.config["seed"]), int(mine[0]), B, s, 1, M, nat.ptr(edges_d),
                                              nat.ptr(psd_d), nat.ptr(fx_d[s]), nat.ptr(fy_d[s]), nat.ptr(cf_d[s]), stream))
    nat.check(ctx.lib.pa_propagate(ctx.handle, desc.ref(), nat.ptr(field), B, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(cf_d), stream))
    tab_d = torch.as_tensor(tab, device=dev) if len(pupils_fixed) else None
    nat.check(ctx.lib.pa_measure(ctx.handle, nat.ptr(field), B, nat.ptr(tab_d), len(pupils_fixed), 0, nat.ptr(table), stride, stream))
    out = np.empty((B, len(cols)), dtype=np.float64)
    t1 = table[:B].cpu().numpy()
    out[:, :nm] = t1[:, :nm]
    out[:, nm:nm + len(pupils_fixed)] = t1[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + len(pupils_fixed)]
    if pupils_tracked:
        # aperture re-centred on each realization's centroid: shift = (mean_x, mean_y) (simulations/pdt.py:62-66)
        per = np.empty((B, len(pupils_tracked), 3), dtype=np.float32)
        for j, r in enumerate(pupils_tracked):
            per[:, j, 0] = np.float32(r**2)
            per[:, j, 1] = t1[:, 1].astype(np.float32)
            per[:, j, 2] = t1[:, 2].astype(np.float32)
        per_d = torch.as_tensor(per, device=dev)
        nat.check(ctx.lib.pa_measure(ctx.handle, nat.ptr(field), B, nat.ptr(per_d), len(pupils_tracked), 1, nat.ptr(table2), stride, stream))
        t2 = table2[:B].cpu().numpy()
        out[:, nm + len(pupils_fixed):] = t2[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + len(pupils_tracked)]
    return out


class HostBatch:
    """One batch of realizations with HOST-drawn coefficients (numpy RNG mode), enqueued and not waited for: the spectra are
    drawn from numpy's global RNG in the reference's order when the object is made, copied to pinned memory and handed to
    pa_simulate_batch_async; `result()` waits for the batch and returns its rows.  Simulation.run keeps one HostBatch in
    flight, so the (RNG-bound) drawing of batch k+1 overlaps the GPU work of batch k.  Statistics-only route: fixed
    apertures (at most nat.MAX_PUPILS), no tracked ones."""

    def __init__(self, channel, first, count, mine, pupils_fixed):
        torch = nat.torch_mod()
        self.count, self.mine, self.first = int(count), np.asarray(mine), int(first)
        self.cols = table_columns(pupils_fixed, [])
        self.nf = len(pupils_fixed)
        path = channel.path
        path.init_phase_screens()
        host = draw_spectra_numpy(path, count)            # every rank draws the whole batch: identical RNG streams
        self.B = B = len(self.mine)
        self.event = None
        if B == 0:
            return
        ctx = channel_context(channel)
        sel = slice(int(self.mine[0] - first), int(self.mine[0] - first) + B)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()      # noqa: E731
        self.fx = pin(host[0][sel].transpose(1, 0, 2))                                # [S][B][M]
        self.fy = pin(host[1][sel].transpose(1, 0, 2))
        self.cf = pin(np.ascontiguousarray(host[2][sel].transpose(1, 0, 2)).view(np.float32))
        tab = np.array([[np.float32(r**2), 0, 0] for r in pupils_fixed], dtype=np.float32).reshape(-1, 3)
        self.tab = pin(tab) if len(tab) else None
        self.stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
        self.out = torch.empty((B, self.stride), dtype=torch.float64).pin_memory()
        self.desc = path._descriptor((0, 0), through_output=False, from_field=False)
        nat.check(ctx.lib.pa_simulate_batch_async(ctx.handle, self.desc.ref(), B, nat.ptr(self.fx), nat.ptr(self.fy), nat.ptr(self.cf), 0, 0,
                                                  None, None, nat.ptr(self.tab), self.nf, nat.ptr(self.out), self.stride, nat.stream_ptr()))
        self.event = torch.cuda.Event()
        self.event.record()

    def result(self):
        """Rows of this rank's share, [len(mine)][len(cols)] float64 (waits for the batch)."""
        out = np.zeros((self.B, len(self.cols)), dtype=np.float64)
        if self.B == 0:
            return out
        self.event.synchronize()
        raw = self.out.numpy()
        nm = len(nat.MEASURE_NAMES)
        out[:, :nm] = raw[:, :nm]
        out[:, nm:] = raw[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + self.nf]
        return out


class BlockRunner:
    """Device-RNG Monte-Carlo blocks with NOTHING but kernel launches between the first and the last realization: every
    batch is enqueued on the current stream (pa_simulate_batch_device, or pa_rng_spectrum + pa_propagate + pa_measure when
    apertures are tracked / the screens differ), its records land in a device table, and the caller reads that table
    back ONCE per block (simulations/simulation.py: one gather per save_step / at the end).  Mirrors the per-iteration
    work of simulations/simulation.py:89-114 for BeamResult / PDTResult / TrackedPDTResult records."""

    def __init__(self, channel, pupils_fixed, pupils_tracked):
        torch = nat.torch_mod()
        self.channel, self.path = channel, channel.path
        self.ctx = channel_context(channel)
        self.path.init_phase_screens()
        screens = self.path.phase_screens
        if not all(getattr(ps, "device_rng", False) for ps in screens):
            raise ValueError("gpu.config['rng'] = 'philox' draws sparse-spectrum coefficients with fixed ring powers "
                             "(SSPhaseScreen); use rng = 'numpy' for the other screen generators")
        if len(pupils_fixed) > nat.MAX_PUPILS or len(pupils_tracked) > nat.MAX_PUPILS:
            raise ValueError(f"at most {nat.MAX_PUPILS} fixed and {nat.MAX_PUPILS} tracked apertures per simulation")
        self.fixed, self.tracked = list(pupils_fixed), list(pupils_tracked)
        self.cols = table_columns(self.fixed, self.tracked)
        self.S, self.M = len(screens), screens[0].f_grid.points
        self.desc = self.path._descriptor((0, 0), through_output=False, from_field=False)
        self.one_call = not self.tracked and uniform_ring_powers(screens)
        dev = self.ctx.tdevice
        tab = np.array([[np.float32(r**2), 0, 0] for r in self.fixed], dtype=np.float32).reshape(-1, 3)
        self.pup_d = torch.as_tensor(tab, device=dev) if len(self.fixed) else None
        self.r2_tracked = torch.as_tensor(np.array([np.float32(r**2) for r in self.tracked], dtype=np.float32), device=dev)
        self.stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
        self.ws = None

    def _workspace(self, B):
        torch = nat.torch_mod()
        if self.ws is None or self.ws["B"] < B:
            dev, S, M = self.ctx.tdevice, self.S, self.M
            self.ws = {"B": B, "field": self.ctx.empty_field(B),
                       "fx": torch.empty((S, B, M), dtype=torch.float32, device=dev),
                       "fy": torch.empty((S, B, M), dtype=torch.float32, device=dev),
                       "cf": torch.empty((S, B, M, 2), dtype=torch.float32, device=dev),
                       "per": torch.empty((B, max(1, len(self.tracked)), 3), dtype=torch.float32, device=dev)}
        return self.ws

    def run(self, first: int, count: int):
        """Enqueue realizations [first, first + count); returns the device table [count][len(cols)] (float64).  No
        synchronisation, no host copies."""
        torch = nat.torch_mod()
        ctx, lib, h = self.ctx, self.ctx.lib, self.ctx.handle
        dev, stride = ctx.tdevice, self.stride
        # realizations per C call: the library works through a call in chunks of its own choosing (pa_simulate_batch*), so
        # calls are made large; gpu.config['batch'] is only a lower bound here
        seed, B0 = int(gpu.config["seed"]), max(256 if self.one_call else 1, int(gpu.config["batch"]))
        nm, nf, nt = len(nat.MEASURE_NAMES), len(self.fixed), len(self.tracked)
        raw = torch.empty((count, stride), dtype=torch.float64, device=dev)
        raw2 = torch.empty((count, stride), dtype=torch.float64, device=dev) if nt else None
        stream = nat.stream_ptr()
        screens = self.path.phase_screens
        for b0 in range(0, count, B0):
            B = min(B0, count - b0)
            if self.one_call:
                edges_d, psd_d = ring_tables(ctx, screens[0])
                nat.check(lib.pa_simulate_batch_device(h, self.desc.ref(), B, seed, first + b0, nat.ptr(edges_d), nat.ptr(psd_d),
                                                       nat.ptr(self.pup_d), nf, nat.ptr(raw[b0:]), stride, stream))
                continue
            ws = self._workspace(B0)
            fx, fy, cf = ws["fx"][:, :B].contiguous(), ws["fy"][:, :B].contiguous(), ws["cf"][:, :B].contiguous()
            for s, ps in enumerate(screens):                  # records are keyed (seed; realization, screen, ring)
                edges_d, psd_d = ring_tables(ctx, ps)
                nat.check(lib.pa_rng_spectrum(h, seed, first + b0, B, s, 1, self.M, nat.ptr(edges_d), nat.ptr(psd_d),
                                              nat.ptr(fx[s]), nat.ptr(fy[s]), nat.ptr(cf[s]), stream))
            field = ws["field"]
            nat.check(lib.pa_propagate(h, self.desc.ref(), nat.ptr(field), B, nat.ptr(fx), nat.ptr(fy), nat.ptr(cf), stream))
            nat.check(lib.pa_measure(h, nat.ptr(field), B, nat.ptr(self.pup_d), nf, 0, nat.ptr(raw[b0:]), stride, stream))
            if nt:
                # aperture re-centred on each realization's centroid: shift = (mean_x, mean_y) (simulations/pdt.py:62-66)
                per = ws["per"][:B]
                per[:, :, 0] = self.r2_tracked
                per[:, :, 1] = raw[b0:b0 + B, 1].to(torch.float32)[:, None]
                per[:, :, 2] = raw[b0:b0 + B, 2].to(torch.float32)[:, None]
                nat.check(lib.pa_measure(h, nat.ptr(field), B, nat.ptr(per), nt, 1, nat.ptr(raw2[b0:]), stride, stream))
        parts = [raw[:, :nm], raw[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + nf]]
        if nt:
            parts.append(raw2[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + nt])
        return torch.cat(parts, dim=1)
