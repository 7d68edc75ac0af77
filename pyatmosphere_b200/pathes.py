"""Propagation paths.  Mirror of /root/reference/pyatmosphere/pathes.py:10-95; all field arithmetic runs in
libpyatm_b200.so.

`output()` / `lossless_output()` use the fused propagator (pa_propagate: the trailing IFFT_x of a leg, the screen
multiply and the leading FFT_x of the next leg are one kernel); `generator()` -- which must hand the field and
the screen back after every slab -- uses the per-step building blocks (pa_vacuum_leg, pa_screen_ss,
pa_apply_screen) and is numerically the same path."""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod

import numpy as np

from . import _engine as eng
from . import _native as nat
from . import gpu
from .gpu import DeviceArray
from .sources import SourceField


def _as_field(ctx, input):
    """torch complex tensor [1][N][N] holding a private copy of `input` (paths work in place)."""
    torch = nat.torch_mod()
    if isinstance(input, DeviceArray):
        t = input.t
    else:
        t = torch.as_tensor(np.asarray(input), device=ctx.tdevice)
    t = t.to(ctx.cdtype)
    return t.reshape(1, ctx.n, ctx.n).clone().contiguous()


class AbstractPath(ABC):
    def __init__(self, length, losses_db=0):
        self.length = length
        self.losses_db = losses_db

    @abstractmethod
    def lossless_output(self, input, *args, **kwargs):
        pass

    def append_losses(self, input, losses_db=None):
        losses_db = losses_db or self.losses_db          # a share of exactly 0 falls back to the full loss
        return input * 10 ** (-losses_db / 20) if losses_db else input

    def output(self, input, *args, **kwargs):
        return self.append_losses(self.lossless_output(input, *args, **kwargs))


class VacuumPath(AbstractPath):
    """Free-space leg: IFFT2c(e^{ikL} e^{-i pi L lambda f^2} FFT2c(u)) (pathes.py:27-40, theory/vacuum.py:5-7)."""

    def lossless_output(self, input, length=None):
        length = length if length is not None else self.length
        if not length > 0:
            return input
        ctx = eng.channel_context(self.channel)
        field = _as_field(ctx, input)
        nat.check(ctx.lib.pa_vacuum_leg(ctx.handle, nat.ptr(field), 1, float(length), float(self.channel.source.wvl),
                                        nat.stream_ptr()))
        return DeviceArray(field[0])


class PhaseScreensPath(AbstractPath):
    def __init__(self, length, phase_screens, positions, losses_db=0):
        self.positions = positions
        self.phase_screens = phase_screens
        super().__init__(length, losses_db)

    def init_phase_screens(self):
        for phase_screen in self.phase_screens:
            phase_screen.channel = self.channel

    # ---- geometry ----------------------------------------------------------------------------------------------
    def leg_lengths(self):
        """Leg in front of every screen plus the closing leg (pathes.py:68-69,75)."""
        pos = self.positions
        legs = [pos[i] - pos[i - 1] if i > 0 else pos[0] for i in range(len(pos))]
        legs.append(self.length - pos[-1])
        return legs

    # ---- fused path ------------------------------------------------------------------------------------------
    def _descriptor(self, shift, through_output, from_field):
        legs = self.leg_lengths()
        scales = eng.path_losses(self, legs)
        final = eng.loss_amplitude(self.losses_db) if through_output else 1.0
        ps = self.phase_screens[0]
        src = self.channel.source
        plans = [q.low_ring_plan(shift) for q in self.phase_screens]
        m_split, degree = min(p[0] for p in plans), max(p[1] for p in plans)
        if m_split == 0:
            degree = -1
        return eng.PathDescriptor(legs, scales, final, src.wvl, getattr(src, "w0", 1.0), getattr(src, "F0", np.inf),
                                  ps.f_grid.points, m_split, degree, shift,
                                  eng.screen_method(self.channel.grid.resolution[0]), from_field,
                                  coef_bound=max(eng.coef_bound(q._ring_power(), m_split) for q in self.phase_screens))

    def _draw_spectra(self, wind):
        """Spectra of all screens in path order -- the order in which the reference's generator draws them."""
        return [ps._get_spectrum(use_cached_spectrum=wind) for ps in self.phase_screens]

    def _run_fused(self, input, shift=(0, 0), wind=False, through_output=False):
        ctx = eng.channel_context(self.channel)
        torch = nat.torch_mod()
        self.init_phase_screens()
        virtual = isinstance(input, SourceField) and input.is_virtual
        field = ctx.empty_field(1) if virtual else _as_field(ctx, input)
        spectra = self._draw_spectra(wind)
        fg = self.phase_screens[0].f_grid
        fx = np.stack([fg.get_x(s.rho, s.theta) for s in spectra]).astype(np.float32)
        fy = np.stack([fg.get_y(s.rho, s.theta) for s in spectra]).astype(np.float32)
        cf = np.stack([np.asarray(s.value, dtype=np.complex64) for s in spectra])
        dev = ctx.tdevice
        fx_d = torch.as_tensor(fx, device=dev)
        fy_d = torch.as_tensor(fy, device=dev)
        cf_d = torch.as_tensor(cf.view(np.float32), device=dev)
        desc = self._descriptor(shift, through_output, from_field=not virtual)
        nat.check(ctx.lib.pa_propagate(ctx.handle, desc.ref(), nat.ptr(field), 1, nat.ptr(fx_d), nat.ptr(fy_d),
                                       nat.ptr(cf_d), nat.stream_ptr()))
        return DeviceArray(field[0])

    def _fusable(self):
        """The fused propagator synthesises every screen itself from (fx, fy, c): it serves paths whose screens are all
        sums of harmonics over one shared log-polar grid (SSPhaseScreen, SUPhaseScreen)."""
        from .phase_screens import HarmonicSumScreen
        ps = self.phase_screens
        return len(ps) > 0 and all(isinstance(p, HarmonicSumScreen) and p.fusable and p.f_grid is ps[0].f_grid for p in ps)

    def lossless_output(self, input, *args, **kwargs):
        if self._fusable() and not args:
            return self._run_fused(input, through_output=False, **kwargs)
        generator = self.generator(input, *args, **kwargs)
        try:
            while True:
                next(generator)
        except StopIteration as stop:
            return stop.value

    def output(self, input, *args, **kwargs):
        if self._fusable() and not args:
            return self._run_fused(input, through_output=True, **kwargs)
        return self.append_losses(self.lossless_output(input, *args, **kwargs))

    # ---- step-by-step path (pathes.py:61-75) ------------------------------------------------------------------
    def generator(self, input, *args, **kwargs):
        ctx = eng.channel_context(self.channel)
        lib, h = ctx.lib, ctx.handle
        self.init_phase_screens()
        wvl = float(self.channel.source.wvl)
        field = _as_field(ctx, input)
        legs = self.leg_lengths()
        scales = eng.path_losses(self, legs)
        for i, phase_screen in enumerate(self.phase_screens):
            nat.check(lib.pa_vacuum_leg(h, nat.ptr(field), 1, float(legs[i]), wvl, nat.stream_ptr()))
            if hasattr(phase_screen, "_screen_for_path"):
                turns, phi = phase_screen._screen_for_path(*args, **kwargs)
            else:
                # any other generator (FFTPhaseScreen): the real phase as the screen returns it, reduced to turns
                phi = phase_screen.generate(*args, **kwargs).t.to(ctx.rdtype).contiguous()
                turns = nat.torch_mod().empty_like(phi)
                nat.check(lib.pa_phase_to_turns(h, nat.ptr(phi), 1 if ctx.precision == nat.PA_C128 else 0, nat.ptr(turns),
                                                phi.numel(), nat.stream_ptr()))
            nat.check(lib.pa_apply_screen(h, nat.ptr(field), 1, nat.ptr(turns), float(scales[i]), nat.stream_ptr()))
            yield DeviceArray(field[0].clone()), DeviceArray(phi)
        nat.check(lib.pa_vacuum_leg(h, nat.ptr(field), 1, float(legs[-1]), wvl, nat.stream_ptr()))
        return DeviceArray(field[0])


class IdenticalPhaseScreensPath(PhaseScreensPath):
    """`count` copies of one screen at equal spacing (pathes.py:78-95)."""

    def __init__(self, length, count, phase_screen, position_in_slab="middle", losses_db=0):
        thickness = length / count
        offsets = {"before": 0.0, "middle": 1 / 2, "after": 1.0}
        if position_in_slab not in offsets:
            raise ValueError("Available values for position_in_slab: 'before', 'middle' and 'after'")
        positions = (np.arange(count) + offsets[position_in_slab]) * thickness
        phase_screen.thickness = thickness
        phase_screens = [copy.copy(phase_screen) for _ in range(count)]
        self.phase_screen = phase_screens[0]
        super().__init__(length=length, phase_screens=phase_screens, positions=positions, losses_db=losses_db)
