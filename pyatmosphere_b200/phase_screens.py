"""Random phase screens.  Mirror of /root/reference/pyatmosphere/phase_screens.py:9-34 (PhaseScreen) and
:70-136 (SSPhaseScreen).  Spectra are drawn on the host from numpy's global RNG in the reference's order;
the sum of harmonics runs in libpyatm_b200.so (pa_screen_ss)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.integrate import quad

from . import _engine as eng
from . import _native as nat
from . import gpu
from .gpu import DeviceArray
from .utils import Default, PolarDiscreteFunction

_PSD_CACHE = {}


class PhaseScreen:
    wvl = Default("channel.source.wvl")
    grid = Default("channel.grid")

    def __init__(self, model, thickness=None, wvl=None, grid=None):
        self.model = model
        self.thickness = thickness
        if wvl:
            self.wvl = wvl
        if grid:
            self.grid = grid

    def generate_phase_screen(self, *args, **kwargs):
        """Return the complex phase screen."""
        raise NotImplementedError

    def generate(self, complex=False, *args, **kwargs):
        if complex:
            return self.generate_phase_screen(*args, **kwargs)
        return self.generate_phase_screen(*args, real_only=True, **kwargs)

    def generator(self, *args, **kwargs):
        while True:
            ps = self.generate(complex=True, *args, **kwargs)
            yield ps.real
            yield ps.imag


class SSPhaseScreen(PhaseScreen):
    """Sparse-spectrum screen: sum of `f_grid.points` random harmonics on a randomised log-polar grid."""

    def __init__(self, f_grid, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.f_grid = f_grid
        self._psd = None
        self.cache_clear()

    def cache_clear(self):
        self._cached_spectrum = None

    # ---- ring powers (phase_screens.py:82-91) --------------------------------------------------------------
    def _get_psd(self):
        """Phase variance of every annulus, float32, integrated once with scipy.quad like the reference.
        Identical screens of one path share the result through a module-level cache (same numbers)."""
        if self._psd is not None:
            return self._psd
        m = self.model
        key = (type(m).__name__, m.Cn2, m.l0, m.L0, float(self.wvl), float(self.thickness), self.f_grid.points,
               float(self.f_grid.f_min), float(self.f_grid.f_max))
        if key not in _PSD_CACHE:
            edges = self.f_grid.base
            k = 2 * np.pi / self.wvl

            def ring_integrand(f):
                return (2 * np.pi) ** 2 * f * m.psd_phi_f(f, k, self.thickness)

            _PSD_CACHE[key] = np.array(
                [2 * np.pi * quad(ring_integrand, edges[i - 1] if i != 0 else 0, edges[i])[0] for i in range(len(edges))],
                dtype=np.float32)
        self._psd = _PSD_CACHE[key]
        return self._psd

    # ---- random spectrum (phase_screens.py:93-106) ---------------------------------------------------------
    def _get_spectrum(self, use_cached_spectrum):
        if use_cached_spectrum and self._cached_spectrum:
            return self._cached_spectrum
        spectrum = PolarDiscreteFunction(
            rho=self.f_grid.get_rho(),
            theta=self.f_grid.get_theta(),
            value=(np.array([1, 1j]) @ np.random.normal(size=(2, self.f_grid.points))).astype(np.complex64)
            * np.sqrt(self._get_psd()))
        if use_cached_spectrum:
            self._cached_spectrum = spectrum
        return spectrum

    # ---- synthesis ---------------------------------------------------------------------------------------------
    def low_ring_plan(self, shift=(0, 0)):
        """(m_split, degree) of the polynomial part for this screen on its grid (see _engine.plan_low_rings)."""
        x, y = self.grid.get_xy()
        xe = float(np.max(np.abs(x + np.float32(shift[0]))))
        ye = float(np.max(np.abs(y + np.float32(shift[1]))))
        return eng.plan_low_rings(self.f_grid.base, self._get_psd(), xe, ye, eng.theta_cut(self.grid.resolution[0]),
                                  eng.screen_tolerance())

    def _synthesize(self, spectrum, shift, want_turns=True, want_phi=False, imag_part=False):
        """Run pa_screen_ss for one spectrum.  Returns (turns, phi) torch tensors (None when not requested)."""
        ctx = eng.grid_context(self.grid)
        torch = nat.torch_mod()
        fx, fy = self.f_grid.get_xy(spectrum.rho, spectrum.theta)
        coef = np.asarray(spectrum.value, dtype=np.complex64)
        if imag_part:
            coef = (coef * np.complex64(-1j)).astype(np.complex64)       # Re(-i z) = Im(z)
        m = coef.shape[0]
        m_split, degree = self.low_ring_plan(shift)
        dev = ctx.tdevice
        fx_d = torch.as_tensor(np.ascontiguousarray(fx, dtype=np.float32).ravel(), device=dev)
        fy_d = torch.as_tensor(np.ascontiguousarray(fy, dtype=np.float32).ravel(), device=dev)
        c_d = torch.as_tensor(np.ascontiguousarray(coef).view(np.float32), device=dev)
        n = ctx.n
        turns = torch.empty((n, n), dtype=ctx.rdtype, device=dev) if want_turns else None
        phi = torch.empty((n, n), dtype=ctx.rdtype, device=dev) if want_phi else None
        nat.check(ctx.lib.pa_screen_ss(ctx.handle, nat.ptr(fx_d), nat.ptr(fy_d), nat.ptr(c_d), m, m_split, degree,
                                       float(shift[0]), float(shift[1]), 1, nat.ptr(turns), nat.ptr(phi),
                                       1 if ctx.precision == nat.PA_C128 else 0,
                                       eng.screen_method(n),
                                       float(np.max(np.abs(coef[m_split:]))) if m_split < m else 1.0, nat.stream_ptr()))
        return turns, phi

    def generate_phase_screen(self, shift: Tuple[float, float] = (0, 0), wind: bool = False, real_only: bool = False):
        """phase_screens.py:108-136.  `wind=True` reuses the spectrum cached on this object (frozen flow) with a
        new `shift`; the reference's partial re-use of the previous screen's columns is an optimisation only
        (and is broken there) so the whole screen is re-synthesised."""
        gpu.require_gpu()
        spectrum = self._get_spectrum(use_cached_spectrum=wind)
        _, re = self._synthesize(spectrum, shift, want_turns=False, want_phi=True)
        if real_only:
            return DeviceArray(re)
        _, im = self._synthesize(spectrum, shift, want_turns=False, want_phi=True, imag_part=True)
        return DeviceArray(nat.torch_mod().complex(re, im))
