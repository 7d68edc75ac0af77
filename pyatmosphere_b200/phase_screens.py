"""Random phase screens.  Mirror of /root/reference/pyatmosphere/phase_screens.py:9-34 (PhaseScreen), :70-136
(SSPhaseScreen), :154-179 (SUPhaseScreen), :182-215 (WindSUPhaseScreen) and :37-67 (FFTPhaseScreen).  Spectra are drawn on the host from numpy's
global RNG in the reference's order; the sum of harmonics (pa_screen_ss) and the inverse transform of the FFT
screens (pa_screen_fft) run in libpyatm_b200.so."""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.integrate import quad

from . import _engine as eng
from . import _native as nat
from . import gpu
from .gpu import DeviceArray
from .grids import RectGrid
from .utils import Default, PolarDiscreteFunction

_PSD_CACHE = {}


def _accepts_real_only(method) -> bool:
    """The screens of this package synthesise only the real part when asked to (half the work)."""
    import inspect
    try:
        params = inspect.signature(method).parameters
    except (TypeError, ValueError):
        return False
    return "real_only" in params or any(q.kind is q.VAR_KEYWORD for q in params.values())


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
        if _accepts_real_only(self.generate_phase_screen):
            return self.generate_phase_screen(*args, real_only=True, **kwargs)
        # a subclass written against the reference's contract (phase_screens.py:21-28): complex screen, real part taken here
        return self.generate_phase_screen(*args, **kwargs).real

    def generator(self, *args, **kwargs):
        while True:
            ps = self.generate(complex=True, *args, **kwargs)
            yield ps.real
            yield ps.imag


class HarmonicSumScreen(PhaseScreen):
    """Screens of the form  sum_m c_m exp(2 pi i (y fy_m + x fx_m))  over one harmonic per annulus of a randomised
    log-polar grid (SSPhaseScreen, SUPhaseScreen).  Subclasses provide `_get_spectrum` (the draws, in the reference's
    order) and `_ring_power` (an upper bound of E|c_m|^2 / 2 per annulus, used to split the sum into the float64
    polynomial and the contraction and to scale the fp16 operands of the tensor-core method)."""

    device_rng = False          # pa_rng_spectrum draws c_m = n sqrt(power_m) with fixed ring powers: SSPhaseScreen only
    fusable = True              # the fused propagator may synthesise this screen itself from (fx, fy, c) and one shift

    def __init__(self, f_grid, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.f_grid = f_grid
        self.cache_clear()

    def cache_clear(self):
        self._cached_spectrum = None

    def _ring_power(self):
        raise NotImplementedError

    def _get_spectrum(self, use_cached_spectrum):
        raise NotImplementedError

    # ---- synthesis ---------------------------------------------------------------------------------------------
    def low_ring_plan(self, shift=(0, 0)):
        """(m_split, degree) of the polynomial part for this screen on its grid (see _engine.plan_low_rings)."""
        x, y = self.grid.get_xy()
        xe = float(np.max(np.abs(x + np.float32(shift[0]))))
        ye = float(np.max(np.abs(y + np.float32(shift[1]))))
        return eng.plan_low_rings(self.f_grid.base, self._ring_power(), xe, ye, eng.theta_cut(self.grid.resolution[0]),
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

    def _screen_for_path(self, shift=(0, 0), wind=False):
        """(turns, phi) of the next screen for the step-by-step path: the phase is reduced mod 2 pi in float64 on the
        device BEFORE it is rounded to the field's precision (a float32 phase of ~1e3 rad has an ulp of 6e-5 rad)."""
        spectrum = self._get_spectrum(use_cached_spectrum=wind)
        return self._synthesize(spectrum, shift, want_turns=True, want_phi=True)

    def generate_phase_screen(self, shift: Tuple[float, float] = (0, 0), wind: bool = False, real_only: bool = False):
        """phase_screens.py:108-136 / :166-179.  `wind=True` reuses the spectrum cached on this object (frozen flow)
        with a new `shift`; the reference's partial re-use of the previous screen's columns is an optimisation only
        (and is broken there) so the whole screen is re-synthesised."""
        gpu.require_gpu()
        spectrum = self._get_spectrum(use_cached_spectrum=wind)
        _, re = self._synthesize(spectrum, shift, want_turns=False, want_phi=True)
        if real_only:
            return DeviceArray(re)
        _, im = self._synthesize(spectrum, shift, want_turns=False, want_phi=True, imag_part=True)
        return DeviceArray(nat.torch_mod().complex(re, im))


class SSPhaseScreen(HarmonicSumScreen):
    """Sparse-spectrum screen: sum of `f_grid.points` random harmonics on a randomised log-polar grid, each carrying
    the whole phase variance of its annulus (phase_screens.py:70-136)."""

    device_rng = True

    def __init__(self, f_grid, *args, **kwargs):
        self._psd = None
        super().__init__(f_grid, *args, **kwargs)

    def _ring_power(self):
        return self._get_psd()

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


class SUPhaseScreen(HarmonicSumScreen):
    """Sparse-uniform screen (phase_screens.py:154-179): the same sum of harmonics with the spectrum SAMPLED at the
    drawn radius, c_m = (n0 + i n1) sqrt(psd_phi_f(rho_m) pi dk_m), dk_m = (2 pi)^2 (f_m^2 - f_{m-1}^2)."""

    def __init__(self, f_grid, *args, **kwargs):
        self._delta_k_base = None
        super().__init__(f_grid, *args, **kwargs)

    @property
    def delta_k_base(self):
        if self._delta_k_base is None:
            outer = self.f_grid.base
            self._delta_k_base = (2 * np.pi) ** 2 * np.array(outer**2 - np.insert(outer, 0, 0)[:-1] ** 2, dtype=np.float32)
        return self._delta_k_base

    def _ring_power(self):
        """max over each annulus of psd_phi_f(f) pi dk (sampled at 9 radii incl. both edges, float64, 5 % margin)."""
        outer = np.asarray(self.f_grid.base, dtype=np.float64)
        inner = np.insert(outer, 0, 0)[:-1]
        k = 2 * np.pi / self.wvl
        frac = np.linspace(0, 1, 9)[:, np.newaxis]
        radii = np.sqrt(inner**2 + frac * (outer**2 - inner**2))
        with np.errstate(all="ignore"):
            dens = np.asarray(self.model.psd_phi_f(radii, k, self.thickness), dtype=np.float64)
        dens = np.where(np.isfinite(dens), dens, 0.0).max(axis=0)
        return 1.05 * dens * np.pi * np.asarray(self.delta_k_base, dtype=np.float64)

    def _get_spectrum(self, use_cached_spectrum):
        if use_cached_spectrum and self._cached_spectrum:
            return self._cached_spectrum
        rho = self.f_grid.get_rho()
        theta = self.f_grid.get_theta()
        cn = (np.array([1, 1j]) @ np.random.normal(size=(2, self.f_grid.points))).astype(np.complex64) * \
            np.sqrt(self.model.psd_phi_f(rho, 2 * np.pi / self.wvl, self.thickness) * np.pi * self.delta_k_base)
        spectrum = PolarDiscreteFunction(rho=rho, theta=theta, value=cn)
        if use_cached_spectrum:
            self._cached_spectrum = spectrum
        return spectrum


class WindSUPhaseScreen(SUPhaseScreen):
    """Frozen-flow sparse-uniform screen (phase_screens.py:182-215): the coefficients are drawn once, every call
    returns the same screen translated by `speed` along x (offset = call index * speed).  As in the reference the unit
    normals are rounded to float32 before they are combined and `generate*` takes no shift / wind arguments.  The
    coefficients enter the kernels as complex64 (the reference carries the float32 x float32 products in double)."""

    fusable = False             # carries its own translation state: served by the step-by-step path

    def __init__(self, f_grid, speed, *args, **kwargs):
        self.speed = speed
        self.cnp = None
        super().__init__(f_grid, *args, **kwargs)

    def generate_cn(self):
        self.rho = self.f_grid.get_rho()
        self.theta = self.f_grid.get_theta()
        unit = np.random.normal(size=(2, self.f_grid.points)).astype(np.complex64)
        self.cnp = np.array([1, 1j]) @ unit
        self.iteration = 0

    def _get_spectrum(self, use_cached_spectrum=True):
        if self.cnp is None:
            self.generate_cn()
        outer = self.f_grid.base
        ring = np.array(outer**2 - np.insert(outer, 0, 0)[:-1] ** 2, dtype=np.float32)
        # float32 products in the reference's order, ((psd pi) (2 pi)^2) ring, not psd (pi delta_k_base)
        scale = np.sqrt(self.model.psd_phi_f(self.rho, 2 * np.pi / self.wvl, self.thickness) * np.pi * (2 * np.pi) ** 2 * ring)
        return PolarDiscreteFunction(rho=self.rho, theta=self.theta, value=self.cnp * scale)

    def _next_offset(self):
        if self.cnp is None:
            self.generate_cn()
        offset = self.iteration * self.speed
        self.iteration += 1
        return (offset, 0)

    def _screen_for_path(self, *args, **kwargs):
        if args or kwargs:
            raise TypeError("WindSUPhaseScreen.generate_phase_screen() takes no shift / wind arguments")
        shift = self._next_offset()
        return self._synthesize(self._get_spectrum(), shift, want_turns=True, want_phi=True)

    def generate_phase_screen(self, real_only: bool = False):
        gpu.require_gpu()
        shift = self._next_offset()
        spectrum = self._get_spectrum()
        _, re = self._synthesize(spectrum, shift, want_turns=False, want_phi=True)
        if real_only:
            return DeviceArray(re)
        _, im = self._synthesize(spectrum, shift, want_turns=False, want_phi=True, imag_part=True)
        return DeviceArray(nat.torch_mod().complex(re, im))

    def generator(self):
        while True:
            yield self.generate(complex=False)


class FFTPhaseScreen(PhaseScreen):
    """Classic FFT screen with optional 3x3 subharmonic levels (phase_screens.py:37-67): white complex noise shaped by
    sqrt(psd_phi_f) 2 pi df on the reciprocal grid, inverse-transformed (pa_screen_fft), plus the low-frequency
    harmonics of every subharmonic level, minus the mean.  The noise is drawn on the host from numpy's global RNG in
    the reference's order (real part first, then imaginary, main grid before the levels)."""

    def __init__(self, subharmonics, *args, **kwargs):
        self.subharmonics = subharmonics
        super().__init__(*args, **kwargs)

    def _draw_coefficients(self, f_grid):
        """phase_screens.py:43-48 on one frequency grid (the main one or a 3x3 subharmonic patch)."""
        noise = (np.random.normal(size=f_grid.shape) + 1j * np.random.normal(size=f_grid.shape)).astype(np.complex64)
        cn = noise * np.sqrt(self.model.psd_phi_f(f_grid.get_rho(), 2 * np.pi / self.wvl, self.thickness)) * 2 * np.pi * f_grid.delta
        cn[f_grid.origin_index] = 0
        return cn

    def _draw(self):
        """(cn [N][N], terms [T][4] = fx, fy, Re c, Im c) of one screen."""
        f_grid = self.grid.get_f_grid()
        cn = self._draw_coefficients(f_grid)
        terms = []
        for level in range(self.subharmonics):
            patch = RectGrid(3, f_grid.delta / 3 ** (level + 1))
            c = self._draw_coefficients(patch)
            f = patch.get_x()
            for i in range(patch.resolution[0]):
                for j in range(patch.resolution[1]):
                    if c[i, j] != 0:            # the zeroed centre contributes nothing
                        terms.append((float(f[0, i]), float(f[0, j]), float(c[i, j].real), float(c[i, j].imag)))
        return cn, np.array(terms, dtype=np.float64).reshape(-1, 4)

    def generate_phase_screen(self, real_only: bool = False):
        gpu.require_gpu()
        ctx = eng.grid_context(self.grid)
        torch = nat.torch_mod()
        cn, terms = self._draw()
        n = ctx.n
        if cn.shape != (n, n):
            raise ValueError("FFTPhaseScreen needs a square grid")
        host = np.ascontiguousarray(cn, dtype=np.complex64 if ctx.precision == nat.PA_C64 else np.complex128)
        spec = torch.as_tensor(host, device=ctx.tdevice)
        out_c = None if real_only else torch.empty((n, n), dtype=ctx.cdtype, device=ctx.tdevice)
        out_r = torch.empty((n, n), dtype=ctx.rdtype, device=ctx.tdevice) if real_only else None
        terms = np.ascontiguousarray(terms)
        nat.check(ctx.lib.pa_screen_fft(ctx.handle, nat.ptr(spec), 1, nat.ptr(terms) if len(terms) else None, len(terms),
                                        nat.ptr(out_c), nat.ptr(out_r), nat.stream_ptr()))
        nat.torch_mod().cuda.current_stream().synchronize()       # `terms` / `host` are host buffers of this call
        return DeviceArray(out_r if real_only else out_c)
