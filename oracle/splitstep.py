"""CPU oracle for the split-step beam-propagation hot path of KlenM/pyAtmosphere.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference's algorithm.  It may be
imported by `tests/`, by `__graft_entry__.smoke()` and by the `cpu_baseline` / `--impl reference` legs of
`bench.py` -- as the checker or as the timed CPU baseline -- and by nothing else.  The product package
(`pyatmosphere_b200`) never imports it and has no CPU path.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against fixtures produced
by importing and running the unmodified reference in the build container (`oracle/make_golden.py`, numpy
2.3.5 / scipy 1.18.1), and against the four asserts of the reference's own vacuum notebook
(`tests/itest_vacuum_propagation.ipynb` cells 3-7, restated).

Two arithmetic modes are offered for every array function:

* ``mode="ref"``  -- the dtypes the reference actually computes in under numpy >= 2 (NEP 50): float32 grids,
  complex64 screen contraction, complex128 transfer-function product, cast to complex64 after every leg.
  This mode is expected to agree with the imported reference to rounding (checked in the golden tests) and
  it is the mode timed as the CPU baseline, because it does the same work as the reference's numpy path.
* ``mode="f64"``  -- the same formulas with the float32 *inputs* (coordinates, frequencies, coefficients)
  promoted exactly to float64 and all arithmetic in float64/complex128.  This is the meaningful target for a
  new implementation: the reference's own complex64 screens are ~2.6e-4 rad away from it.

All file:line citations are relative to /root/reference/pyatmosphere/.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "rect_axis", "rect_xy", "f_grid_delta", "logpolar_base", "ring_rho", "mvk_psd_n", "ring_psd",
    "draw_spectrum", "spectrum_to_fxy", "ss_screen", "gaussian_source", "gaussian_width", "vacuum_leg",
    "screen_positions", "leg_lengths", "propagate", "circle_mask", "intensity", "moments", "rytov2",
    "analytic_gaussian_field", "pdt_histogram", "beam_statistics",
    "psd_phi_f", "andrews_psd_n", "su_delta_k", "draw_su_spectrum", "fft_screen_coefficients", "draw_fft_screen",
    "fft_screen", "draw_wind_su_spectrum",
]


# --------------------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------------------
def rect_axis(n: int, delta: float) -> np.ndarray:
    """Centred coordinate axis, float32: fl32(j) * delta for j in [-n//2, n//2) (even n).

    grids.py:39-53 (bounds, odd sizes start one later) and grids.py:63-69 (`arange(..., float32) * delta`;
    under NEP 50 the python float is weak, so the product is rounded to float32)."""
    odd = n % 2
    lo = -n // 2 + odd
    hi = n // 2 + odd
    return np.arange(lo, hi, dtype=np.float32) * delta


def rect_xy(n: int, delta: float):
    """x as a (1,n) row, y as an (n,1) column (grids.py:63-72). Row index <-> y, column index <-> x."""
    a = rect_axis(n, delta)
    return a.reshape(1, -1), a.reshape(-1, 1)


def f_grid_delta(n: int, delta: float) -> np.float64:
    """Spacing of the frequency grid, 1/(n*delta) as a numpy float64 (grids.py:82-85)."""
    return 1 / (np.min((n, n)) * delta)


def logpolar_base(points: int, f_min: float, f_max: float) -> np.ndarray:
    """Outer edges of the log-spaced annuli, float32 (grids.py:93-96)."""
    return np.exp(np.linspace(np.log(f_min), np.log(f_max), points, dtype=np.float32))


def ring_rho(base: np.ndarray, rand: np.ndarray) -> np.ndarray:
    """Radius of the harmonic drawn inside each annulus from ONE shared uniform number (grids.py:98-103).

    `rand` is the float32 array of shape (1,) the reference draws with np.random.random."""
    inner = np.insert(base, 0, 0)[:-1]
    return np.sqrt(inner**2 + rand * (base**2 - inner**2))


# --------------------------------------------------------------------------------------------------
# turbulence spectrum
# --------------------------------------------------------------------------------------------------
def mvk_psd_n(kappa, Cn2: float, l0: float, L0: float):
    """Modified von Karman refractive-index spectrum (theory/models.py:80-85)."""
    k0 = (2 * np.pi) / L0
    km = 5.92 / l0
    return 0.033 * Cn2 * np.exp(-(kappa / km) ** 2) / (kappa**2 + k0**2) ** (11 / 6)


def ring_psd(base: np.ndarray, Cn2: float, l0: float, L0: float, wvl: float, thickness: float) -> np.ndarray:
    """Phase variance of every annulus, float32 (phase_screens.py:82-91 with theory/models.py:22-23).

    psd_m = 2*pi * Integral_{f_{m-1}}^{f_m} (2*pi)^2 f * [2*pi k^2 dz Phi_n(2*pi f)] df, f_{-1} = 0,
    evaluated with scipy.integrate.quad defaults exactly as the reference does."""
    from scipy.integrate import quad

    k = 2 * np.pi / wvl

    def integrand(f):
        return (2 * np.pi) ** 2 * f * (2 * np.pi * k**2 * thickness * mvk_psd_n(2 * np.pi * f, Cn2, l0, L0))

    out = [2 * np.pi * quad(integrand, base[i - 1] if i != 0 else 0, base[i])[0] for i in range(len(base))]
    return np.array(out, dtype=np.float32)


def rytov2(Cn2: float, k: float, length: float) -> float:
    """Rytov variance (theory/atmosphere/__init__.py:5-6)."""
    return 1.23 * Cn2 * k ** (7 / 6) * length ** (11 / 6)


# --------------------------------------------------------------------------------------------------
# random spectrum of one screen
# --------------------------------------------------------------------------------------------------
def draw_spectrum(base: np.ndarray, psd: np.ndarray):
    """Draw (rho, theta, value) from numpy's GLOBAL legacy RNG in the reference's order.

    Order (verified against the reference): np.random.random(1) [grids.py:100],
    np.random.random(M) [grids.py:107], np.random.normal(size=(2,M)) [phase_screens.py:101].
    rho, theta are float32, value is complex64 = (n0 + i n1) * sqrt(psd)."""
    m = len(base)
    rand = np.random.random(size=(1,)).astype(np.float32)
    rho = ring_rho(base, rand)
    theta = 2 * np.pi * np.random.random(size=(m,)).astype(np.float32)
    value = (np.array([1, 1j]) @ np.random.normal(size=(2, m))).astype(np.complex64) * np.sqrt(psd)
    return rho, theta, value


def spectrum_to_fxy(rho: np.ndarray, theta: np.ndarray):
    """Cartesian frequencies, float32: fx (1,M) row and fy (M,1) column (grids.py:117-119)."""
    return (rho * np.cos(theta)).reshape(1, -1), (rho * np.sin(theta)).reshape(-1, 1)


# --------------------------------------------------------------------------------------------------
# sparse-spectrum screen
# --------------------------------------------------------------------------------------------------
def ss_screen(x, y, fx, fy, value, shift=(0.0, 0.0), mode: str = "ref", complex_out: bool = False,
              diag_product: bool = False):
    """phi[i,j] = Re sum_m value_m exp(2 pi i (y_i+sy) fy_m) exp(2 pi i fx_m (x_j+sx)).

    phase_screens.py:108-126 (contraction at :125-126), `.real` at :25-28.
    x (1,N), y (N,1) float32; fx (1,M), fy (M,1) float32; value (M,) complex64.
    mode="ref": float32/complex64 arithmetic like the reference; mode="f64": inputs promoted exactly.
    diag_product=True evaluates the "ref" mode as  E_y @ diag(value) @ E_x, the operation order of SUPhaseScreen
    (phase_screens.py:179); the value is the same sum, only the complex64 rounding differs."""
    if mode == "ref" and diag_product:
        full = np.exp(1j * 2 * np.pi * (y + shift[1]) @ fy.T) @ np.diag(value) @ np.exp(1j * 2 * np.pi * fx.T @ (x + shift[0]))
    elif mode == "ref":
        xs = x + shift[0]
        ys = y + shift[1]
        left = value * np.exp(1j * 2 * np.pi * ys @ fy.T)          # (N,M) complex64
        right = np.exp(1j * 2 * np.pi * fx.T @ xs)                 # (M,N) complex64
        full = left @ right
    elif mode == "f64":
        xs = (x + shift[0]).astype(np.float64)                     # the f32 sum is part of the input definition
        ys = (y + shift[1]).astype(np.float64)
        fx64 = fx.astype(np.float64)
        fy64 = fy.astype(np.float64)
        c = value.astype(np.complex128)
        left = c * np.exp(2j * np.pi * (ys @ fy64.T))
        right = np.exp(2j * np.pi * (fx64.T @ xs))
        full = left @ right
    else:
        raise ValueError(mode)
    return full if complex_out else full.real


# --------------------------------------------------------------------------------------------------
# the other screen generators
# --------------------------------------------------------------------------------------------------
def andrews_psd_n(kappa, Cn2: float, l0: float, L0: float):
    """Andrews' spectrum with the inner-scale bump (theory/models.py:94-101)."""
    kl = 3.3 / l0
    k0 = (2 * np.pi) / L0
    q = kappa / kl
    return 0.033 * Cn2 * (1 + 1.802 * q - 0.254 * q ** (7 / 6)) * np.exp(-(q) ** 2) / (kappa**2 + k0**2) ** (11 / 6)


def psd_phi_f(f, Cn2: float, l0: float, L0: float, k: float, thickness: float, psd_n=mvk_psd_n):
    """Phase spectrum of a slab over spatial frequency f: 2 pi k^2 dz Phi_n(2 pi f) (theory/models.py:16-23)."""
    return 2 * np.pi * k**2 * thickness * psd_n(2 * np.pi * f, Cn2, l0, L0)


def su_delta_k(base: np.ndarray) -> np.ndarray:
    """(2 pi)^2 (f_m^2 - f_{m-1}^2), float32: area of annulus m in kappa-space over pi (phase_screens.py:160-165)."""
    return (2 * np.pi) ** 2 * np.array(base**2 - np.insert(base, 0, 0)[:-1] ** 2, dtype=np.float32)


def draw_su_spectrum(base: np.ndarray, Cn2: float, l0: float, L0: float, wvl: float, thickness: float, psd_n=mvk_psd_n):
    """Sparse-uniform coefficients (phase_screens.py:166-176), drawn from numpy's GLOBAL RNG in the reference's order
    (random(1), random(M), normal(2,M)):  c_m = (n0 + i n1) sqrt(psd_phi_f(rho_m) pi dk_m), complex64."""
    m = len(base)
    rand = np.random.random(size=(1,)).astype(np.float32)
    rho = ring_rho(base, rand)
    theta = 2 * np.pi * np.random.random(size=(m,)).astype(np.float32)
    noise = (np.array([1, 1j]) @ np.random.normal(size=(2, m))).astype(np.complex64)
    value = noise * np.sqrt(psd_phi_f(rho, Cn2, l0, L0, 2 * np.pi / wvl, thickness, psd_n) * np.pi * su_delta_k(base))
    return rho, theta, value


def draw_wind_su_spectrum(base: np.ndarray, Cn2: float, l0: float, L0: float, wvl: float, thickness: float, psd_n=mvk_psd_n):
    """Coefficients of the frozen-flow sparse-uniform screen (phase_screens.py:189-206): same draw order as
    draw_su_spectrum, but the unit normals are rounded to float32 BEFORE they are combined with [1, i] (the
    `.astype(complex64)` sits on the normal array), so the product with the float32 amplitudes is complex128.
    Call k of the screen evaluates ss_screen(..., shift=(k * speed, 0)) with these coefficients (:208-211)."""
    m = len(base)
    rand = np.random.random(size=(1,)).astype(np.float32)
    rho = ring_rho(base, rand)
    theta = 2 * np.pi * np.random.random(size=(m,)).astype(np.float32)
    unit = np.array([1, 1j]) @ np.random.normal(size=(2, m)).astype(np.complex64)
    ring = np.array(base**2 - np.insert(base, 0, 0)[:-1] ** 2, dtype=np.float32)
    # float32 products in the reference's order: ((psd * pi) * (2 pi)^2) * ring   (phase_screens.py:203-205)
    value = unit * np.sqrt(psd_phi_f(rho, Cn2, l0, L0, 2 * np.pi / wvl, thickness, psd_n) * np.pi * (2 * np.pi) ** 2 * ring)
    return rho, theta, value


def fft_screen_coefficients(points: int, delta_f, Cn2: float, l0: float, L0: float, wvl: float, thickness: float,
                            psd_n=mvk_psd_n):
    """One draw of phase_screens.py:43-48 on the centred frequency grid RectGrid(points, delta_f): real normals
    first, then the imaginary ones; cn = noise(c64) * sqrt(psd_phi_f(|f|)) * 2 pi delta_f, centre element zeroed.
    `delta_f` is a numpy float64 in the reference (grids.py:82-85), which promotes the frequency axis -- and with
    it cn -- to double precision under numpy >= 2."""
    noise = (np.random.normal(size=(points, points)) + 1j * np.random.normal(size=(points, points))).astype(np.complex64)
    ax = rect_axis(points, 1.0).astype(np.float32) * delta_f
    rho = np.sqrt(ax.reshape(1, -1) ** 2 + ax.reshape(-1, 1) ** 2)
    cn = noise * np.sqrt(psd_phi_f(rho, Cn2, l0, L0, 2 * np.pi / wvl, thickness, psd_n)) * 2 * np.pi * delta_f
    cn[points // 2, points // 2] = 0
    return cn


def draw_fft_screen(n: int, delta: float, subharmonics: int, Cn2: float, l0: float, L0: float, wvl: float,
                    thickness: float, psd_n=mvk_psd_n):
    """All random inputs of one FFT screen in the reference's draw order (phase_screens.py:50-58): the n x n
    coefficients, then for every subharmonic level a 3 x 3 patch with spacing delta_f / 3^(level+1).
    Returns (cn, terms) with terms[t] = (fx, fy, c): the screen gains c exp(2 pi i (fx x + fy y))."""
    df = f_grid_delta(n, delta)
    cn = fft_screen_coefficients(n, df, Cn2, l0, L0, wvl, thickness, psd_n)
    terms = []
    for level in range(subharmonics):
        dsub = df / 3 ** (level + 1)
        c = fft_screen_coefficients(3, dsub, Cn2, l0, L0, wvl, thickness, psd_n)
        f = rect_axis(3, 1.0).astype(np.float32) * dsub
        for i in range(3):
            for j in range(3):
                terms.append((f[i], f[j], c[i, j]))          # f[i] pairs with x, f[j] with y (phase_screens.py:63-65)
    return cn, terms


def fft_screen(cn, terms, x, y, mode: str = "ref"):
    """Complex FFT screen (phase_screens.py:50-67): ifft2(cn, 1) [utils.py:47-50] + sum of the subharmonic terms, minus
    the mean.  mode="ref" lets numpy pick the dtypes as in the reference (double precision under numpy >= 2 because cn
    and the subharmonic frequencies are float64-based); mode="f64" promotes every input explicitly."""
    if mode == "f64":
        cn = np.asarray(cn, dtype=np.complex128)
        x = x.astype(np.float64)
        y = y.astype(np.float64)
    elif mode != "ref":
        raise ValueError(mode)
    screen = _centred_ifft2(cn, 1)
    for fx, fy, c in terms:
        if mode == "f64":
            fx, fy, c = np.float64(fx), np.float64(fy), np.complex128(c)
        screen = screen + c * np.exp(1j * 2 * np.pi * (fx * x + fy * y))
    return screen - np.mean(screen)


# --------------------------------------------------------------------------------------------------
# source
# --------------------------------------------------------------------------------------------------
def gaussian_source(x, y, w0: float, wvl: float, F0: float = np.inf, mode: str = "ref"):
    """Unit-power Gaussian beam sqrt(2/pi)/w0 exp(-(1/w0^2 + i k/(2 F0)) rho^2).

    theory/sources.py:16-18 on rho^2 = x^2+y^2 (float32, grids.py:74-76).  In "ref" mode the exponential is
    evaluated in complex64 and the float64 prefactor promotes the product to complex128, as numpy 2 does."""
    rho2 = x**2 + y**2
    a = 1 / w0**2 + 1j * 2 * np.pi / wvl / 2 / F0
    if mode == "ref":
        return np.sqrt(2 / np.pi) / w0 * np.exp(-a * rho2)
    if mode == "f64":
        return np.sqrt(2 / np.pi) / w0 * np.exp(-a * rho2.astype(np.float64))
    raise ValueError(mode)


def gaussian_width(w0: float, wvl: float, F0: float, length: float) -> float:
    """Analytic 1/e^2 beam radius after `length` (theory/sources.py:20-33)."""
    k = 2 * np.pi / wvl
    theta0 = 1 - length / F0
    lam0 = 2 * length / k / w0**2
    return w0 * np.sqrt(theta0**2 + lam0**2)


def analytic_gaussian_field(x, y, w0: float, wvl: float, length: float):
    """Closed-form collimated Gaussian beam after vacuum propagation, complex128."""
    k = 2 * np.pi / wvl
    zr = np.pi * w0**2 / wvl
    q = 1 + 1j * length / zr
    rho2 = x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2
    return np.sqrt(2 / np.pi) / w0 * np.exp(1j * k * length) / q * np.exp(-rho2 / (w0**2 * q))


# --------------------------------------------------------------------------------------------------
# vacuum leg
# --------------------------------------------------------------------------------------------------
def _centred_fft2(u, delta):
    """utils.py:42-44."""
    return np.fft.fftshift(np.fft.fft2(np.fft.fftshift(u))) * delta**2


def _centred_ifft2(u, delta_f):
    """utils.py:47-50."""
    n = u.shape[0]
    return np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(u))) * (n * delta_f) ** 2


def vacuum_leg(u, length, wvl: float, delta: float, mode: str = "ref"):
    """One Fresnel angular-spectrum leg: IFFT2c( e^{ikL} e^{-i pi L lambda f^2} FFT2c(u) ).

    theory/vacuum.py:5-7 called from pathes.py:27-40: f^2 is the rho^2 of RectGrid(N, 1/(N delta))
    (float32 axis times float64 spacing -> float64), the result is cast to complex64 (pathes.py:38);
    a non-positive length returns the input untouched (pathes.py:30,39-40)."""
    if not length > 0:
        return u
    n = u.shape[0]
    k = 2 * np.pi / wvl
    df = f_grid_delta(n, delta)
    fax = rect_axis(n, 1.0).astype(np.float32) * df        # float32 integer axis times float64 spacing
    f2 = fax.reshape(1, -1) ** 2 + fax.reshape(-1, 1) ** 2
    if mode == "ref":
        out = _centred_ifft2(np.exp(1j * k * length) * np.exp(-1j * np.pi * length * (2 * np.pi / k) * f2)
                             * _centred_fft2(u, delta), df)
        return out.astype(np.complex64)
    if mode == "f64":
        u128 = u.astype(np.complex128)
        out = _centred_ifft2(np.exp(1j * k * length) * np.exp(-1j * np.pi * length * (2 * np.pi / k) * f2)
                             * _centred_fft2(u128, delta), df)
        return out
    raise ValueError(mode)


# --------------------------------------------------------------------------------------------------
# path
# --------------------------------------------------------------------------------------------------
def screen_positions(length: float, count: int, where: str = "middle") -> np.ndarray:
    """Screen positions inside equal slabs (pathes.py:80-89)."""
    thickness = length / count
    if where == "before":
        return np.arange(count) * thickness
    if where == "middle":
        return (np.arange(count) + 1 / 2) * thickness
    if where == "after":
        return (np.arange(count) + 1) * thickness
    raise ValueError("Available values for position_in_slab: 'before', 'middle' and 'after'")


def leg_lengths(length: float, positions) -> list:
    """Vacuum-leg length in front of every screen plus the closing leg (pathes.py:68-69,75)."""
    legs = [positions[i] - positions[i - 1] if i > 0 else positions[0] for i in range(len(positions))]
    legs.append(length - positions[-1])
    return legs


def propagate(u0, screens, length: float, positions, wvl: float, delta: float, mode: str = "ref",
              losses_db: float = 0.0, keep_legs: bool = False, through_output: bool = True):
    """Split-step loop of pathes.py:61-75, optionally followed by AbstractPath.output (pathes.py:23-24).

    `screens` is a sequence of real phase arrays (N,N), already generated.  Each step does
    u <- exp(-i phi) * leg(u)  (pathes.py:72-73; propagate first, then multiply), then the closing leg.

    Losses, exactly as the reference applies them: the inner VacuumPath has losses_db = 0, so its `output`
    adds nothing; after each screen the field is attenuated by the leg's share `losses_db*leg/length`
    (pathes.py:71-73) -- and because `append_losses` uses `losses_db or self.losses_db` (pathes.py:20), a
    share of exactly 0 (zero-length first leg of "before") falls back to the FULL loss.  The closing leg gets
    no share.  `through_output=True` (Channel.run) applies the full loss once more (pathes.py:23-24);
    `through_output=False` is what Channel.generator stores in `channel.output` (channels.py:40-42)."""
    def lose(a, db):
        return a * 10 ** (-db / 20) if db else a

    legs = leg_lengths(length, positions)
    u = u0
    per_leg = []
    for i, phi in enumerate(screens):
        stepped = vacuum_leg(u, legs[i], wvl, delta, mode)
        share = losses_db * legs[i] / length
        u = lose(np.exp(-1j * phi) * stepped, share or losses_db)
        if keep_legs:
            per_leg.append(u)
    u = vacuum_leg(u, legs[-1], wvl, delta, mode)
    if through_output:
        u = lose(u, losses_db)
    return (u, per_leg) if keep_legs else u


# --------------------------------------------------------------------------------------------------
# pupil + measures
# --------------------------------------------------------------------------------------------------
def circle_mask(x, y, radius: float, shift=(0.0, 0.0)):
    """(x - sx)^2 + (y + sy)^2 <= r^2 in float32 (pupils.py:8-10; note the +sy)."""
    return (x - shift[0]) ** 2 + (y + shift[1]) ** 2 <= radius**2


def intensity(u):
    """measures.py:1-4."""
    return abs(u) ** 2


def moments(u, x, y, delta: float, pupils=(), mode: str = "ref") -> dict:
    """All seven reductions of measures.py:7-38 plus BeamResult.mean_x2_r (simulations/beam.py:26-33) and the
    aperture transmittance for each (radius, (sx, sy)) in `pupils` (simulations/pdt.py:21-26 -> measures.eta).

    mode="ref" follows the reference's dtype flow (float32 weights, numpy pairwise float32 sums);
    mode="f64" accumulates in float64 with the float32 coordinates promoted exactly."""
    if mode == "ref":
        inten = intensity(u)
        d2 = delta**2
        out = {
            "eta": (inten.sum(axis=(-1, -2)) * d2).item(),
            "mean_x": ((inten * x).sum(axis=(-1, -2)) * d2).item(),
            "mean_y": ((inten * (-1) * y).sum(axis=(-1, -2)) * d2).item(),
            "mean_x2": ((inten * x**2).sum(axis=(-1, -2)) * d2).item(),
            "mean_xy": ((inten * (-1 * x * y)).sum(axis=(-1, -2)) * d2).item(),
            "mean_y2": ((inten * y**2).sum(axis=(-1, -2)) * d2).item(),
        }
        r0 = np.sqrt(out["mean_x"] ** 2 + out["mean_y"] ** 2)
        cx, sx = out["mean_x"] / r0, out["mean_y"] / r0
        rot = (x * cx + ((-1) * y) * sx) ** 2
        out["mean_x2_r"] = ((inten * rot).sum(axis=(-1, -2)) * d2).item()
        out["eta_pupil"] = [((intensity(u * circle_mask(x, y, r, s))).sum(axis=(-1, -2)) * d2).item()
                            for r, s in pupils]
        return out
    if mode == "f64":
        inten = (u.real.astype(np.float64) ** 2 + u.imag.astype(np.float64) ** 2)
        x64, y64 = x.astype(np.float64), y.astype(np.float64)
        d2 = float(delta) ** 2
        out = {
            "eta": float(inten.sum() * d2),
            "mean_x": float((inten * x64).sum() * d2),
            "mean_y": float(-(inten * y64).sum() * d2),
            "mean_x2": float((inten * x64**2).sum() * d2),
            "mean_xy": float(-(inten * x64 * y64).sum() * d2),
            "mean_y2": float((inten * y64**2).sum() * d2),
        }
        r0 = np.hypot(out["mean_x"], out["mean_y"])
        cx, sx = out["mean_x"] / r0, out["mean_y"] / r0
        out["mean_x2_r"] = float((inten * (x64 * cx - y64 * sx) ** 2).sum() * d2)
        out["eta_pupil"] = [float((inten * circle_mask(x, y, r, s)).sum() * d2) for r, s in pupils]
        return out
    raise ValueError(mode)


# --------------------------------------------------------------------------------------------------
# Monte-Carlo post-processing
# --------------------------------------------------------------------------------------------------
def pdt_histogram(etas, bins: int = 200):
    """Histogram the reference draws for one pupil: `bins` equal bins on [0, 1] (simulations/pdt.py:30-31;
    matplotlib's hist == numpy.histogram: right edge of the last bin closed, out-of-range values dropped)."""
    return np.histogram(np.asarray(etas, dtype=np.float64), bins=bins, range=(0, 1))[0].astype(np.int64)


def beam_statistics(mean_x, mean_x2):
    """(value, error) for beam wander, long-term and short-term widths (simulations/beam.py:35-71)."""
    bw2 = np.asarray(mean_x, dtype=np.float64) ** 2
    lt2 = 4 * np.asarray(mean_x2, dtype=np.float64)
    st2 = lt2 - 4 * bw2

    def stat(v):
        m = np.sqrt(v.mean())
        return float(m), float(v.std(ddof=1) / np.sqrt(len(v)) / 2 / m)

    return {"bw": stat(bw2), "lt": stat(lt2), "st": stat(st2)}
