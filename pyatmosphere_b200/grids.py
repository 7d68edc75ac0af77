"""Spatial and spectral grids (API of /root/reference/pyatmosphere/grids.py:12-119).

Only small host-side float32 vectors live here (a few KiB): the N x N arrays the reference derives from them
(rho^2, f^2, aperture masks) are never materialised; the kernels rebuild them from the axes, which are uploaded
once per context exactly as numpy produces them (float32 integer axis times the python-float spacing)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class Grid:
    def get_array_module(self):
        return np


@dataclass
class RectGrid(Grid):
    resolution: tuple
    delta: float

    def __post_init__(self):
        if isinstance(self.resolution, int):
            self.resolution = (self.resolution,) * 2

    # ---- geometry ---------------------------------------------------------------------------------------
    def _bounds(self, axis):
        n = self.resolution[axis]
        odd = bool(n % 2)
        return -n // 2 + odd, n // 2 + odd

    @property
    def _left_bound(self):
        return self._bounds(0)[0]

    @property
    def _right_bound(self):
        return self._bounds(0)[1]

    @property
    def _top_bound(self):
        return self._bounds(1)[0]

    @property
    def _bottom_bound(self):
        return self._bounds(1)[1]

    @property
    def shape(self):
        return self.resolution

    @property
    def size(self):
        return self.delta * np.array(self.resolution)

    @property
    def origin_index(self):
        nx, ny = self.resolution
        return nx // 2, ny // 2

    @property
    def extent(self):
        return self.delta * np.array([*self._bounds(0), *self._bounds(1)])

    # ---- coordinates ----------------------------------------------------------------------------------------
    def _axis(self, axis):
        lo, hi = self._bounds(axis)
        return np.arange(lo, hi, dtype=np.float32)

    def get_x(self):
        """(1, N) row, float32."""
        return self._axis(0)[np.newaxis, :] * self.delta

    def get_y(self):
        """(N, 1) column, float32."""
        return self._axis(1)[:, np.newaxis] * self.delta

    def get_xy(self):
        return self.get_x(), self.get_y()

    def get_NxNy(self):
        (x0, x1), (y0, y1) = self._bounds(0), self._bounds(1)
        return np.ogrid[y0:y1, x0:x1]

    def get_N2(self):
        rows, cols = self.get_NxNy()
        return rows**2 + cols**2

    def get_rho2(self):
        x, y = self.get_xy()
        return x**2 + y**2

    def get_rho(self):
        return np.sqrt(self.get_rho2())

    def get_f_grid(self):
        """Reciprocal grid: same (square) resolution, spacing 1/(N delta)."""
        n = np.min(self.resolution)
        return RectGrid(resolution=int(n), delta=1 / (n * self.delta))


@dataclass
class RandLogPolarGrid(Grid):
    """Log-spaced annuli with one random harmonic each.  Draws come from numpy's global legacy RNG in the
    reference's order, so `np.random.seed(s)` reproduces the reference's spectra."""
    points: int
    f_min: float
    f_max: float

    @property
    def base(self):
        """Outer edges of the annuli, float32."""
        log_edges = np.linspace(np.log(self.f_min), np.log(self.f_max), self.points, dtype=np.float32)
        return np.exp(log_edges)

    def get_rho(self):
        """Radius inside every annulus, uniform in area, from ONE uniform number shared by all annuli."""
        shared = np.random.random(size=(1,)).astype(np.float32)
        outer = self.base
        inner = np.insert(outer, 0, 0)[:-1]
        return np.sqrt(inner**2 + shared * (outer**2 - inner**2))

    def get_theta(self):
        return 2 * np.pi * np.random.random(size=(self.points,)).astype(np.float32)

    def get_x(self, rho, theta):
        return rho * np.cos(theta)

    def get_y(self, rho, theta):
        return rho * np.sin(theta)

    def get_xy(self, rho, theta):
        return self.get_x(rho, theta)[np.newaxis, :], self.get_y(rho, theta)[:, np.newaxis]
