"""Spatial and spectral grids.  Host-side float32 vectors only (a few KiB): the N x N arrays the reference
builds from them (rho^2, f^2, masks) are never materialised here, the kernels rebuild them from the axes.
Behavioural mirror of /root/reference/pyatmosphere/grids.py:12-119."""
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
            self.resolution = (self.resolution, self.resolution)

    # ---- geometry (grids.py:23-53) ----------------------------------------------------------------------
    @property
    def size(self):
        return np.array(self.resolution) * self.delta

    @property
    def shape(self):
        return self.resolution

    @property
    def origin_index(self):
        return (self.resolution[0] // 2, self.resolution[1] // 2)

    def _bounds(self, axis):
        n = self.resolution[axis]
        odd = bool(n % 2)
        return -n // 2 + odd, n // 2 + odd

    @property
    def extent(self):
        (l, r), (t, b) = self._bounds(0), self._bounds(1)
        return np.array([l, r, t, b]) * self.delta

    # ---- coordinates (grids.py:55-80): float32 integer axis times the python-float spacing ---------------
    def get_NxNy(self):
        (l, r), (t, b) = self._bounds(0), self._bounds(1)
        return np.ogrid[t:b, l:r]

    def get_N2(self):
        ny, nx = self.get_NxNy()
        return ny**2 + nx**2

    def get_x(self):
        l, r = self._bounds(0)
        return np.arange(l, r, dtype=np.float32).reshape((1, -1)) * self.delta

    def get_y(self):
        t, b = self._bounds(1)
        return np.arange(t, b, dtype=np.float32).reshape((-1, 1)) * self.delta

    def get_xy(self):
        return self.get_x(), self.get_y()

    def get_rho2(self):
        x, y = self.get_xy()
        return x**2 + y**2

    def get_rho(self):
        return np.sqrt(self.get_rho2())

    def get_f_grid(self):
        n = int(np.min(self.resolution))
        return RectGrid(resolution=n, delta=1 / (np.min(self.resolution) * self.delta))


@dataclass
class RandLogPolarGrid(Grid):
    """Log-spaced annuli with one random harmonic each (grids.py:88-119).  Draws come from numpy's global
    legacy RNG in the reference's order, so `np.random.seed(s)` reproduces the reference's spectra."""
    points: int
    f_min: float
    f_max: float

    @property
    def base(self):
        return np.exp(np.linspace(np.log(self.f_min), np.log(self.f_max), self.points, dtype=np.float32))

    def get_rho(self):
        u = np.random.random(size=(1,)).astype(np.float32)      # ONE number shared by all annuli (grids.py:100)
        outer = self.base
        inner = np.insert(outer, 0, 0)[:-1]
        return np.sqrt(inner**2 + u * (outer**2 - inner**2))

    def get_theta(self):
        return 2 * np.pi * np.random.random(size=(self.points,)).astype(np.float32)

    def get_x(self, rho, theta):
        return rho * np.cos(theta)

    def get_y(self, rho, theta):
        return rho * np.sin(theta)

    def get_xy(self, rho, theta):
        return self.get_x(rho, theta).reshape((1, -1)), self.get_y(rho, theta).reshape((-1, 1))
