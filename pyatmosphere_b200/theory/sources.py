"""Closed forms of a Gaussian beam in vacuum (Andrews & Phillips parameters), API of
/root/reference/pyatmosphere/theory/sources.py:6-33.  The amplitude on the grid itself is produced on the device
(pa_source_gaussian / the analytic first leg); only the scalar formulas live here."""
from __future__ import annotations

import math


class GaussianBeam:
    def __init__(self, wvl, w0, F0):
        self.wvl, self.w0, self.F0 = wvl, w0, F0

    @property
    def k(self):
        return 2 * math.pi / self.wvl

    def _input_plane(self, length):
        """(Theta_0, Lambda_0): curvature and Fresnel parameters of the transmitter plane for a path `length`."""
        return 1 - length / self.F0, 2 * length / (self.k * self.w0**2)

    def get_theta0(self, length):
        return self._input_plane(length)[0]

    def get_Lambda0(self, length):
        return self._input_plane(length)[1]

    def get_theta(self, length):
        t0, l0 = self._input_plane(length)
        return t0 / (t0 * t0 + l0 * l0)

    def get_Lambda(self, length):
        t0, l0 = self._input_plane(length)
        return l0 / (t0 * t0 + l0 * l0)

    def get_w(self, length):
        """1/e^2 radius after `length` of vacuum: w0 sqrt(Theta_0^2 + Lambda_0^2)."""
        return self.w0 * math.hypot(*self._input_plane(length))
