"""Analytic Gaussian beam (mirror of /root/reference/pyatmosphere/theory/sources.py:6-33).  The amplitude on
the grid is produced by the native library (pa_source_gaussian); only closed forms live here."""
from __future__ import annotations

import numpy as np


class GaussianBeam:
    def __init__(self, wvl, w0, F0):
        self.wvl = wvl
        self.w0 = w0
        self.F0 = F0

    @property
    def k(self):
        return 2 * np.pi / self.wvl

    def get_theta0(self, length):
        return 1 - length / self.F0

    def get_Lambda0(self, length):
        return 2 * length / self.k / self.w0**2

    def _norm(self, length):
        return self.get_theta0(length) ** 2 + self.get_Lambda0(length) ** 2

    def get_theta(self, length):
        return self.get_theta0(length) / self._norm(length)

    def get_Lambda(self, length):
        return self.get_Lambda0(length) / self._norm(length)

    def get_w(self, length):
        return self.w0 * np.sqrt(self._norm(length))
