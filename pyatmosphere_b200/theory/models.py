"""Turbulence spectrum models, host float64 (one-time ring integrals only).
Mirror of /root/reference/pyatmosphere/theory/models.py:10-32,80-101 (Model, MVKModel, AndrewsModel)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Model:
    Cn2: float
    l0: float
    L0: float

    def psd_n_f(self, f):
        return self.psd_n(2 * np.pi * f)

    def psd_phi(self, kappa, k, thickness):
        return 2 * np.pi * k**2 * thickness * self.psd_n(kappa)

    def psd_phi_f(self, f, k, thickness):
        return 2 * np.pi * k**2 * thickness * self.psd_n_f(f)


class MVKModel(Model):
    """Modified von Karman: 0.033 Cn2 exp(-(kappa/km)^2) / (kappa^2 + k0^2)^(11/6)."""

    def psd_n(self, kappa):
        k0 = (2 * np.pi) / self.L0
        km = 5.92 / self.l0
        return 0.033 * self.Cn2 * np.exp(-(kappa / km) ** 2) / (kappa**2 + k0**2) ** (11 / 6)


class AndrewsModel(Model):
    """Andrews' modified atmospheric spectrum (high-wavenumber bump): the von Karman form with the inner-scale factor
    (1 + 1.802 q - 0.254 q^(7/6)) exp(-q^2), q = kappa / kl, kl = 3.3 / l0  (theory/models.py:94-101)."""

    def psd_n(self, kappa):
        kl = 3.3 / self.l0
        k0 = (2 * np.pi) / self.L0
        q = kappa / kl
        bump = 1 + 1.802 * q - 0.254 * q ** (7 / 6)
        return 0.033 * self.Cn2 * bump * np.exp(-(q) ** 2) / (kappa**2 + k0**2) ** (11 / 6)
