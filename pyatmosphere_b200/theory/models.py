"""Turbulence spectrum models, host float64 (one-time ring integrals and coefficient amplitudes only).
Mirror of /root/reference/pyatmosphere/theory/models.py:10-32,80-101 (Model, MVKModel, AndrewsModel): same names,
same arithmetic order (the amplitudes feed seeded parity tests bit for bit), own layout."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_KOLMOGOROV = 0.033
_POWER = 11 / 6


@dataclass
class Model:
    """Refractive-index spectrum Phi_n(kappa) with structure constant Cn2, inner scale l0 and outer scale L0."""
    Cn2: float
    l0: float
    L0: float

    # ---- conversions shared by every model (theory/models.py:16-23) -------------------------------------------
    @staticmethod
    def _slab_factor(k, thickness):
        """Phi_phi = 2 pi k^2 dz Phi_n for a slab of thickness dz at optical wavenumber k."""
        return 2 * np.pi * k**2 * thickness

    def psd_n_f(self, f):
        kappa = 2 * np.pi * f
        return self.psd_n(kappa)

    def psd_phi(self, kappa, k, thickness):
        return self._slab_factor(k, thickness) * self.psd_n(kappa)

    def psd_phi_f(self, f, k, thickness):
        return self._slab_factor(k, thickness) * self.psd_n_f(f)

    def _outer_cutoff(self):
        return (2 * np.pi) / self.L0


class MVKModel(Model):
    """Modified von Karman: 0.033 Cn2 exp(-(kappa/km)^2) / (kappa^2 + k0^2)^(11/6), km = 5.92 / l0, k0 = 2 pi / L0."""

    def psd_n(self, kappa):
        outer = self._outer_cutoff()
        inner = 5.92 / self.l0
        damping = np.exp(-(kappa / inner) ** 2)
        return _KOLMOGOROV * self.Cn2 * damping / (kappa**2 + outer**2) ** _POWER


class AndrewsModel(Model):
    """Andrews' modified atmospheric spectrum (high-wavenumber bump): the von Karman form with the inner-scale factor
    (1 + 1.802 q - 0.254 q^(7/6)) exp(-q^2), q = kappa / kl, kl = 3.3 / l0  (theory/models.py:94-101)."""

    def psd_n(self, kappa):
        q = kappa / (3.3 / self.l0)
        outer = (2 * np.pi) / self.L0
        bump = 1 + 1.802 * q - 0.254 * q ** (7 / 6)
        return _KOLMOGOROV * self.Cn2 * bump * np.exp(-(q) ** 2) / (kappa**2 + outer**2) ** _POWER
