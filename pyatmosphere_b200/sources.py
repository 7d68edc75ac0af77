"""Light sources.  Mirror of /root/reference/pyatmosphere/sources.py:7-23."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _engine as eng
from . import _native as nat
from .gpu import DeviceArray
from .theory.sources import GaussianBeam


@dataclass
class Source:
    wvl: float

    @property
    def k(self):
        return 2 * np.pi / self.wvl


class SourceField(DeviceArray):
    """Device field produced by GaussianSource.output().  It remembers that it is the untouched Gaussian of its
    channel, which lets the path generate it inside its first FFT pass instead of reading it back from HBM;
    any access to `.t` materialises it."""

    def __init__(self, source, ctx):
        self._source, self._ctx, self._t = source, ctx, None

    @property
    def is_virtual(self):
        return self._t is None

    @property
    def t(self):
        if self._t is None:
            f = self._ctx.empty_field(1)
            nat.check(self._ctx.lib.pa_source_gaussian(self._ctx.handle, nat.ptr(f), 1, float(self._source.w0),
                                                       float(self._source.wvl), float(self._source.F0), nat.stream_ptr()))
            self._t = f[0]
        return self._t

    @t.setter
    def t(self, value):
        self._t = value


class PlaneSource(Source):
    """Unit-amplitude plane wave (sources.py:16-18).  The reference returns the scalar 1, which none of its paths accepts
    (SURVEY.md App. B); here it is the field of ones on the channel grid, which they do."""

    def output(self):
        ctx = eng.channel_context(self.channel)
        return DeviceArray(nat.torch_mod().ones((ctx.n, ctx.n), dtype=ctx.cdtype, device=ctx.tdevice))


class GaussianSource(GaussianBeam, Source):
    """sqrt(2/pi)/w0 exp(-(1/w0^2 + i k/(2 F0)) rho^2) on the channel grid (sources.py:21-23 ->
    theory/sources.py:16-18), produced by pa_source_gaussian."""

    def output(self):
        return SourceField(self, eng.channel_context(self.channel))
