"""Closed forms of a Gaussian beam in vacuum (Andrews & Phillips parameters), API of
/root/reference/pyatmosphere/theory/sources.py:6-33.  The amplitude on the channel grid is produced on the device inside
the first pass (pa_source_gaussian / the analytic first leg); `amplitude` on caller-supplied radii is pa_gaussian_amplitude;
the scalar formulas live here."""
from __future__ import annotations

import math


class GaussianBeam:
    def __init__(self, wvl, w0, F0):
        self.wvl, self.w0, self.F0 = wvl, w0, F0

    @property
    def k(self):
        return 2 * math.pi / self.wvl

    def amplitude(self, r2):
        """theory/sources.py:16-18: sqrt(2/pi)/w0 exp(-(1/w0^2 + i k/(2 F0)) r2) for an array of squared radii (host or
        device, any shape) -> DeviceArray in gpu.config['dtype']."""
        import numpy as np

        from .. import _native as nat
        from .. import gpu
        gpu.require_gpu()
        torch = nat.torch_mod()
        prec = gpu.precision()
        rdt = torch.float32 if prec == 0 else torch.float64
        t = r2.t if isinstance(r2, gpu.DeviceArray) else torch.as_tensor(np.asarray(r2), device="cuda")
        t = t.to(rdt).contiguous()
        out = torch.empty(t.shape, dtype=torch.complex64 if prec == 0 else torch.complex128, device=t.device)
        ctx = nat.any_context(prec)
        nat.check(ctx.lib.pa_gaussian_amplitude(ctx.handle, nat.ptr(t), nat.ptr(out), t.numel(), float(self.w0), float(self.wvl),
                                                float(self.F0), nat.stream_ptr()))
        return gpu.DeviceArray(out)

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
