"""Receiving apertures.  Mirror of /root/reference/pyatmosphere/pupils.py:4-13."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _engine as eng
from . import _native as nat
from .gpu import DeviceArray


@dataclass
class CirclePupil:
    radius: float

    def get_pupil(self, shift=(0, 0)):
        """Host boolean mask (x - sx)^2 + (y + sy)^2 <= r^2 in float32 (pupils.py:8-10), for inspection only;
        `output` evaluates the same predicate on the device."""
        x, y = self.channel.grid.get_xy()
        return (x - shift[0]) ** 2 + (y + shift[1]) ** 2 <= self.radius**2

    def output(self, input, shift=(0, 0)):
        ctx = eng.channel_context(self.channel)
        t = input.t if isinstance(input, DeviceArray) else nat.torch_mod().as_tensor(np.asarray(input), device=ctx.tdevice).to(ctx.cdtype)
        t = t.contiguous()
        out = nat.torch_mod().empty_like(t)
        batch = 1 if t.ndim == 2 else t.shape[0]
        nat.check(ctx.lib.pa_pupil_apply(ctx.handle, nat.ptr(t), nat.ptr(out), batch, float(self.radius), float(shift[0]),
                                         float(shift[1]), nat.stream_ptr()))
        return DeviceArray(out)
