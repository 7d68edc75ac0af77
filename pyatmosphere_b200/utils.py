"""Descriptors that wire the parts of a channel together, and the container of a drawn spectrum.
Behavioural mirror of /root/reference/pyatmosphere/utils.py:7-39; the centred fft2/ifft2 helpers of utils.py:42-50 are
calls into the native library (pa_fft2c) -- the split-step path itself never uses them, its legs are fused passes
(pathes.VacuumPath -> pa_vacuum_leg)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence


class CrossRef:
    """Attribute whose value gets a back-reference to its owner (utils.py:7-21).
    Setting a falsy value is silently ignored, exactly as in the reference (`Channel(pupil=None)`)."""

    def __init__(self, back_name: str):
        self.back_name = back_name

    def __set_name__(self, owner, name):
        self.slot = "_" + name

    def __get__(self, obj, objtype=None):
        return getattr(obj, self.slot)

    def __set__(self, obj, value):
        if not value:
            return
        setattr(obj, self.slot, value)
        setattr(value, self.back_name, obj)


class Default:
    """Read-only attribute resolved by walking a dotted path from the instance (utils.py:24-32); assigning
    on the instance shadows it, because this is a non-data descriptor."""

    def __init__(self, dotted: str):
        self.chain = dotted.split(".")

    def __get__(self, obj, objtype=None):
        node = obj
        for name in self.chain:
            node = getattr(node, name)
        return node


@dataclass
class PolarDiscreteFunction:
    rho: Sequence[float]
    theta: Sequence[float]
    value: Sequence[float]


def _centred_transform(data, scale, forward):
    """[N][N] or [batch][N][N] complex array (host or device) -> DeviceArray of the same shape; see pa_fft2c."""
    import numpy as np

    from . import _engine as eng
    from . import _native as nat
    from . import gpu
    from .grids import RectGrid
    gpu.require_gpu()
    torch = nat.torch_mod()
    t = data.t if isinstance(data, gpu.DeviceArray) else torch.as_tensor(np.asarray(data), device="cuda")
    if t.ndim not in (2, 3) or t.shape[-1] != t.shape[-2]:
        raise ValueError("fft2 / ifft2 take square [N][N] (or [batch][N][N]) arrays")
    n = int(t.shape[-1])
    ctx = eng.grid_context(RectGrid(n, 1.0))          # the transform itself does not depend on the spacing: one context per N
    src = t.to(ctx.cdtype).reshape(-1, n, n).contiguous()
    out = torch.empty_like(src)
    nat.check(ctx.lib.pa_fft2c(ctx.handle, nat.ptr(src), nat.ptr(out), int(src.shape[0]), int(forward), float(scale),
                               nat.stream_ptr()))
    return gpu.DeviceArray(out.reshape(tuple(t.shape)))


def fft2(x, delta):
    """utils.py:42-44: fftshift(fft2(fftshift(x))) * delta**2 (grid sizes are even, so fftshift == ifftshift)."""
    return _centred_transform(x, float(delta) ** 2, True)


def ifft2(x, delta):
    """utils.py:47-50: ifftshift(ifft2(ifftshift(x))) * (N * delta)**2 with numpy's 1/N**2 inside ifft2, i.e. the plain
    sum times delta**2; `delta` is the frequency step."""
    return _centred_transform(x, float(delta) ** 2, False)
