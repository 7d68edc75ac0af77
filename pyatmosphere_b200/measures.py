"""Observables of an output field.  Mirror of /root/reference/pyatmosphere/measures.py:1-38: same names,
same arguments (`channel`, optional `output=`, anything else is forwarded to `channel.run`), python floats out.
One fused device sweep (pa_measure) yields all of them; each function picks its entry."""
from __future__ import annotations

import numpy as np

from . import _engine as eng
from . import _native as nat
from .gpu import DeviceArray


def _field_tensor(channel, output, args, kwargs):
    ctx = eng.channel_context(channel)
    if output is None:
        output = channel.run(*args, **kwargs)
    if isinstance(output, DeviceArray):
        t = output.t
    else:
        t = nat.torch_mod().as_tensor(np.asarray(output), device=ctx.tdevice)
    return ctx, t.to(ctx.cdtype).contiguous()


def all_moments(channel, output=None, pupils=(), *args, **kwargs):
    """dict of every measure of `output` (and the transmittance of each (radius, (sx, sy)) in `pupils`)."""
    ctx, t = _field_tensor(channel, output, args, kwargs)
    torch = nat.torch_mod()
    batch = 1 if t.ndim == 2 else t.shape[0]
    etas = []
    head = None
    pupils = list(pupils)
    chunks = [pupils[i:i + nat.MAX_PUPILS] for i in range(0, len(pupils), nat.MAX_PUPILS)] or [[]]
    for chunk in chunks:
        stride = nat.MEASURE_HEAD + nat.MAX_PUPILS
        out = torch.empty((batch, stride), dtype=torch.float64, device=ctx.tdevice)
        tab = np.array([[np.float32(r**2), np.float32(s[0]), np.float32(s[1])] for r, s in chunk], dtype=np.float32).reshape(-1, 3)
        tab_d = torch.as_tensor(tab, device=ctx.tdevice) if len(chunk) else None
        nat.check(ctx.lib.pa_measure(ctx.handle, nat.ptr(t), batch, nat.ptr(tab_d), len(chunk), 0, nat.ptr(out), stride,
                                     nat.stream_ptr()))
        host = out.cpu().numpy()
        head = host[:, :nat.MEASURE_HEAD]
        etas.append(host[:, nat.MEASURE_HEAD:nat.MEASURE_HEAD + len(chunk)])
    res = {name: head[:, i] for i, name in enumerate(nat.MEASURE_NAMES)}
    res["eta_pupil"] = np.concatenate(etas, axis=1) if etas else np.zeros((batch, 0))
    return res


def _one(channel, name, args, kwargs, force_no_pupil):
    if force_no_pupil:
        kwargs["pupil"] = False
    output = kwargs.pop("output", None)
    if output is None and args and not isinstance(args[0], (bool, int)):
        output, args = args[0], args[1:]
    return float(all_moments(channel, output, (), *args, **kwargs)[name][0])


def I(channel, output=None, *args, **kwargs):
    """|u|^2 (measures.py:1-4) as a device array."""
    ctx, t = _field_tensor(channel, output, args, kwargs)
    torch = nat.torch_mod()
    out = torch.empty(t.shape, dtype=ctx.rdtype, device=ctx.tdevice)
    batch = 1 if t.ndim == 2 else t.shape[0]
    nat.check(ctx.lib.pa_intensity(ctx.handle, nat.ptr(t), nat.ptr(out), batch, nat.stream_ptr()))
    return DeviceArray(out)


def eta(channel, *args, **kwargs):
    return _one(channel, "eta", args, kwargs, force_no_pupil=False)


def mean_x(channel, *args, **kwargs):
    return _one(channel, "mean_x", args, kwargs, force_no_pupil=True)


def mean_y(channel, *args, **kwargs):
    return _one(channel, "mean_y", args, kwargs, force_no_pupil=True)


def mean_x2(channel, *args, **kwargs):
    return _one(channel, "mean_x2", args, kwargs, force_no_pupil=True)


def mean_xy(channel, *args, **kwargs):
    return _one(channel, "mean_xy", args, kwargs, force_no_pupil=True)


def mean_y2(channel, *args, **kwargs):
    return _one(channel, "mean_y2", args, kwargs, force_no_pupil=True)
