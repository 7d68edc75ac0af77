"""Descriptors that wire the parts of a channel together, and the container of a drawn spectrum.
Behavioural mirror of /root/reference/pyatmosphere/utils.py:7-39 (the centred fft2/ifft2 helpers of
utils.py:42-50 live in the native library: see _native / pathes.VacuumPath)."""
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
