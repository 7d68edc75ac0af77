"""Backend switch and device-array wrapper.

Mirrors /root/reference/pyatmosphere/gpu.py:4-20 (`config['use_gpu']`, `get_xp`, `get_array`).  In the
reference `use_gpu=True` swaps numpy for cupy; here it routes every array operation of the split-step path
to libpyatm_b200.so.  This package has NO numpy path for field arithmetic: with `use_gpu=False` the compute
entry points raise (use the reference itself for CPU runs).

Extra keys (all optional, the reference's single key keeps working):
  dtype         'complex64' (default) | 'complex128'
  screen_method 'auto' | 'exact' | 'tc'   float64 CUDA-core contraction | tcgen05 split-fp16 tensor-core contraction;
                'auto' = 'tc' for complex64 grids whose size is a multiple of 256, else 'exact'
  theta_cut     phase-argument bound (rad) below which rings are summed as a float64 polynomial
                (None = 10 for the tensor-core method, 2 for the exact one)
  rng           'numpy' (reference draw order from numpy's global RNG; parity mode) | 'philox' (device RNG)
  seed          seed of the device RNG
  batch         realizations per launch in Simulation's batched fast path
"""
from __future__ import annotations

import numpy as np

config = {
    "use_gpu": True,
    "dtype": "complex64",
    "screen_method": "auto",
    "theta_cut": None,
    "rng": "numpy",
    "seed": 0,
    "batch": 8,
}


class NoCpuPathError(RuntimeError):
    pass


def require_gpu():
    if not config["use_gpu"]:
        raise NoCpuPathError("pyatmosphere_b200 has no CPU path: set gpu.config['use_gpu'] = True "
                             "(or run the reference package for numpy execution)")


def precision() -> int:
    dt = np.dtype(config["dtype"])
    if dt == np.complex64:
        return 0
    if dt == np.complex128:
        return 1
    raise ValueError("gpu.config['dtype'] must be complex64 or complex128")


def get_xp():
    """Host-side array module for the small float32/float64 vectors (axes, ring tables, coefficients).
    Fields never pass through it."""
    return np


class DeviceArray:
    """Thin owner of a torch CUDA tensor that quacks enough like a cupy array for user code written against
    the reference: `.get()`, `__array__`, `.shape`, `.dtype`, `.copy()`, `abs()`, `.real/.imag`, `.item()`,
    `.sum()`, `*`, `**`.  The named measures and the path use native kernels on `.t` directly; the generic
    operators below are conveniences for arbitrary user lambdas and delegate to torch."""

    __array_priority__ = 1000

    def __init__(self, t):
        self.t = t

    # ---- cupy-like surface ------------------------------------------------------------------------------
    def get(self):
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.ndim

    @property
    def dtype(self):
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    @property
    def size(self):
        return self.t.numel()

    def copy(self):
        return DeviceArray(self.t.clone())

    def item(self):
        return self.t.item()

    def astype(self, dtype):
        import torch
        return DeviceArray(self.t.to(getattr(torch, np.dtype(dtype).name)))

    @property
    def real(self):
        return DeviceArray(self.t.real.contiguous()) if self.t.is_complex() else self

    @property
    def imag(self):
        return DeviceArray(self.t.imag.contiguous())

    def __abs__(self):
        return DeviceArray(self.t.abs())

    def sum(self, axis=None):
        return DeviceArray(self.t.sum() if axis is None else self.t.sum(dim=axis))

    def __getitem__(self, idx):
        return DeviceArray(self.t[idx])

    def _other(self, o):
        import torch
        if isinstance(o, DeviceArray):
            return o.t
        if isinstance(o, np.ndarray):
            return torch.as_tensor(o, device=self.t.device)
        return o

    def __mul__(self, o):
        return DeviceArray(self.t * self._other(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return DeviceArray(self.t + self._other(o))

    __radd__ = __add__

    def __sub__(self, o):
        return DeviceArray(self.t - self._other(o))

    def __neg__(self):
        return DeviceArray(-self.t)

    def __truediv__(self, o):
        return DeviceArray(self.t / self._other(o))

    def __pow__(self, p):
        return DeviceArray(self.t ** p)

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.t.device})"


def get_array(xp_array):
    """gpu.py:17-20: bring an array to the host (`.get()` on device arrays, identity otherwise)."""
    if isinstance(xp_array, DeviceArray):
        return xp_array.get()
    return xp_array
