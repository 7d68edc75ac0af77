"""ctypes binding of libpyatm_b200.so (include/pyatm_b200.h) -- the only door between the Python host and the
sm_100a kernels.  There is no CPU fallback: if the library or a CUDA device is missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYATM_LIB", os.path.join(_HERE, "libpyatm_b200.so"))

PA_C64, PA_C128 = 0, 1
PA_SCREEN_EXACT, PA_SCREEN_TC = 0, 1
MEASURE_HEAD = 8
MAX_PUPILS = 8
MEASURE_NAMES = ("eta", "mean_x", "mean_y", "mean_x2", "mean_xy", "mean_y2", "mean_x2_r")

_vp, _int, _dbl, _u64, _sz = C.c_void_p, C.c_int, C.c_double, C.c_ulonglong, C.c_size_t
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class PaPath(C.Structure):
    _fields_ = [("n_screens", _int), ("leg_lengths_host", _dp), ("screen_scale_host", _dp), ("final_scale", _dbl),
                ("wvl", _dbl), ("w0", _dbl), ("F0", _dbl), ("m", _int), ("m_split", _int), ("degree", _int),
                ("shift_x", _dbl), ("shift_y", _dbl), ("screen_method", _int), ("coef_bound", _dbl), ("from_field", _int)]


# name -> (restype, argtypes); must list every symbol include/pyatm_b200.h declares (checked by the tests)
SIGNATURES = {
    "pa_version": (_int, []),
    "pa_last_error": (C.c_char_p, []),
    "pa_device_count": (_int, [_ip]),
    "pa_launch_count": (_u64, [_int]),
    "pa_ctx_create": (_int, [C.POINTER(_vp), _int, _int, _int]),
    "pa_ctx_destroy": (_int, [_vp]),
    "pa_ctx_set_axes": (_int, [_vp, _vp, _vp, _dbl]),
    "pa_ctx_permutation": (_int, [_vp, _vp]),
    "pa_ctx_fft_geometry": (_int, [_vp, _vp]),
    "pa_source_gaussian": (_int, [_vp, _vp, _int, _dbl, _dbl, _dbl, _vp]),
    "pa_vacuum_leg": (_int, [_vp, _vp, _int, _dbl, _dbl, _vp]),
    "pa_fft2c": (_int, [_vp, _vp, _vp, _int, _int, _dbl, _vp]),
    "pa_gaussian_amplitude": (_int, [_vp, _vp, _vp, _sz, _dbl, _dbl, _dbl, _vp]),
    "pa_screen_ss": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _dbl, _dbl, _int, _vp, _vp, _int, _int, _dbl, _vp]),
    "pa_screen_fft": (_int, [_vp, _vp, _int, _vp, _int, _vp, _vp, _vp]),
    "pa_apply_screen": (_int, [_vp, _vp, _int, _vp, _dbl, _vp]),
    "pa_phase_to_turns": (_int, [_vp, _vp, _int, _vp, _sz, _vp]),
    "pa_intensity": (_int, [_vp, _vp, _vp, _int, _vp]),
    "pa_pupil_apply": (_int, [_vp, _vp, _vp, _int, _dbl, _dbl, _dbl, _vp]),
    "pa_measure": (_int, [_vp, _vp, _int, _vp, _int, _int, _vp, _int, _vp]),
    "pa_histogram": (_int, [_vp, _vp, _sz, _sz, _vp, _int, _vp, _vp]),
    "pa_rng_spectrum": (_int, [_vp, _u64, _u64, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pa_fft_pass": (_int, [_vp, _vp, _int, _int, _vp, _dbl, _dbl, _vp]),
    "pa_propagate": (_int, [_vp, C.POINTER(PaPath), _vp, _int, _vp, _vp, _vp, _vp]),
    "pa_simulate_batch": (_int, [_vp, C.POINTER(PaPath), _int, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _int, _vp,
                                 _int, _vp]),
    "pa_simulate_batch_async": (_int, [_vp, C.POINTER(PaPath), _int, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _int, _vp,
                                       _int, _vp]),
    "pa_stream_synchronize": (_int, [_vp, _vp]),
    "pa_simulate_batch_device": (_int, [_vp, C.POINTER(PaPath), _int, _u64, _u64, _vp, _vp, _vp, _int, _vp, _int,
                                        _vp]),
    "pa_comm_unique_id": (_int, [_vp]),
    "pa_comm_create": (_int, [C.POINTER(_vp), _int, _int, _int, _vp]),
    "pa_comm_destroy": (_int, [_vp]),
    "pa_stats_allreduce": (_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
}
COMM_ID_BYTES = 128

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -m pyatmosphere_b200.build` "
                "(pyatmosphere_b200 has no CPU or library fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().pa_last_error()
        raise NativeError((msg or b"unknown error").decode("utf-8", "replace") + f" (status {rc})")


def torch_mod():
    import torch
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device visible: pyatmosphere_b200 runs on B200 (sm_100a) only and has no CPU path")
    return torch


def stream_ptr():
    torch = torch_mod()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or pinned host) address of a torch tensor / host address of a numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


class Context:
    """One pa_ctx: device + grid size + precision + axes.  Cached by `context()`."""

    def __init__(self, device: int, n: int, precision: int, x: np.ndarray, y: np.ndarray, delta: float):
        lib = load()
        torch_mod()
        self.lib, self.device, self.n, self.precision, self.delta = lib, device, n, precision, float(delta)
        h = C.c_void_p()
        check(lib.pa_ctx_create(C.byref(h), device, n, precision))
        self.handle = h
        self.x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        self.y = np.ascontiguousarray(y, dtype=np.float32).ravel()
        check(lib.pa_ctx_set_axes(h, ptr(self.x), ptr(self.y), float(delta)))

    # dtypes -------------------------------------------------------------------------------------------
    @property
    def cdtype(self):
        torch = torch_mod()
        return torch.complex64 if self.precision == PA_C64 else torch.complex128

    @property
    def rdtype(self):
        torch = torch_mod()
        return torch.float32 if self.precision == PA_C64 else torch.float64

    @property
    def tdevice(self):
        return torch_mod().device("cuda", self.device)

    def empty_field(self, batch=1):
        return torch_mod().empty((batch, self.n, self.n), dtype=self.cdtype, device=self.tdevice)

    def permutation(self):
        out = np.empty(self.n, dtype=np.int32)
        check(self.lib.pa_ctx_permutation(self.handle, ptr(out)))
        return out

    def fft_geometry(self):
        out = np.zeros(6, dtype=np.int32)
        check(self.lib.pa_ctx_fft_geometry(self.handle, ptr(out)))
        keys = ("rows_threads", "rows_per_cta", "rows_smem", "cols_threads", "cols_per_cta", "cols_smem")
        return dict(zip(keys, (int(v) for v in out)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pa_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts = {}


def context(n: int, delta: float, x: np.ndarray, y: np.ndarray, precision: int, device: int | None = None) -> Context:
    torch = torch_mod()
    if device is None:
        device = torch.cuda.current_device()
    key = (device, int(n), int(precision), float(delta))
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(device, int(n), int(precision), x, y, delta)
        _contexts[key] = ctx
    return ctx


def any_context(precision: int) -> Context:
    """Some context of this precision on the current device -- for the element-wise entry points that need a device and a
    precision but no grid (pa_gaussian_amplitude); a small one is created when none exists yet."""
    device = torch_mod().cuda.current_device()
    for (dev, _n, prec, _delta), ctx in _contexts.items():
        if dev == device and prec == int(precision):
            return ctx
    axis = (np.arange(64, dtype=np.float32) - 32) * 1.0
    return context(64, 1.0, axis, axis, precision, device)


def clear_contexts():
    _contexts.clear()


def launch_count(reset=False) -> int:
    return int(load().pa_launch_count(1 if reset else 0))
