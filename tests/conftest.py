import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_unavailable():
    """Reason why `-m gpu` tests cannot run here, or None.  (On the B200 box both conditions hold; a plain `pytest` on a
    CPU box then skips the GPU tests instead of failing them.)"""
    lib = os.path.join(ROOT, "pyatmosphere_b200", "libpyatm_b200.so")
    if not os.path.exists(lib):
        return "libpyatm_b200.so is not built (python -m pyatmosphere_b200.build)"
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
    except ImportError:
        return "torch is not importable"
    return None


def pytest_collection_modifyitems(config, items):
    reason = None
    for item in items:
        if "gpu" in item.keywords:
            reason = reason if reason is not None else (_gpu_unavailable() or "")
            if reason:
                item.add_marker(pytest.mark.skip(reason=reason))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        d = {k: z[k] for k in z.files}
    if "param_keys" in d:
        params = {}
        for k, v in zip(d["param_keys"], d["param_vals"]):
            v = str(v)
            try:
                params[str(k)] = int(v)
            except ValueError:
                try:
                    params[str(k)] = float(v)
                except ValueError:
                    params[str(k)] = v
        d["params"] = params
    return d


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
