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
