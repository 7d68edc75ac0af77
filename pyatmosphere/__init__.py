"""`pyatmosphere` import name for the B200-native implementation.

With the repository root (or an install of this tree) ahead of KlenM/pyAtmosphere on `sys.path`, the reference's own
user code -- README.md:26-106, main.ipynb -- runs unmodified:

    from pyatmosphere import gpu, QuickChannel, simulations
    from pyatmosphere.simulations import BeamResult
    import pyatmosphere.theory.models

Every module of `pyatmosphere_b200` is registered under the reference's module path (/root/reference/pyatmosphere/
__init__.py:1-17 and the sub-packages it imports); the objects are THE SAME objects (`pyatmosphere.gpu.config is
pyatmosphere_b200.gpu.config`), this package holds no code of its own.  There is still no CPU path: set
`gpu.config['use_gpu'] = True` as README.md:26-29 says, otherwise the first field operation raises NoCpuPathError.
"""
import importlib
import pkgutil
import sys

import pyatmosphere_b200 as _impl

_SKIP = {"pyatmosphere_b200.build"}            # the nvcc driver is not part of the reference's surface


def _register():
    names = [m.name for m in pkgutil.walk_packages(_impl.__path__, _impl.__name__ + ".")]
    for real in names:
        leaf = real.rsplit(".", 1)[-1]
        if real in _SKIP or leaf.startswith("_") or leaf.startswith("lib"):      # libpyatm_b200.so is a C library
            continue
        module = importlib.import_module(real)
        alias = __name__ + real[len(_impl.__name__):]
        sys.modules.setdefault(alias, module)
        parent, _, leaf = alias.rpartition(".")
        if parent == __name__:
            globals().setdefault(leaf, module)


_register()
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("_")})
__all__ = list(_impl.__all__)
__version__ = _impl.__version__
