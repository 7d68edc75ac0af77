"""pyatmosphere_b200 -- B200-native split-step beam propagation behind pyAtmosphere's Python API.

    import pyatmosphere_b200 as pyatmosphere        # drop-in for the hot path of KlenM/pyAtmosphere

Same names as /root/reference/pyatmosphere/__init__.py:1-17 for the path in scope: Channel, QuickChannel,
RectGrid, RandLogPolarGrid, GaussianSource, IdenticalPhaseScreensPath, SSPhaseScreen, MVKModel, CirclePupil,
measures, simulations, gpu.  Field arithmetic runs in libpyatm_b200.so (hand-written sm_100a CUDA behind a C
ABI, include/pyatm_b200.h); there is no CPU path.
"""
from . import gpu
from . import measures
from . import simulations
from .channels import Channel, QuickChannel
from .grids import *  # noqa: F401,F403
from .pathes import *  # noqa: F401,F403
from .phase_screens import *  # noqa: F401,F403
from .pupils import *  # noqa: F401,F403
from .sources import *  # noqa: F401,F403
from .theory.models import *  # noqa: F401,F403

# `from pyatmosphere import *` exports exactly these two names in the reference too (__init__.py:14-17); every other
# name above is reachable by explicit import, as there
__all__ = ["Channel", "QuickChannel"]
__version__ = "0.2.0"
