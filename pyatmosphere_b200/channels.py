"""Channel facade.  Mirror of /root/reference/pyatmosphere/channels.py:17-80 (matplotlib is imported lazily,
only by `plot`)."""
from __future__ import annotations

import numpy as np

from .gpu import get_array
from .grids import RandLogPolarGrid, RectGrid
from .measures import I
from .pathes import IdenticalPhaseScreensPath
from .phase_screens import SSPhaseScreen
from .pupils import CirclePupil
from .sources import GaussianSource
from .theory.atmosphere import get_rytov2
from .theory.models import MVKModel
from .utils import CrossRef


class Channel:
    grid = CrossRef("channel")
    source = CrossRef("channel")
    path = CrossRef("channel")
    pupil = CrossRef("channel")

    def __init__(self, grid, source, path, pupil=None, name=""):
        self.grid = grid
        self.source = source
        self.path = path
        self.pupil = pupil
        self.output = None
        self.name = name

    def run(self, pupil=True, *args, **kwargs):
        out = self.path.output(self.source.output(), *args, **kwargs)
        return self.pupil.output(out) if pupil else out

    def generator(self, pupil=True, store_output=True, *args, **kwargs):
        """Per-slab (field, screen) pairs of the path (channels.py:37-44); the field behind the last leg is kept in
        `self.output` (behind the aperture if `pupil`) unless `store_output` is off."""
        self.output = None
        final = yield from self.path.generator(self.source.output(), *args, **kwargs)
        if store_output:
            self.output = self.pupil.output(final) if pupil else final

    def get_rythov2(self):
        return get_rytov2(self.path.phase_screen.model.Cn2, self.source.k, self.path.length)

    def plot(self, *args, **kwargs):
        from matplotlib import pyplot as plt
        plt.imshow(get_array(I(self, *args, **kwargs)), extent=self.grid.extent)


def QuickChannel(Cn2=1e-15, length=1e3, count_ps=5, beam_w0=0.09, beam_wvl=808e-9, aperture_radius=0.02,
                 grid_resolution=1024, grid_delta=0.001):
    """The reference's one-call channel (channels.py:53-80): 1024^2 grid at 1 mm, collimated Gaussian beam, `count_ps`
    identical MVK sparse-spectrum screens (l0 = 3 mm, L0 = 1 km, 2^10 rings from 1/(15 L0) to 2/l0), circular aperture."""
    inner_scale, outer_scale = 3e-3, 1e3
    screen = SSPhaseScreen(
        model=MVKModel(Cn2=Cn2, l0=inner_scale, L0=outer_scale),
        f_grid=RandLogPolarGrid(points=2**10, f_min=1 / outer_scale / 15, f_max=1 / inner_scale * 2))
    return Channel(
        grid=RectGrid(resolution=grid_resolution, delta=grid_delta),
        source=GaussianSource(wvl=beam_wvl, w0=beam_w0, F0=np.inf),
        path=IdenticalPhaseScreensPath(phase_screen=screen, length=length, count=count_ps),
        pupil=CirclePupil(radius=aperture_radius))
