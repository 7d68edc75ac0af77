"""On-axis scintillation index per leg.  Mirror of /root/reference/pyatmosphere/simulations/si.py:11-36
(record + `si` statistic).  The closed-form Andrews curves the reference overlays (theory/atmosphere/si.py) are
analytic post-processing outside this path: pass them in `theoretical_functions` if wanted."""
from __future__ import annotations

import numpy as np

from ..gpu import get_array
from .measure import Measure
from .result import Result


def intensity_at_center(channel, output):
    iy, ix = channel.grid.origin_index
    return float(np.abs(get_array(output[iy, ix])) ** 2)


class SIResult(Result):
    def __init__(self, channel, theoretical_functions=(), *args, **kwargs):
        measures = [Measure(channel, "propagation", intensity_at_center)]
        super().__init__(*args, channel=channel, measures=measures, **kwargs)
        self.set_theoretical_functions(*theoretical_functions)

    def set_theoretical_functions(self, *theoretical_functions):
        self.theoretical_functions = theoretical_functions
        self.theoretical_si = [f(self.positions, self.channel.path.phase_screen.model, self.channel.source)
                               for f in theoretical_functions]

    @property
    def intensities_at_center(self):
        return np.asarray(self.measures[0])

    @property
    def positions(self):
        return np.array(list(self.channel.path.positions) + [self.channel.path.length])

    @property
    def si(self):
        i = self.intensities_at_center
        return (i**2).mean(axis=0) / i.mean(axis=0) ** 2 - 1

    def plot_output(self):
        from matplotlib import pyplot as plt
        plt.plot(self.positions, self.si, label=r"On-axis SI $\sigma_I$, m")
        for i, f in enumerate(self.theoretical_functions):
            plt.plot(self.positions, self.theoretical_si[i], label=f"Theoretical on-axis SI: {f.__name__}")
        plt.plot(np.nan, np.nan, label=f"Iterations: {len(self.measures[0])}", alpha=0)
        plt.xlabel("Propagation distance z, m")
        plt.legend()
        plt.show()
