"""On-axis scintillation index per leg: record and statistic of
/root/reference/pyatmosphere/simulations/si.py:11-36.  The closed-form Andrews curves the reference overlays
(theory/atmosphere/si.py) are analytic post-processing outside this path; pass callables in
`theoretical_functions` to overlay them."""
from __future__ import annotations

import numpy as np

from ..gpu import get_array
from .measure import Measure
from .result import Result


def intensity_at_center(channel, output):
    """|u|^2 at the grid origin (one element read back from the device)."""
    row, col = channel.grid.origin_index
    centre = np.asarray(get_array(output[row, col]))
    return float(centre.real**2 + centre.imag**2)


class SIResult(Result):
    def __init__(self, channel, theoretical_functions=(), *args, **kwargs):
        record = Measure(channel, "propagation", intensity_at_center)
        super().__init__(*args, channel=channel, measures=[record], **kwargs)
        self.set_theoretical_functions(*theoretical_functions)

    def set_theoretical_functions(self, *theoretical_functions):
        model, source = self.channel.path.phase_screen.model, self.channel.source
        self.theoretical_functions = theoretical_functions
        self.theoretical_si = [f(self.positions, model, source) for f in theoretical_functions]

    @property
    def positions(self):
        """Screen positions followed by the end of the path: the planes at which the record is taken."""
        return np.append(np.asarray(self.channel.path.positions, dtype=float), self.channel.path.length)

    @property
    def intensities_at_center(self):
        return np.asarray(self.measures[0])

    @property
    def si(self):
        """<I^2>/<I>^2 - 1 per plane."""
        samples = self.intensities_at_center
        mean = samples.mean(axis=0)
        return np.mean(samples * samples, axis=0) / (mean * mean) - 1

    def plot_output(self):
        """Scintillation index against distance, with whichever closed-form curves were registered."""
        from matplotlib import pyplot as plt
        z = self.positions
        fig, axis = plt.subplots()
        axis.plot(z, self.si, marker="o", label=f"simulated on-axis scintillation index ({len(self.measures[0])} iterations)")
        for function, curve in zip(self.theoretical_functions, self.theoretical_si):
            axis.plot(z, curve, linestyle="--", label=f"theory: {function.__name__}")
        axis.set_xlabel("z, m")
        axis.set_ylabel("scintillation index")
        axis.legend()
        plt.show()
