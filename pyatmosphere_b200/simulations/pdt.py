"""Probability distribution of the transmittance.  Mirror of
/root/reference/pyatmosphere/simulations/pdt.py:11-66 (PDTResult, TrackedPDTResult)."""
from __future__ import annotations

from functools import partial

import numpy as np

from ..measures import eta, mean_x, mean_y
from .measure import Measure
from .result import Result


class PDTResult(Result):
    bins_single, bins_multi = 200, 100      # pdt.py:30-31,41-46

    def __init__(self, channel, pupils: list = None, **kwargs):
        self.pupil_shift = (0, 0)
        self.pupils = pupils or [channel.pupil]
        measures = kwargs.pop("measures", None)
        if measures is None:
            measures = [Measure(channel, "atmosphere", partial(self.append_pupil, pupil), eta, name=f"{pupil.radius}",
                                fast_key=("pupil_eta", pupil.radius, "fixed"))
                        for pupil in self.pupils]
        super().__init__(channel, measures, **kwargs)

    def append_pupil(self, pupil, channel, output):
        """Mask `output` with `pupil` at the current `pupil_shift` (pdt.py:21-26)."""
        saved = channel.pupil
        channel.pupil = pupil
        output = channel.pupil.output(output, shift=self.pupil_shift)
        channel.pupil = saved
        return output

    def histogram(self, index=0, bins=None):
        """Counts of the histogram `plot_output` draws: `bins` equal bins on [0, 1]."""
        bins = bins or (self.bins_single if len(self.pupils) == 1 else self.bins_multi)
        return np.histogram(np.asarray(self.pdt_measures()[index].data, dtype=np.float64), bins=bins, range=(0, 1))[0]

    def pdt_measures(self):
        return self.measures[-len(self.pupils):]

    def plot_output(self):
        """Transmittance histograms, one panel per aperture, drawn from `histogram()` (the counts the statistics use)."""
        from matplotlib import pyplot as plt
        records = self.pdt_measures()
        panels = len(records)
        columns = min(panels, 3)
        rows = -(-panels // columns)
        fig, axes = plt.subplots(rows, columns, figsize=(5 * columns, 3 * rows), squeeze=False)
        for index, axis in enumerate(axes.ravel()):
            if index >= panels:
                axis.set_visible(False)
                continue
            counts = self.histogram(index)
            edges = np.linspace(0.0, 1.0, len(counts) + 1)
            caption = f"Count: {len(records[index])}"
            if panels > 1:
                caption = f"Pupil radius: {self.pupils[index].radius:.3f}\n" + caption
            axis.stairs(counts, edges, fill=True, label=caption)
            axis.set_xlim(0.0, 1.0)
            axis.legend()
        plt.show()


class TrackedPDTResult(PDTResult):
    """Aperture re-centred on the instantaneous beam centroid (pdt.py:51-66)."""

    def __init__(self, channel, pupils: list = None, **kwargs):
        pupils = pupils or [channel.pupil]
        beam = [Measure(channel, "atmosphere", mean_x, fast_key=("moment", "mean_x")),
                Measure(channel, "atmosphere", mean_y, fast_key=("moment", "mean_y"))]
        pdt = [Measure(channel, "atmosphere", self.set_pupil_position, partial(self.append_pupil, pupil), eta,
                       name=f"{pupil.radius}", fast_key=("pupil_eta", pupil.radius, "tracked"))
               for pupil in pupils]
        super().__init__(channel, pupils=pupils, measures=beam + pdt, **kwargs)

    def set_pupil_position(self, channel, output):
        self.pupil_shift = (self.measures[0].iteration_data, self.measures[1].iteration_data)
        return output
