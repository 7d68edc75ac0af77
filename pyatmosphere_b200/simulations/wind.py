"""Frozen-flow time series: the same random spectrum re-evaluated on a grid shifted by the wind
(/root/reference/pyatmosphere/simulations/wind.py:13-66).  `Measure(time=...)` makes Simulation.iter run the channel
once per time value with `shift=(0, t), wind=True`, i.e. with the spectrum cached on every screen
(phase_screens.py:93-106) -- on the device this is the same synthesis kernel with a coordinate offset."""
from __future__ import annotations

import ast
from typing import Sequence

import numpy as np
import pandas as pd

from ..measures import eta, mean_x, mean_y
from .measure import Measure
from .result import Result


class WindResult(Result):
    """Records are lists (one value per time lag); the CSV holds their repr, parsed back on load."""

    def load_output(self):
        table = pd.read_csv(self.save_path)
        for record, (_, column) in zip(self.measures, table.items()):
            record.data = [[float(v) for v in ast.literal_eval(cell)] for cell in column]


class TimeCoherenceResult(WindResult):
    """Transmittance behind the channel's aperture at every time lag; `tc` = its correlation with lag 0."""

    def __init__(self, channel, time, *args, **kwargs):
        super().__init__(*args, channel=channel, measures=[Measure(channel, "pupil", eta, time=time)], **kwargs)

    @property
    def tc(self) -> Sequence[float]:
        from scipy.stats import pearsonr
        samples = np.asarray(self.measures[0])
        return [pearsonr(samples[:, 0], samples[:, lag])[0] for lag in range(samples.shape[1])]

    def plot_output(self):
        from matplotlib import pyplot as plt
        if len(self.measures[0]) > 2:
            plt.plot(self.measures[0].time, self.tc)
            plt.ylim((0, 1))
        plt.show()
        print(f"Iteration: {len(self.measures[0].data)}")


class TimeBWcorrSimulation(WindResult):
    """Beam centroid at every time lag; xx / yy / xy = 2 sqrt(|<a_0 b_tau>|) cross-correlations."""

    def __init__(self, channel, time, *args, **kwargs):
        records = [Measure(channel, "atmosphere", op, time=time) for op in (mean_x, mean_y)]
        super().__init__(*args, channel=channel, measures=records, **kwargs)

    def _lagged(self, first, second):
        a, b = np.asarray(self.measures[first]), np.asarray(self.measures[second])
        return np.mean(a[:, :1] * b, axis=0)

    @property
    def xx(self) -> Sequence[float]:
        return 2 * np.sqrt(self._lagged(0, 0))

    @property
    def yy(self) -> Sequence[float]:
        return 2 * np.sqrt(self._lagged(1, 1))

    @property
    def xy(self) -> Sequence[float]:
        return 2 * np.sqrt(np.abs(self._lagged(0, 1)))

    def plot_output(self):
        from matplotlib import pyplot as plt
        lags = self.measures[0].time
        for values, label in ((self.xx, r"$2 \cdot \sqrt{\left<x_0 x_{\tau}\right>}$"), (self.xy, r"$2 \cdot \sqrt{|\left<x_0 y_{\tau}\right>|}$")):
            plt.scatter(lags, values)
            plt.ylabel("Beam wandering " + label + ", m")
            plt.xlabel("Wind shift, m")
            plt.show()
