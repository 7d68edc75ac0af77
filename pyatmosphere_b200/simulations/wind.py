"""Frozen-flow time series: the same random spectrum re-evaluated on a grid shifted by the wind.
Mirror of /root/reference/pyatmosphere/simulations/wind.py:13-66.  `Measure(time=...)` makes Simulation.iter run the
channel once per time value with `shift=(0, t), wind=True`, i.e. with the spectrum cached on every screen
(phase_screens.py:93-106) -- on the device this is the same synthesis kernel with a coordinate offset."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pandas as pd

from ..measures import eta, mean_x, mean_y
from .measure import Measure
from .result import Result


class WindResult(Result):
    def load_output(self):
        for m, column in zip(self.measures, pd.read_csv(self.save_path).T.values):
            m.data = [[float(v) for v in row[1:-1].split(", ")] for row in column]


class TimeCoherenceResult(WindResult):
    def __init__(self, channel, time, *args, **kwargs):
        measures = [Measure(channel, "pupil", eta, time=time)]
        super().__init__(*args, channel=channel, measures=measures, **kwargs)

    @property
    def tc(self) -> Sequence[float]:
        from scipy.stats import pearsonr
        a = np.asarray(self.measures[0])
        return [pearsonr(a[:, 0], a[:, i])[0] for i in range(len(self.measures[0].time))]

    def plot_output(self):
        from matplotlib import pyplot as plt
        if len(self.measures[0]) > 2:
            plt.plot(self.measures[0].time, self.tc)
            plt.ylim((0, 1))
        plt.show()
        print(f"Iteration: {len(self.measures[0].data)}")


class TimeBWcorrSimulation(WindResult):
    def __init__(self, channel, time, *args, **kwargs):
        measures = [Measure(channel, "atmosphere", mean_x, time=time), Measure(channel, "atmosphere", mean_y, time=time)]
        super().__init__(*args, channel=channel, measures=measures, **kwargs)

    def _corr(self, a, b):
        a, b = np.asarray(self.measures[a]), np.asarray(self.measures[b])
        return (a[:, 0, None] * b[:, :]).mean(axis=0)

    @property
    def xx(self) -> Sequence[float]:
        return 2 * np.sqrt(self._corr(0, 0))

    @property
    def yy(self) -> Sequence[float]:
        return 2 * np.sqrt(self._corr(1, 1))

    @property
    def xy(self) -> Sequence[float]:
        return 2 * np.sqrt(abs(self._corr(0, 1)))

    def plot_output(self):
        from matplotlib import pyplot as plt
        plt.scatter(self.measures[0].time, self.xx)
        plt.ylabel(r"Beam wandering $2 \cdot \sqrt{\left<x_0 x_{\tau}\right>}$, m")
        plt.xlabel("Wind shift, m")
        plt.show()
        plt.scatter(self.measures[0].time, self.xy)
        plt.ylabel(r"Beam wandering $2 \cdot \sqrt{|\left<x_0 y_{\tau}\right>|}$, m")
        plt.xlabel("Wind shift, m")
        plt.show()
