"""Beam-wander / beam-width statistics.  Mirror of /root/reference/pyatmosphere/simulations/beam.py:14-80."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from ..measures import all_moments, mean_x, mean_x2, mean_xy, mean_y, mean_y2
from .measure import Measure
from .result import Result


class BeamResult(Result):
    def __init__(self, channel, **kwargs):
        measures = [
            Measure(channel, "atmosphere", mean_x, fast_key=("moment", "mean_x")),
            Measure(channel, "atmosphere", mean_y, fast_key=("moment", "mean_y")),
            Measure(channel, "atmosphere", mean_x2, fast_key=("moment", "mean_x2")),
            Measure(channel, "atmosphere", mean_xy, fast_key=("moment", "mean_xy")),
            Measure(channel, "atmosphere", mean_y2, fast_key=("moment", "mean_y2")),
            Measure(channel, "atmosphere", self.mean_x2_r, name="mean_x2_r", fast_key=("moment", "mean_x2_r")),
        ]
        super().__init__(channel, measures, **kwargs)

    def mean_x2_r(self, channel, output):
        """Second moment along the direction of the instantaneous centroid (beam.py:26-33):
        sum I (x cos(xi) + (-y) sin(xi))^2 delta^2 with (cos, sin)(xi) = (<x>, <y>)/r0 taken from this
        iteration's mean_x / mean_y.  Expanded, it is c^2<x^2> + 2cs<xy> + s^2<y^2> of the same sweep."""
        mx, my = self.measures[0].iteration_data, self.measures[1].iteration_data
        r0 = np.sqrt(mx**2 + my**2)
        c, s = mx / r0, my / r0
        mom = all_moments(channel, output)
        return float(c * c * mom["mean_x2"][0] + 2 * c * s * mom["mean_xy"][0] + s * s * mom["mean_y2"][0])

    @property
    def bw2(self) -> Sequence[float]:
        return np.asarray(self.measures[0]) ** 2

    @property
    def lt2(self) -> Sequence[float]:
        return 4 * np.asarray(self.measures[2])

    @property
    def st2(self) -> Sequence[float]:
        return self.lt2 - 4 * self.bw2

    @staticmethod
    def _root_with_error(v2) -> Tuple[float, float]:
        """sqrt(mean) and its standard error propagated through the square root (beam.py:49-71)."""
        mean = np.sqrt(v2.mean())
        err2 = v2.std(ddof=1) / np.sqrt(len(v2))
        return mean, err2 / 2 / mean

    @property
    def bw(self):
        return self._root_with_error(self.bw2)

    @property
    def lt(self):
        return self._root_with_error(self.lt2)

    @property
    def st(self):
        return self._root_with_error(self.st2)

    def print_output(self):
        bw, lt, st = self.bw, self.lt, self.st
        print(f"sigma_BW_x = {bw[0]:.1e} +- {bw[1]:.1e}")
        print(f"sigma_LT_x = {lt[0]:.1e} +- {lt[1]:.1e}")
        print(f"W_ST = {st[0]:.1e} +- {st[1]:.1e}")
        print(f"Count of measures: {len(self.measures[0])}")
