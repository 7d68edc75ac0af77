"""Container of the measures of one study, with CSV checkpointing.

Behavioural mirror of /root/reference/pyatmosphere/simulations/result.py:5-43.  On-disk format (unchanged, it is the
reference's checkpoint format): one CSV column per measure, header = `Measure.name`, floats written with
'%.3e'; a Result constructed with an existing `save_path` resumes from it."""
from __future__ import annotations

import os

import pandas as pd


def _three_significant(value):
    return "%.3e" % value


class Result:
    save_float_format = staticmethod(_three_significant)

    def __init__(self, channel, measures, max_size=None, save_path: str = ""):
        self.channel, self.measures, self.save_path = channel, measures, save_path
        self.set_max_size(max_size)
        if save_path and os.path.exists(save_path):
            self.load_output()
            print(f"Loaded measures from {save_path}")

    # ---- bookkeeping --------------------------------------------------------------------------------------
    def set_max_size(self, max_size):
        for record in self.measures:
            record.max_size = max_size

    def as_df(self):
        """One column per measure (shorter records are padded with NaN by pandas)."""
        return pd.DataFrame({i: pd.Series(r.data) for i, r in enumerate(self.measures)}).set_axis(
            [r.name for r in self.measures], axis=1)

    # ---- reporting ---------------------------------------------------------------------------------------
    def print_output(self):
        print(f"Len of the first measures: {len(self.measures[0])}")

    def plot_output(self):
        self.print_output()

    # ---- checkpoint ----------------------------------------------------------------------------------------
    def save_output(self):
        if self.save_path:
            self.as_df().to_csv(self.save_path, index=False, float_format=self.save_float_format)

    def load_output(self):
        table = pd.read_csv(self.save_path)
        for record, (_, column) in zip(self.measures, table.items()):
            record.data = column.tolist()
