"""Container of the measures of one study + CSV checkpointing.
Mirror of /root/reference/pyatmosphere/simulations/result.py:5-43 (same on-disk format: one column per
measure, header = Measure.name, floats written as '%.3e')."""
from __future__ import annotations

import pandas as pd


class Result:
    save_float_format = '{:.3e}'.format

    def __init__(self, channel, measures, max_size=None, save_path: str = ""):
        self.channel = channel
        self.measures = measures
        self.set_max_size(max_size)
        self.save_path = save_path
        if self.save_path:
            try:
                self.load_output()
                print(f"Loaded measures from {self.save_path}")
            except FileNotFoundError:
                pass

    def set_max_size(self, max_size):
        for m in self.measures:
            m.max_size = max_size

    def print_output(self):
        print(f"Len of the first measures: {len(self.measures[0])}")

    def plot_output(self):
        self.print_output()

    def as_df(self):
        df = pd.DataFrame([m.data for m in self.measures]).T
        df.columns = [m.name for m in self.measures]
        return df

    def save_output(self):
        if not self.save_path:
            return
        self.as_df().to_csv(self.save_path, index=False, float_format=self.save_float_format)

    def load_output(self):
        for m, column in zip(self.measures, pd.read_csv(self.save_path).T.values):
            m.data = column.tolist()
