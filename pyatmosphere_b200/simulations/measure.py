"""Per-sample record of one observable.  Mirror of /root/reference/pyatmosphere/simulations/measure.py:5-38."""
from __future__ import annotations

from typing import Sequence

import numpy as np


class Measure:
    def __init__(self, channel, measure_type: str, *operations, name: str = "", max_size: int = None,
                 time: Sequence[float] = None, save_path: str = None, save_name: str = None, fast_key=None):
        self.channel = channel
        self.measure_type = measure_type
        self.operations = tuple(operations)
        single = len(operations) == 1 and getattr(operations[0], "__name__", "<lambda>") != "<lambda>"
        self.name = name or (operations[0].__name__ if single else "")
        self.max_size = max_size
        self.time = tuple(time) if time else None
        self.data = []
        self.iteration_data = None
        # not in the reference: tells Simulation's batched device path which column of the fused measure table
        # this record takes (None -> only the generic per-iteration loop can serve it)
        self.fast_key = fast_key

    @property
    def is_done(self):
        return self.max_size is not None and len(self) >= self.max_size

    def __len__(self):
        return len(self.data)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.data)
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return self.name
