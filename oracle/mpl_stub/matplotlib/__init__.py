"""Import shim so the read-only reference (which imports matplotlib at module import time,
channels.py:1, simulations/beam.py:3, pdt.py:2) can be imported on a box without matplotlib.
Only used by oracle/make_golden.py; nothing here ever plots."""
