"""Stub pyplot: any attribute is a function that raises if it is actually called."""


def __getattr__(name):
    def _no_plot(*args, **kwargs):
        raise RuntimeError(f"matplotlib.pyplot.{name} called on the import shim (plotting is out of scope)")
    return _no_plot
