"""Closed forms imported by in-scope modules (/root/reference/pyatmosphere/theory/atmosphere/__init__.py:5-11)."""


def get_rytov2(Cn2, k, length):
    return 1.23 * Cn2 * k ** (7 / 6) * length ** (11 / 6)


def get_r0(Cn2, k, length):
    """Coherence diameter for plane waves."""
    return (0.423 * k**2 * Cn2 * length) ** (-3 / 5)
