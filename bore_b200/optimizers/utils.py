"""Bounds normalisation (mirror of bore/optimizers/utils.py:4-16)."""
from scipy.optimize import Bounds


def from_bounds(bounds):
    """``scipy.optimize.Bounds`` or a sequence of ``(low, high)`` pairs -> ``((low, high), dim)``."""
    if isinstance(bounds, Bounds):
        low, high = bounds.lb, bounds.ub
        dim = len(low)
        assert dim == len(high), "lower and upper bounds sizes do not match!"
        return (low, high), dim
    pairs = list(bounds)
    low = tuple(p[0] for p in pairs)
    high = tuple(p[1] for p in pairs)
    return (low, high), len(pairs)
