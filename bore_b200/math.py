"""Integer helpers of the hot path's host side (mirror of bore/math.py:4-29)."""
import numpy as np


def ceil_divide(a, b, *args, **kwargs):
    """Ceiling division through floor division of the negated numerator (bore/math.py:4-5);
    extra arguments go to ``np.floor_divide`` like in the reference."""
    return np.negative(np.floor_divide(np.negative(a), b, *args, **kwargs))


def steps_per_epoch(dataset_size, batch_size):
    """Gradient steps in one pass over ``dataset_size`` samples; a trailing short batch still
    counts as a step (bore/math.py:8-29).

    >>> steps_per_epoch(dataset_size=32, batch_size=64)
    1
    >>> steps_per_epoch(dataset_size=64, batch_size=64)
    1
    >>> steps_per_epoch(dataset_size=100, batch_size=64)
    2
    >>> steps_per_epoch(dataset_size=1000, batch_size=64)
    16
    """
    return int(ceil_divide(dataset_size, batch_size))
