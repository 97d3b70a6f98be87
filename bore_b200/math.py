"""Integer helpers of the hot path's host side (mirror of bore/math.py:4-29)."""
import numpy as np


def ceil_divide(a, b, *args, **kwargs):
    """Ceiling division through floor division of the negated numerator (bore/math.py:4-5);
    extra arguments go to ``np.floor_divide`` like in the reference."""
    return np.negative(np.floor_divide(np.negative(a), b, *args, **kwargs))


def steps_per_epoch(dataset_size, batch_size):
    """Gradient steps in one pass over ``dataset_size`` samples; a trailing short batch still
    counts as a step (bore/math.py:8-29).

    The training lengths of the BASELINE configurations (README net on 10 .. 110 observations,
    Hartmann-6 on 500, Ackley-50 on 2,000; the reference's own doctest values are held in
    tests/golden/host_golden.json):

    >>> [steps_per_epoch(n, 64) for n in (10, 110, 500, 2000)]
    [1, 2, 8, 32]
    >>> steps_per_epoch(65, 64), steps_per_epoch(64, 64), steps_per_epoch(1, 1)
    (2, 1, 1)
    """
    return int(ceil_divide(dataset_size, batch_size))
