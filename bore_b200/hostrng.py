"""The candidate points of an argmax: ``random_state.uniform(low, high, size=(n, dim))`` from the caller's
``numpy.random.RandomState`` -- the reference's call (bore/mixins.py:49, bore/optimizers/base.py), the same numbers
and the same generator state afterwards, produced by the library's own MT19937 (csrc/hostrng.cu) instead of
numpy's broadcasting path: 65,536 x 50 doubles took numpy longer than the GPU takes to train the classifier.

Anything that is not a legacy ``RandomState`` on MT19937 (a ``Generator``, a module-level ``np.random``, a
subclass with its own ``uniform``) keeps numpy's call.  Host code either way: this is not a device path.
"""
import ctypes as C

import numpy as np

from . import _lib

# below this many numbers numpy's own call is as fast as reading and writing the generator state
FAST_MIN = 1 << 14


def _is_plain_mt19937(rs):
    return type(rs) is np.random.RandomState


def uniform(random_state, low, high, n, dim, out=None):
    """``random_state.uniform(low=low, high=high, size=(n, dim))`` as float64, optionally into ``out`` (a
    C-contiguous float64 array of that shape, e.g. a view of a pinned staging buffer)."""
    low = np.ascontiguousarray(np.broadcast_to(np.asarray(low, np.float64), (dim,)))
    high = np.ascontiguousarray(np.broadcast_to(np.asarray(high, np.float64), (dim,)))
    if out is None:
        out = np.empty((n, dim), np.float64)
    assert out.shape == (n, dim) and out.dtype == np.float64 and out.flags.c_contiguous
    if n * dim < FAST_MIN or not _is_plain_mt19937(random_state):
        out[...] = random_state.uniform(low=low, high=high, size=(n, dim))
        return out
    name, key, pos, has_gauss, cached = random_state.get_state(legacy=True)
    if name != "MT19937":
        out[...] = random_state.uniform(low=low, high=high, size=(n, dim))
        return out
    key = np.ascontiguousarray(key, np.uint32).copy()
    cpos = C.c_int(int(pos))
    lib = _lib.load()
    _lib.check(lib.bore_mt19937_uniform(key.ctypes.data, C.byref(cpos), low.ctypes.data, high.ctypes.data,
                                        int(dim), int(n), out.ctypes.data))
    random_state.set_state((name, key, int(cpos.value), has_gauss, cached))
    return out
