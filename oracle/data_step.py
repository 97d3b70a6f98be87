"""CPU oracle for the data step either side of the path.  TEST INFRASTRUCTURE ONLY.

Restates bore/data.py:31-35 (quantile labelling) and bore/data.py:42-48 (duplicate test) for
many problems at once.  The arithmetic lives in NumPy (`np.quantile`, method "linear";
`np.allclose`), which is what the reference calls -- so this oracle IS the reference's
arithmetic, and it is additionally **pinned** to outputs of the reference's own
`bore.data.Record` (tests/golden/data_step_golden.npz, written by tests/golden/make_golden.py
with /root/reference imported).
"""
import numpy as np


def quantile_labels(y, gamma):
    """y (M, N) -> (z bool (M, N), tau (M,)): bore/data.py:33-34 row by row."""
    y = np.atleast_2d(np.asarray(y, np.float64))
    tau = np.array([np.quantile(row, q=gamma) for row in y])
    with np.errstate(invalid="ignore"):
        z = np.stack([np.less(row, t) for row, t in zip(y, tau)])
    return z, tau


def lerp_quantile(y, gamma):
    """The same threshold spelled out (numpy/lib/_function_base_impl.py::_quantile, "linear"):
    what the CUDA kernel implements -- kept here so the formula itself is pinned to np.quantile."""
    s = np.sort(np.asarray(y, np.float64))
    n = s.size
    if np.isnan(s[-1]):
        return np.nan
    vi = (n - 1) * np.float64(gamma)
    if vi >= n - 1:
        return s[-1]
    if vi < 0:
        return s[0]
    prev = int(np.floor(vi))
    t = vi - prev
    a, b = s[prev], s[prev + 1]
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def is_duplicate(x, x_prev, rtol=1e-5, atol=1e-8):
    """x (G, K, D) candidates, x_prev (G, N, D) stored rows -> bool (G, K): bore/data.py:42-48."""
    x, x_prev = np.asarray(x, np.float64), np.asarray(x_prev, np.float64)
    out = np.zeros(x.shape[:2], bool)
    for g in range(x.shape[0]):
        for k in range(x.shape[1]):
            out[g, k] = any(np.allclose(p, x[g, k], rtol=rtol, atol=atol) for p in x_prev[g])
    return out
