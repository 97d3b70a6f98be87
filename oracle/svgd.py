"""CPU oracle for the SVGD batch argmax.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of bore/optimizers/svgd/base.py:36-131 and bore/optimizers/svgd/kernels.py:4-28
(the reference code is itself plain NumPy, so the restatement uses the same array operations in
the same order) plus ``argmax_batch`` (bore/mixins.py:100-116) with the oracle MLP as the model.
**Pinned**: the reference's SVGD modules import in the build container (numpy / scipy / sklearn
only); tests/golden/svgd_golden.npz holds trajectories produced by the reference's own
``SVGD.optimize_from_init`` / ``RadialBasis.value_and_grad`` (tests/golden/make_golden.py), and
tests/test_svgd.py holds this file to them exactly.
"""
import numpy as np

from . import keras_mlp as km


def check_length_scale(n_samples, sum_sqr_diff, length_scale=None, eps=1e-6):
    """kernels.py:4-10: median heuristic when ``length_scale`` is None, floored at eps."""
    if length_scale is None:
        h = np.median(sum_sqr_diff)
        length_scale = np.sqrt(.5 * h / np.log(n_samples + 1))
    if eps is not None:
        length_scale = np.maximum(length_scale, eps)
    return length_scale


def rbf_value_and_grad(X, length_scale=1.0):
    """kernels.py:18-28."""
    n_samples = X.shape[0]
    diff = np.expand_dims(X, axis=1) - X
    sum_sqr_diff = np.sum(np.square(diff), axis=-1)
    length_scale = check_length_scale(n_samples, sum_sqr_diff, length_scale)
    gamma = .5 / length_scale**2
    K = np.exp(-gamma * sum_sqr_diff)
    K_grad = 2. * np.sum(gamma * diff * np.expand_dims(K, axis=-1), axis=1)
    return K, K_grad


def rank(a):
    """svgd/base.py:36-64: empirical CDF, ties counted weakly."""
    return np.less_equal(a, np.expand_dims(a, axis=1)).mean(axis=1)


def optimize_from_init(func, x_init, bounds=None, length_scale=1.0, n_iter=1000, step_size=1e-3,
                       alpha=.9, eps=1e-6, tau=1., lambd=None, c=1., callback=None):
    """svgd/base.py:78-118; ``lambd=None`` = DistortionConstant(c), else DistortionExpDecay."""
    if bounds is not None:
        low = np.array([b[0] for b in bounds], np.float64)
        high = np.array([b[1] for b in bounds], np.float64)
    n_init = x_init.shape[0]
    grad_hist = None
    x = x_init.copy()
    for i in range(n_iter):
        K, K_grad = rbf_value_and_grad(x, length_scale)
        f, f_grad = func(x)
        zeta = c if lambd is None else np.power(rank(f), -lambd)
        Zeta = np.expand_dims(zeta, axis=-1)
        grad = (K @ (Zeta * f_grad) + tau * K_grad)
        grad /= n_init
        if grad_hist is None:
            grad_hist = grad**2
        else:
            grad_hist *= alpha
            grad_hist += (1 - alpha) * grad**2
        adj_grad = np.true_divide(grad, eps + np.sqrt(grad_hist))
        x += step_size * adj_grad
        if bounds is not None:
            x = x.clip(low, high)
        if callback is not None:
            callback(x)
    return x


def make_func_max(weights, acts, transform="identity", dtype=np.float32):
    """``self._func_max`` (bore/mixins.py:98): value and input-gradient of transform(model(x)) for a
    BATCH of particles -- value (n,) in the model dtype, gradient (n, D) as float64 (it keeps
    the dtype of x, bore/decorators.py:54-56)."""
    def func(x):
        f, g = km.value_and_input_grad(weights, acts, np.atleast_2d(x), transform, False, dtype)
        return f, g.astype(np.float64)
    return func


def argmax_batch(weights, acts, batch_size, bounds, transform="identity", length_scale=None, n_iter=1000,
                 step_size=1e-3, alpha=.9, eps=1e-6, tau=1.0, lambd=None, random_state=None):
    """bore/mixins.py:100-116 + svgd/base.py:120-131 (uniform start points from random_state)."""
    rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
    low = [b[0] for b in bounds]
    high = [b[1] for b in bounds]
    x_init = rs.uniform(low=low, high=high, size=(batch_size, len(bounds)))
    return optimize_from_init(make_func_max(weights, acts, transform), x_init, bounds, length_scale, n_iter,
                              step_size, alpha, eps, tau, lambd)
