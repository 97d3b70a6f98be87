"""Stein variational gradient descent, the reference's batch argmax optimiser (mirror of
bore/optimizers/svgd/base.py:11-131): same classes, arguments and RNG consumption.

Every iteration runs on the GPU (``svgd_step_kernel``, csrc/svgd.cu).  When ``func`` is the
value-and-gradient closure of a ``bore_b200`` model (what ``convert`` returns -- the only ``func``
the reference itself passes, bore/mixins.py:98,115) and no ``callback`` is given, the whole
``n_iter`` loop is enqueued at once (MLP kernel + step kernel per iteration, no host round trip).
Any other callable is the caller's own objective: it is evaluated on the host each iteration, as
it must be, and only the SVGD arithmetic runs on the device.
"""
from abc import ABC, abstractmethod

import numpy as np
from sklearn.utils import check_random_state

from .kernels import RadialBasis
from ..utils import from_bounds
from ... import engine


class Distortion(ABC):

    @abstractmethod
    def __call__(self, beta):
        pass


class DistortionConstant(Distortion):

    def __init__(self, c=1.):
        self.c = c

    def __call__(self, beta):
        return self.c


class DistortionExpDecay(Distortion):
    """Importance weight ``omega(beta) = beta ** -lambd`` of the rank ``beta``."""

    def __init__(self, lambd=1.):
        self.lambd = lambd

    def __call__(self, beta):
        return np.power(beta, -self.lambd)


def rank(a):
    """Empirical CDF of the entries of a 1-d array (bore/optimizers/svgd/base.py:36-64).

    >>> rank(np.array([0.4532752, 0.858725 , 0.3792093, 0.6631048, 0.7619765]))
    array([0.4, 1. , 0.2, 0.6, 0.8])
    >>> rank(np.array([0.4532752, 0.858725 , 0.3792093, 0.3792093, 0.7619765]))
    array([0.6, 1. , 0.4, 0.4, 0.8])
    """
    assert a.ndim == 1, "only support 1d arrays!"
    return np.less_equal(a, np.expand_dims(a, axis=1)).mean(axis=1)


class SVGD:

    def __init__(self, kernel=RadialBasis(), n_iter=1000, step_size=1e-3,
                 alpha=.9, eps=1e-6, tau=1., distortion=DistortionConstant()):
        self.kernel = kernel
        self.n_iter = n_iter
        self.step_size = step_size
        self.alpha = alpha
        self.eps = eps
        self.tau = tau
        self.distortion = distortion

    def _device_options(self):
        if not isinstance(self.kernel, RadialBasis):
            raise NotImplementedError("only the RadialBasis kernel has a device path")
        ls = self.kernel.length_scale
        opts = dict(length_scale=float("nan") if ls is None else float(ls), step_size=self.step_size,
                    alpha=self.alpha, eps=self.eps, tau=self.tau, lambd=float("nan"), zeta_c=1.0)
        if isinstance(self.distortion, DistortionExpDecay):
            opts["lambd"] = float(self.distortion.lambd)
        elif isinstance(self.distortion, DistortionConstant):
            opts["zeta_c"] = float(self.distortion.c)
        else:
            raise NotImplementedError("distortion must be DistortionConstant or DistortionExpDecay")
        return opts

    def optimize_from_init(self, func, x_init, bounds=None, callback=None):
        """Optimize from specified starting points (svgd/base.py:78-118)."""
        low = high = None
        if bounds is not None:
            (low, high), _ = from_bounds(bounds)
        x_init = np.asarray(x_init, np.float64)
        opts = self._device_options()
        model = getattr(func, "_bore_model", None)
        if model is not None and callback is None:
            return model._native_svgd(x_init, func._bore_transform, low, high, self.n_iter, opts)
        state = engine.SvgdStepper(x_init, low, high, **opts)
        for i in range(self.n_iter):
            x = state.x()
            f, f_grad = func(x)
            state.step(np.asarray(f, np.float64), np.asarray(f_grad, np.float64))
            if callback is not None:
                callback(state.x())
        return state.x()

    def optimize(self, func, batch_size, bounds=None, callback=None, random_state=None):
        """Optimize from ``batch_size`` uniformly sampled starting points (svgd/base.py:120-131)."""
        random_state = check_random_state(random_state)
        (low, high), dims = from_bounds(bounds)
        x_init = random_state.uniform(low=low, high=high, size=(batch_size, dims))
        return self.optimize_from_init(func, x_init, bounds, callback)
