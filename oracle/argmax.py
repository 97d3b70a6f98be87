"""CPU restatement of the reference's multi-start argmax.  TEST INFRASTRUCTURE ONLY.

Follows bore/mixins.py:16-89 line by line (control flow, ``argpartition``, filter rule,
``num_starts == 0`` shortcut) and bore/optimizers/base.py:10-62, with the Keras model
replaced by ``oracle.keras_mlp`` (parity unpinned, see that file) and the optimiser being
the *installed* ``scipy.optimize.minimize(method="L-BFGS-B")`` -- the real thing, so the
L-BFGS-B half of this oracle is pinned to SciPy itself (version recorded in fixtures).

``LockstepLBFGSB`` drives SciPy's raw reverse-communication step ``_lbfgsb.setulb`` for S
independent starts at once: it is bit-identical to calling
``minimize`` per start but exposes per-request trial points, which is what the CUDA
stepper is compared against request by request.
"""
import numpy as np
from scipy.optimize import minimize, OptimizeResult, Bounds

from . import keras_mlp as km


def from_bounds(bounds):
    """bore/optimizers/utils.py:4-16."""
    if isinstance(bounds, Bounds):
        low, high = bounds.lb, bounds.ub
        dim = len(low)
        assert dim == len(high), "lower and upper bounds sizes do not match!"
    else:
        low, high = zip(*bounds)
        dim = len(bounds)
    return (low, high), dim


def make_func_min(weights, acts, transform="identity", dtype=np.float32, counter=None):
    """``convert(model, lambda u: transform(-u))`` (bore/mixins.py:20, bore/base.py:35-42):
    x:(D,) f64 -> [f:(), g:(D,)] as a LIST of arrays (bore/decorators.py:75).  The value is
    an fp32 scalar array and the gradient is fp64-typed/fp32-accurate (decorators.py:54-56:
    x stays fp64, Keras casts to fp32 inside the model)."""
    def fn(x):
        if counter is not None:
            counter[0] += 1
        x = np.asarray(x)
        f, g = km.value_and_input_grad(weights, acts, x.reshape(-1, x.shape[-1]),
                                       transform, True, dtype)
        if x.ndim == 1:
            return [f[0], g[0].astype(np.float64)]
        # batched call (bore/optimizers/base.py:53): the tape returns d(sum f)/dX
        return [f, g.astype(np.float64)]
    return fn


def maxima(weights, acts, bounds, num_starts=5, num_samples=1024, method="L-BFGS-B",
           options=dict(maxiter=1000, ftol=1e-9), print_fn=print, random_state=None,
           transform="identity", dtype=np.float32, counter=None):
    """bore/mixins.py:22-72."""
    from sklearn.utils import check_random_state
    random_state = check_random_state(random_state)
    assert num_samples is not None, "`num_samples` must be specified!"
    assert num_samples > 0, "`num_samples` must be positive integer!"
    assert num_starts is not None, "`num_starts` must be specified!"
    assert num_starts >= 0, "`num_starts` must be nonnegative integer!"
    assert num_samples >= num_starts

    (low, high), dim = from_bounds(bounds)
    X_init = random_state.uniform(low=low, high=high, size=(num_samples, dim))
    z_init = km.predict(weights, acts, X_init, dtype).squeeze(axis=-1)
    f_init = -z_init  # raw model output, NO transform (mixins.py:50-52)

    func_min = make_func_min(weights, acts, transform, dtype, counter)
    results = []
    if num_starts > 0:
        ind = np.argpartition(f_init, kth=num_starts - 1, axis=None)
        for i in range(num_starts):
            x0 = X_init[ind[i]]
            result = minimize(func_min, x0=x0, method=method, jac=True, bounds=bounds,
                              options=options)
            results.append(result)
            print_fn(f"[Maximum {i+1:02d}: value={result.fun:.3f}] "
                     f"success: {result.success}, "
                     f"iterations: {result.nit:02d}, "
                     f"status: {result.status} ({result.message})")
    else:
        i = np.argmin(f_init, axis=None)
        results.append(OptimizeResult(x=X_init[i], fun=f_init[i], success=True))
    return results


def argmax(weights, acts, bounds, filter_fn=lambda res: True, *args, **kwargs):
    """bore/mixins.py:74-89: first minimum of ``fun`` over results that
    ``(success or status == 1) and filter_fn(res)``; None if none qualify."""
    res_best = None
    for res in maxima(weights, acts, bounds, *args, **kwargs):
        if (res.success or res.status == 1) and filter_fn(res):
            if res_best is None or res.fun < res_best.fun:
                res_best = res
    return res_best


def minimize_starts(weights, acts, X0, bounds, options=dict(maxiter=1000, ftol=1e-9),
                    transform="identity", dtype=np.float32):
    """One ``scipy.optimize.minimize`` per row of X0, serial (the loop of
    bore/mixins.py:57-61) -> dict of arrays x, fun, nit, nfev, status."""
    func_min = make_func_min(weights, acts, transform, dtype)
    S, D = X0.shape
    out = dict(x=np.zeros((S, D)), fun=np.zeros(S), nit=np.zeros(S, np.int32),
               nfev=np.zeros(S, np.int32), status=np.zeros(S, np.int32))
    for i in range(S):
        r = minimize(func_min, x0=X0[i], method="L-BFGS-B", jac=True, bounds=bounds,
                     options=options)
        out["x"][i], out["fun"][i] = r.x, r.fun
        out["nit"][i], out["nfev"][i], out["status"][i] = r.nit, r.nfev, r.status
    return out


class LockstepLBFGSB:
    """S independent SciPy L-BFGS-B states advanced in lockstep through ``setulb``.

    Usage::

        ls = LockstepLBFGSB(X0, lo, hi, maxiter=1000, ftol=1e-9)
        while ls.pending.any():
            f, g = fun(ls.X[ls.pending])           # batched evaluation
            ls.feed(f, g)                          # advance every pending start
        ls.result() -> dict(x, fun, nit, nfev, status, task)

    Mirrors the driver loop of scipy/optimize/_lbfgsb_py.py:406-443 exactly, including the
    x0 clip (:359), the nfev accounting of ``ScalarFunction`` (first evaluation at x0 is
    shared with FG_START), ``maxiter`` -> 504 and ``nfev > maxfun`` -> 502.
    """

    def __init__(self, X0, lo, hi, m=10, maxiter=1000, ftol=1e-9, gtol=1e-5,
                 maxfun=15000, maxls=20):
        from scipy.optimize import _lbfgsb
        self._setulb = _lbfgsb.setulb
        X0 = np.asarray(X0, np.float64)
        S, n = X0.shape
        self.S, self.n, self.m = S, n, m
        self.maxiter, self.maxfun, self.maxls = maxiter, maxfun, maxls
        self.factr = ftol / np.finfo(float).eps
        self.pgtol = gtol
        lo = np.asarray(lo, np.float64)
        hi = np.asarray(hi, np.float64)
        self.nbd = np.zeros(n, np.int32)
        self.lo = np.where(np.isinf(lo), 0.0, lo)
        self.hi = np.where(np.isinf(hi), 0.0, hi)
        for i in range(n):
            L, U = not np.isinf(lo[i]), not np.isinf(hi[i])
            self.nbd[i] = {(False, False): 0, (True, False): 1, (True, True): 2,
                           (False, True): 3}[(L, U)]
        self.X = np.clip(X0, lo, hi)
        self.f = np.zeros(S)
        self.G = np.zeros((S, n))
        self.wa = [np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m) for _ in range(S)]
        self.iwa = [np.zeros(3 * n, np.int32) for _ in range(S)]
        self.task = [np.zeros(2, np.int32) for _ in range(S)]
        self.ln_task = [np.zeros(2, np.int32) for _ in range(S)]
        self.lsave = [np.zeros(4, np.int32) for _ in range(S)]
        self.isave = [np.zeros(44, np.int32) for _ in range(S)]
        self.dsave = [np.zeros(29) for _ in range(S)]
        self.nit = np.zeros(S, np.int32)
        self.nfev = np.zeros(S, np.int32)
        self.pending = np.zeros(S, bool)
        self.done = np.zeros(S, bool)
        # ScalarFunction memoises on x: a request at the point evaluated last is served
        # from its cache and does NOT bump nfev (scipy/optimize/_differentiable_functions.py)
        self._x_last = np.full((S, n), np.nan)
        for i in range(S):
            self._advance(i)

    def _advance(self, i):
        """Run start i until it asks for f,g or terminates."""
        while True:
            x = self.X[i]
            f = np.array(self.f[i])
            g = self.G[i]
            self._setulb(self.m, x, self.lo, self.hi, self.nbd, f, g, self.factr,
                         self.pgtol, self.wa[i], self.iwa[i], self.task[i], self.lsave[i],
                         self.isave[i], self.dsave[i], self.maxls, self.ln_task[i])
            t = self.task[i]
            if t[0] == 3:
                self.pending[i] = True
                return
            elif t[0] == 1:
                self.nit[i] += 1
                if self.nit[i] >= self.maxiter:
                    t[0], t[1] = 5, 504
                elif self.nfev[i] > self.maxfun:
                    t[0], t[1] = 5, 502
            else:
                self.pending[i] = False
                self.done[i] = True
                return

    def feed(self, f, G):
        """f (P,), G (P, n) for the currently pending starts, in index order."""
        idx = np.flatnonzero(self.pending)
        for k, i in enumerate(idx):
            self.f[i] = f[k]
            self.G[i] = np.asarray(G[k], np.float64)
            if not np.array_equal(self.X[i], self._x_last[i]):
                self.nfev[i] += 1
                self._x_last[i] = self.X[i]
            self.pending[i] = False
            self._advance(i)

    def result(self):
        status = np.zeros(self.S, np.int32)
        for i in range(self.S):
            if self.task[i][0] == 4:
                status[i] = 0
            elif self.nfev[i] > self.maxfun or self.nit[i] >= self.maxiter:
                status[i] = 1
            else:
                status[i] = 2
        return dict(x=self.X.copy(), fun=self.f.copy(), nit=self.nit.copy(),
                    nfev=self.nfev.copy(), status=status,
                    task=np.array([t.copy() for t in self.task]))
