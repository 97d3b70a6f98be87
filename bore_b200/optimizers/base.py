"""Generic multi-start minimiser (mirror of bore/optimizers/base.py:8-67).

``minimize_multi_start(fn, bounds, num_starts, num_samples=None, random_state=None, ...)``
keeps the reference's contract -- ``fn`` maps ``x`` to ``(value, gradient)`` and also accepts a
batch -- but the L-BFGS-B iterations of ALL starts run on the GPU:

* ``fn`` built by ``bore_b200.convert`` : screening (K2), sort (K4) and the batched L-BFGS-B
  with the MLP objective inlined (K3) never leave the device;
* any other callable: the on-device L-BFGS-B stepper drives ``fn`` through reverse
  communication (one batched round of evaluations per lock-step), so SciPy is still not on
  the path.

The reference marks this function "Deprecated until minor bug fixed" (base.py:7): it never
defaults ``jac=True``.  Here ``jac`` defaults to True; passing a falsy ``jac`` still asserts.
"""
import ctypes as C

import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.utils import check_random_state

from .utils import from_bounds
from .. import hostrng
from ..engine import lbfgsb_message


def _device_stepper_minimize(fn, X0, low, high, options):
    """Reverse-communication loop around bore_lbfgsb_init/step (objective on the host)."""
    import torch
    from .. import _lib
    lib = _lib.require_cuda()
    S, n = X0.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    m = int(options.get("maxcor", 10))
    nbytes = lib.bore_lbfgsb_workspace_bytes(S, n, m)
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    xreq = torch.empty(S, n, dtype=torch.float64, device=dev)
    pend = torch.empty(S, dtype=torch.int32, device=dev)
    X0d = torch.from_numpy(np.ascontiguousarray(X0, np.float64)).to(dev)
    P = lambda t: C.c_void_p(t.data_ptr())
    NP = lambda a: a.ctypes.data_as(C.c_void_p)
    lo = np.ascontiguousarray(low, np.float64)
    hi = np.ascontiguousarray(high, np.float64)
    _lib.check(lib.bore_lbfgsb_init(P(X0d), S, n, NP(lo), NP(hi), m,
                                    float(options.get("ftol", 2.2204460492503131e-09)),
                                    float(options.get("gtol", 1e-5)),
                                    int(options.get("maxiter", 15000)),
                                    int(options.get("maxfun", 15000)),
                                    int(options.get("maxls", 20)), P(work), nbytes, P(xreq),
                                    P(pend), dev.index, None))
    fh = np.zeros(S)
    gh = np.zeros((S, n))
    fd = torch.zeros(S, dtype=torch.float64, device=dev)
    gd = torch.zeros(S, n, dtype=torch.float64, device=dev)
    pending = C.c_int(S)
    while pending.value > 0:
        xr = xreq.cpu().numpy()
        for i in np.flatnonzero(pend.cpu().numpy()):
            f, g = fn(xr[i])
            fh[i] = float(np.asarray(f))
            gh[i] = np.asarray(g, np.float64)
        fd.copy_(torch.from_numpy(fh))
        gd.copy_(torch.from_numpy(gh))
        _lib.check(lib.bore_lbfgsb_step(P(fd), P(gd), 1, S, n, P(work), P(xreq), P(pend),
                                        C.byref(pending), dev.index, None))
    x = torch.empty(S, n, dtype=torch.float64, device=dev)
    fun = torch.empty(S, dtype=torch.float64, device=dev)
    ints = torch.empty(4, S, dtype=torch.int32, device=dev)
    _lib.check(lib.bore_lbfgsb_results(S, n, P(work), P(x), P(fun), P(ints[0]), P(ints[1]),
                                       P(ints[2]), P(ints[3]), dev.index, None))
    return dict(x=x.cpu().numpy(), fun=fun.cpu().numpy(), nit=ints[0].cpu().numpy(),
                nfev=ints[1].cpu().numpy(), status=ints[2].cpu().numpy(),
                task=ints[3].cpu().numpy())


def multi_start(minimizer_fn=None):
    """Decorator form kept for surface parity (bore/optimizers/base.py:8).  ``minimizer_fn`` is
    accepted for signature compatibility only: the minimiser is the device L-BFGS-B."""

    def new_minimizer(fn, bounds, num_starts, num_samples=None, random_state=None, *args,
                      **kwargs):
        random_state = check_random_state(random_state)

        assert "x0" not in kwargs, "`x0` should not be specified"
        assert "jac" not in kwargs or kwargs["jac"], "`jac` must be true"
        method = kwargs.pop("method", "L-BFGS-B")
        kwargs.pop("jac", None)
        options = dict(kwargs.pop("options", None) or {})
        if method != "L-BFGS-B":
            raise NotImplementedError(f"method={method!r}: only L-BFGS-B has a device path")
        if args or kwargs:
            raise TypeError(f"unsupported minimizer arguments: {args} {sorted(kwargs)}")

        if num_samples is None:
            num_samples = num_starts

        assert num_samples >= num_starts, \
            "number of random samples (`num_samples`) must be " \
            "greater than number of starting points (`num_starts`)"

        (low, high), dims = from_bounds(bounds)
        low = np.asarray(low, np.float64)
        high = np.asarray(high, np.float64)
        X_init = hostrng.uniform(random_state, low, high, num_samples, dims)  # numpy's stream, faster

        values, _ = fn(X_init)  # batched screening call (bore/optimizers/base.py:53)
        ind = np.argsort(np.asarray(values), kind="stable")
        X0 = X_init[ind[:num_starts]]

        model = getattr(fn, "_bore_model", None)
        if model is not None:
            from .. import ops
            e = fn._bore_transform(ops.Expr(model, (1,), (1,)))
            if e.sign != -1 and e.sign != 1:
                raise NotImplementedError
            net = model._engine(dims)
            if e.sign == -1:
                res = net.lbfgsb(X0, low, high, transform=e.transform,
                                 m=options.get("maxcor", 10),
                                 ftol=options.get("ftol", 2.2204460492503131e-09),
                                 gtol=options.get("gtol", 1e-5),
                                 maxiter=options.get("maxiter", 15000),
                                 maxfun=options.get("maxfun", 15000),
                                 maxls=options.get("maxls", 20))
            else:  # T(+u): no fused objective for that sign -> reverse communication with K2
                res = _device_stepper_minimize(fn, X0, low, high, options)
        else:
            res = _device_stepper_minimize(fn, X0, low, high, options)

        results = []
        for i in range(num_starts):
            st = int(res["status"][i])
            results.append(OptimizeResult(x=res["x"][i], fun=res["fun"][i], nit=int(res["nit"][i]),
                                          nfev=int(res["nfev"][i]), njev=int(res["nfev"][i]),
                                          status=st, success=(st == 0),
                                          message=lbfgsb_message(st, res["task"][i])))
        return results

    return new_minimizer


minimize_multi_start = multi_start()
