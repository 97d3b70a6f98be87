"""Multi-start maximisation of a model's output wrt its input (mirror of bore/mixins.py:14-89).

Same signatures, assertions, RNG consumption and selection rules as the reference; the work
moves to the device:

    X_init (host MT19937, as the reference)  ->  K0 predict  ->  K4 top-k  ->  K3 batched
    L-BFGS-B with K2 inlined as the objective  ->  K4 first-minimum selection

``BatchMaximizableMixin`` (bore/mixins.py:92-116) adds the SVGD batch argmax, whose iterations
run on the device as well (csrc/svgd.cu).
"""
import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.utils import check_random_state

from . import ops
from .base import convert
from .engine import lbfgsb_message
from .optimizers.utils import from_bounds


def _accept_all(res):
    return True


def _bounds_arrays(bounds):
    (low, high), dim = from_bounds(bounds)
    return np.asarray(low, np.float64), np.asarray(high, np.float64), dim


class MaximizableMixin:

    def __init__(self, transform=ops.identity, *args, **kwargs):
        super(MaximizableMixin, self).__init__(*args, **kwargs)
        # negate to turn into a minimisation problem (bore/mixins.py:18-20): note T(-u), not -T(u)
        self._transform_fn = transform
        self._func_min = convert(self, transform=lambda u: transform(-u))

    # ------------------------------------------------------------------ device pipeline
    def _min_transform_name(self):
        """Trace ``transform(-u)`` once to learn which device transform code it is."""
        e = self._transform_fn(-ops.Expr(self, (1,), (1,)))
        if not isinstance(e, ops.Expr) or e.sign != -1:
            raise NotImplementedError("transform must be one of bore_b200.ops.identity/sigmoid/exp")
        return e.transform

    def _maxima_device(self, bounds, num_starts, num_samples, method, options, random_state):
        """Runs the whole screening + multi-start minimisation on the GPU.  Returns
        (X_init host array, dict of device tensors) -- or (X_init, None, i, f_i) when
        ``num_starts == 0``."""
        import torch
        random_state = check_random_state(random_state)

        assert num_samples is not None, "`num_samples` must be specified!"
        assert num_samples > 0, "`num_samples` must be positive integer!"
        assert num_starts is not None, "`num_starts` must be specified!"
        assert num_starts >= 0, "`num_starts` must be nonnegative integer!"
        assert num_samples >= num_starts, \
            "number of random samples (`num_samples`) must be " \
            "greater than number of starting points (`num_starts`)"
        if method != "L-BFGS-B":
            raise NotImplementedError(f"method={method!r}: only L-BFGS-B has a device path "
                                      "(bore_b200 has no CPU fallback)")
        options = dict(options or {})
        known = {"maxiter", "ftol", "gtol", "maxcor", "maxfun", "maxls"}
        unknown = set(options) - known
        if unknown:
            raise TypeError(f"unknown L-BFGS-B options {sorted(unknown)}")

        low, high, dim = _bounds_arrays(bounds)
        net = self._engine(dim)
        # host RNG exactly like the reference (bore/mixins.py:49): fp64 MT19937 uniforms from the caller's
        # generator -- the same numbers and the same state afterwards (bore_b200/hostrng.py), written straight
        # into the pinned staging buffer of the upload
        X64, X_init = net.uniform_to_device(random_state, low, high, num_samples, dim)
        X32 = X64.to(torch.float32)
        z_init = net.predict_dev(X32)  # raw model output, NO transform (bore/mixins.py:50-52)
        if num_starts == 0:
            i = int(net.topk_smallest(z_init, 1, negate=True)[0].item())
            return np.array(X_init, copy=True), None, i, -float(z_init[i].item())
        ind = net.topk_smallest(z_init, num_starts, negate=True)  # k smallest of f_init = -z
        X0 = X64.index_select(0, ind.to(torch.int64)) if num_starts < num_samples else X64
        res = net.lbfgsb_dev(X0, low, high, transform=self._min_transform_name(),
                             m=options.get("maxcor", 10), ftol=options.get("ftol", 2.2204460492503131e-09),
                             gtol=options.get("gtol", 1e-5), maxiter=options.get("maxiter", 15000),
                             maxfun=options.get("maxfun", 15000), maxls=options.get("maxls", 20))
        res["ind"] = ind
        # counters of the last argmax (evals = value+input-gradient evaluations performed)
        self._last_stats = dict(evals=res["evals"], rounds=res["rounds"], num_starts=num_starts)
        return X_init, res, None, None

    @staticmethod
    def _result(x, fun, nit, nfev, status, task):
        return OptimizeResult(x=x, fun=fun, nit=int(nit), nfev=int(nfev), njev=int(nfev),
                              status=int(status), success=bool(status == 0),
                              message=lbfgsb_message(status, task))

    def maxima(self, bounds, num_starts=5, num_samples=1024, method="L-BFGS-B",
               options=dict(maxiter=1000, ftol=1e-9), print_fn=print, random_state=None):
        """List of ``scipy.optimize.OptimizeResult``, one per start (bore/mixins.py:22-72)."""
        X_init, res, i0, f0 = self._maxima_device(bounds, num_starts, num_samples, method, options,
                                                  random_state)
        if res is None:
            return [OptimizeResult(x=X_init[i0], fun=f0, success=True)]
        host = {k: res[k].cpu().numpy() for k in ("x", "fun", "nit", "nfev", "status", "task")}
        results = []
        for i in range(num_starts):
            result = self._result(host["x"][i], np.float32(host["fun"][i]), host["nit"][i],
                                  host["nfev"][i], host["status"][i], host["task"][i])
            results.append(result)
            if print_fn is not None:
                print_fn(f"[Maximum {i+1:02d}: value={result.fun:.3f}] "
                         f"success: {result.success}, "
                         f"iterations: {result.nit:02d}, "
                         f"status: {result.status} ({result.message})")
        return results

    def argmax(self, bounds, filter_fn=_accept_all, *args, **kwargs):
        """The result with the smallest ``fun`` among those that ``(success or status == 1)
        and filter_fn(res)``; ``None`` if none qualifies (bore/mixins.py:74-89).

        ``num_start_points`` is accepted as an alias of ``num_starts`` (README.rst:96).  With the
        default ``filter_fn`` and ``print_fn=None`` nothing but the winner leaves the GPU.
        """
        if "num_start_points" in kwargs:
            kwargs["num_starts"] = kwargs.pop("num_start_points")
        names = ("num_starts", "num_samples", "method", "options", "print_fn", "random_state")
        kw = dict(num_starts=5, num_samples=1024, method="L-BFGS-B",
                  options=dict(maxiter=1000, ftol=1e-9), print_fn=print, random_state=None)
        kw.update(dict(zip(names, args)))
        kw.update(kwargs)
        print_fn = kw.pop("print_fn")
        from .data import UniqueFilter
        on_device = filter_fn is _accept_all or isinstance(filter_fn, UniqueFilter)
        if print_fn is not None or not on_device:
            # the reference's scan, verbatim (bore/mixins.py:80-87): every qualifying result is
            # shown to filter_fn once, in index order; first strict minimum wins (a NaN `fun`
            # never compares smaller, so it cannot win)
            res_best = None
            for res in self.maxima(bounds, print_fn=print_fn, **kw):
                if (res.success or res.status == 1) and filter_fn(res):
                    if res_best is None or res.fun < res_best.fun:
                        res_best = res
            return res_best
        X_init, res, i0, f0 = self._maxima_device(bounds, **kw)
        if res is None:
            return OptimizeResult(x=X_init[i0], fun=f0, success=True)
        net = self._engine(X_init.shape[1])
        keep = None
        if filter_fn is not _accept_all:
            # the plugin's duplicate filter for every result in one launch (bore/data.py:42-48)
            prev = filter_fn.stored()
            if prev.shape[0] > 0:
                keep = net.keep_unique_dev(res["x"].unsqueeze(0), net.to_device(prev[None], np.float64),
                                           rtol=filter_fn.rtol, atol=filter_fn.atol).reshape(-1)
                filter_fn.report_dropped(keep)
        key = int(net.select_best(res["fun"], res["status"], keep_dev=keep).item())
        if key == 0:
            return None
        return self._result_at(res, 0x7fffffff - (key & 0x7fffffff))

    def _result_at(self, res, i):
        import torch
        D = res["x"].shape[1]
        rec = torch.cat([res["x"][i], res["fun"][i:i + 1],
                         torch.stack([res[k][i] for k in ("nit", "nfev", "status", "task")]).to(torch.float64)])
        return self._result_from_record(rec.cpu().numpy(), D)

    @classmethod
    def _result_from_record(cls, rec, D):
        return cls._result(rec[:D].copy(), np.float32(rec[D]), rec[D + 1], rec[D + 2], rec[D + 3],
                           rec[D + 4])

    def argmax_sharded(self, bounds, num_starts, num_samples=None, method="L-BFGS-B",
                       options=dict(maxiter=1000, ftol=1e-9), random_state=None, group=None):
        """Multi-GPU argmax (build extension; the reference is single-process): every rank of the
        torch.distributed group holds the same weights and optimises ITS OWN ``num_starts`` start
        points (drawn from its own ``random_state``); one NCCL max all-reduce on the packed
        (value, index) key picks the global winner, whose record is then broadcast.  Returns the
        same ``OptimizeResult`` (or None) on every rank."""
        import torch
        import torch.distributed as dist
        from . import distributed as bd
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if num_samples is None:
            num_samples = num_starts
        X_init, res, i0, f0 = self._maxima_device(bounds, num_starts, num_samples, method, options,
                                                  random_state)
        assert res is not None, "argmax_sharded needs num_starts > 0"
        net = self._engine(X_init.shape[1])
        D = X_init.shape[1]
        key = net.select_best(res["fun"], res["status"], idx_offset=rank * num_starts)

        def record(i):
            return torch.cat([res["x"][i], res["fun"][i:i + 1],
                              torch.stack([res[k][i] for k in ("nit", "nfev", "status", "task")]).to(torch.float64)])
        gidx, rec = bd.global_winner(key, record, num_starts * world, D + 5, group=group)
        if gidx is None:
            return None
        out = self._result_from_record(rec.cpu().numpy(), D)
        out["global_index"] = gidx
        return out


class BatchMaximizableMixin(MaximizableMixin):
    """Adds ``argmax_batch``: a batch of maximisers by SVGD (bore/mixins.py:92-116)."""

    def __init__(self, transform=ops.identity, *args, **kwargs):
        super(BatchMaximizableMixin, self).__init__(transform=transform, *args, **kwargs)
        # maximization problem for SVGD (bore/mixins.py:97-98)
        self._func_max = convert(self, transform=transform)

    def argmax_batch(self, batch_size, bounds, length_scale=None, n_iter=1000, step_size=1e-3,
                     alpha=.9, eps=1e-6, tau=1.0, lambd=None, random_state=None):
        from .optimizers.svgd.base import SVGD, DistortionConstant, DistortionExpDecay
        from .optimizers.svgd.kernels import RadialBasis
        distortion = DistortionConstant() if lambd is None else DistortionExpDecay(lambd=lambd)
        kernel = RadialBasis(length_scale=length_scale)
        svgd = SVGD(kernel=kernel, n_iter=n_iter, step_size=step_size, alpha=alpha, eps=eps, tau=tau,
                    distortion=distortion)
        return svgd.optimize(self._func_max, batch_size, bounds=bounds, random_state=random_state)
