"""Function adapters (mirror of bore/decorators.py:6-79) -- the "value-and-gradient wrappers".

Same names, same shape algebra, same return conventions as the reference; what differs is the
engine: ``value_and_gradient`` does not run a TensorFlow tape, it traces the wrapped function
once (``bore_b200.ops``) and executes the fused CUDA forward + reverse-through-input kernel.
"""
from functools import wraps

import numpy as np

from . import ops


def stack(fn):
    """``fn(stacked)`` -> ``new_fn(*args)`` (bore/decorators.py:6-12)."""
    @wraps(fn)
    def new_fn(*args):
        return fn(ops.stack(args))
    return new_fn


def unstack(fn):
    """``fn(*components)`` -> ``new_fn(stacked)`` along the last axis (bore/decorators.py:15-21)."""
    @wraps(fn)
    def new_fn(args):
        return fn(*ops.unstack(args, axis=-1))
    return new_fn


def squeeze(axis):
    """Squeeze ``axis`` of the wrapped function's output (bore/decorators.py:24-34)."""
    def squeeze_dec(fn):
        @wraps(fn)
        def new_fn(*args, **kwargs):
            return ops.squeeze(fn(*args, **kwargs), axis=axis)
        return new_fn
    return squeeze_dec


def unbatch(fn):
    """Batched ``fn`` -> function of a single input (bore/decorators.py:37-45)."""
    @wraps(fn)
    def new_fn(input):
        return ops.squeeze(fn(ops.expand_dims(input, axis=0)), axis=0)
    return new_fn


def value_and_gradient(value_fn):
    """``x -> (value_fn(x), d sum(value_fn(x)) / dx)`` (bore/decorators.py:48-65).

    ``value_fn`` must reduce to ``transform(+-model(x))`` for a ``bore_b200`` model; it is traced
    per input shape and evaluated by the CUDA kernel.  Like the reference, the value comes back
    in the model's dtype (float32) and the gradient in the dtype of ``x``.
    """
    cache = {}

    @wraps(value_fn)
    def value_and_gradient_fn(x):
        x = np.asarray(x)
        expr = cache.get(x.shape)
        if expr is None:
            expr = value_fn(ops.Tracer(x.shape))
            if not isinstance(expr, ops.Expr):
                raise TypeError("value_and_gradient: the wrapped function must be built from a "
                                "bore_b200 model call (no generic autodiff on this path)")
            cache[x.shape] = expr
        D = expr.in_shape[-1]
        flat = x.reshape(-1, D)
        f, g = expr.model._native_value_and_grad(flat, expr.transform, expr.sign < 0)
        val = f.reshape(expr.shape)
        grad = g.reshape(x.shape).astype(x.dtype if x.dtype.kind == "f" else np.float64)
        return val, grad

    value_and_gradient_fn._bore_value_fn = value_fn
    return value_and_gradient_fn


def numpy_io(fn):
    """Array in, LIST of arrays out (bore/decorators.py:68-79)."""
    @wraps(fn)
    def new_fn(*args):
        outputs = fn(*[np.asarray(a) for a in args])
        return [np.asarray(o) for o in outputs]
    return new_fn
