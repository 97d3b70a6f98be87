"""Output transforms and the tiny expression tracer behind ``convert``.

The reference composes TensorFlow ops (``tf.identity / tf.sigmoid / tf.exp``, unary minus,
``expand_dims`` / ``squeeze``) around a Keras model and lets ``GradientTape`` differentiate the
result (bore/base.py:35-42, bore/decorators.py).  Here the same composition is *traced* once
into ``(model, sign, transform, output shape)`` and then executed by the fused CUDA
value-and-input-gradient kernel -- there is no autodiff framework on the path.

``identity``, ``sigmoid`` and ``exp`` are the TRANSFORMS of bore/plugins/hpbandster/base.py:18;
on numpy arrays they simply compute the function.
"""
import numpy as np


class Tracer:
    """Symbolic stand-in for the input tensor: tracks only its shape."""

    def __init__(self, shape):
        self.shape = tuple(shape)


class Expr:
    """T(sign * model(x)) with a tracked output shape."""

    def __init__(self, model, in_shape, shape, sign=1, transform="identity"):
        self.model, self.in_shape, self.shape = model, tuple(in_shape), tuple(shape)
        self.sign, self.transform = sign, transform

    def _with(self, **kw):
        d = dict(model=self.model, in_shape=self.in_shape, shape=self.shape, sign=self.sign,
                 transform=self.transform)
        d.update(kw)
        return Expr(**d)

    def __neg__(self):
        if self.transform != "identity":
            raise NotImplementedError("only transform(+-model(x)) has a device path")
        return self._with(sign=-self.sign)


def _apply(name, fn, v):
    if isinstance(v, Expr):
        if v.transform != "identity":
            raise NotImplementedError("only ONE output transform around the model has a device path")
        return v._with(transform=name)
    if isinstance(v, Tracer):
        raise TypeError("transforms apply to model outputs, not to the raw input")
    return fn(np.asarray(v))


def identity(v):
    return v if isinstance(v, (Expr, Tracer)) else np.asarray(v)


def sigmoid(v):
    return _apply("sigmoid", lambda a: 1.0 / (1.0 + np.exp(-a)), v)


def exp(v):
    return _apply("exp", np.exp, v)


TRANSFORMS = dict(identity=identity, sigmoid=sigmoid, exp=exp)


def expand_dims(v, axis):
    if isinstance(v, (Tracer, Expr)):
        shp = list(v.shape)
        ax = axis if axis >= 0 else axis + len(shp) + 1
        shp.insert(ax, 1)
        if isinstance(v, Tracer):
            return Tracer(shp)
        return v._with(shape=shp)
    return np.expand_dims(v, axis)


def squeeze(v, axis):
    if isinstance(v, (Tracer, Expr)):
        shp = list(v.shape)
        ax = axis if axis >= 0 else axis + len(shp)
        if shp[ax] != 1:
            raise ValueError(f"cannot squeeze axis {axis} of shape {tuple(shp)}")
        del shp[ax]
        if isinstance(v, Tracer):
            return Tracer(shp)
        return v._with(shape=shp)
    return np.squeeze(v, axis=axis)


def stack(values):
    return np.stack([np.asarray(v) for v in values])


def unstack(value, axis=-1):
    value = np.asarray(value)
    return [np.take(value, i, axis=axis) for i in range(value.shape[axis])]
