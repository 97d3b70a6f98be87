"""``convert`` and the suggestion-distortion helpers (mirror of bore/base.py:7-64)."""
from scipy.stats import truncnorm

from . import ops
from .decorators import unbatch, value_and_gradient, numpy_io, squeeze


def convert(model, transform=ops.identity):
    """Model -> ``fn(x: (D,)) -> [f: (), g: (D,)]`` ready for ``scipy.optimize`` with
    ``jac=True`` (bore/base.py:7-42).  Also accepts ``(S, D)`` batches, in which case ``g`` is
    the gradient of ``sum(f)`` -- the property bore/optimizers/base.py:53 relies on.  Every
    evaluation runs the fused CUDA value-and-input-gradient kernel."""
    @numpy_io
    @value_and_gradient
    @squeeze(axis=-1)
    @unbatch
    def fn(x):
        return transform(model(x))

    fn._bore_model = model
    fn._bore_transform = transform
    return fn


def truncated_normal(loc, scale, lower, upper):
    """Frozen ``truncnorm`` on [lower, upper] around loc (bore/base.py:45-48)."""
    return truncnorm(a=(lower - loc) / scale, b=(upper - loc) / scale, loc=loc, scale=scale)


def maybe_distort(loc, distortion=None, bounds=None, random_state=None, print_fn=print):
    """Optionally resample the suggestion from a truncated normal (bore/base.py:51-64)."""
    if distortion is None:
        return loc
    assert bounds is not None, "must specify bounds!"
    ret = truncated_normal(loc=loc, scale=distortion, lower=bounds.lb,
                           upper=bounds.ub).rvs(random_state=random_state)
    print_fn(f"Suggesting x={ret} (after applying distortion={distortion:.3E})")
    return ret
