"""``convert`` and the suggestion-distortion helpers (mirror of bore/base.py:7-64)."""
import numpy as np
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


def maybe_distort_batch(loc, distortion=None, bounds=None, random_state=None, net=None):
    """``maybe_distort`` for a batch of suggestions ``loc`` (P, D) on the device:
    the uniform variates scipy's ``truncnorm.rvs`` would consume -- one per coordinate, suggestion
    after suggestion, i.e. what P sequential ``maybe_distort`` calls sharing ``random_state`` draw --
    come from the caller's MT19937 stream on the host; the truncated-normal ppf runs on the GPU
    (``bore_truncnorm_distort``).  Same values as the reference's host path to 1e-9."""
    from sklearn.utils import check_random_state
    from .optimizers.utils import from_bounds
    loc = np.atleast_2d(np.asarray(loc, np.float64))
    if distortion is None:
        return loc
    assert bounds is not None, "must specify bounds!"
    assert net is not None, "maybe_distort_batch runs on a NativeMLP's device"
    (low, high), dim = from_bounds(bounds)
    assert dim == loc.shape[1]
    rs = check_random_state(random_state)
    u = rs.uniform(size=loc.shape)
    out = net.distort_dev(net.to_device(loc, np.float64), distortion, low, high, net.to_device(u, np.float64))
    return out.cpu().numpy()
