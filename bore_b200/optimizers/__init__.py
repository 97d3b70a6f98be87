"""Multi-start minimisation on the device (mirror of bore/optimizers) and the SVGD batch argmax."""
from .base import minimize_multi_start, multi_start
from .utils import from_bounds

__all__ = ["minimize_multi_start", "multi_start", "from_bounds"]
