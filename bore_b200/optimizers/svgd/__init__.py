"""SVGD batch argmax (device-backed mirror of bore/optimizers/svgd)."""
from .base import SVGD, DistortionConstant, DistortionExpDecay, rank
from .kernels import RadialBasis

__all__ = ["SVGD", "RadialBasis", "DistortionConstant", "DistortionExpDecay", "rank"]
