from .base import SVGD

__all__ = ["SVGD"]
