from .base import minimize_multi_start

__all__ = ["minimize_multi_start"]
