"""RBF kernel of the SVGD batch argmax (mirror of bore/optimizers/svgd/kernels.py:13-28);
``value_and_grad`` runs the ``bore_svgd_kernel_value_and_grad`` CUDA kernel."""
import numpy as np

from ... import engine


class RadialBasis:
    """``length_scale=None`` selects the median heuristic (kernels.py:4-10)."""

    def __init__(self, length_scale=1.0):
        self.length_scale = length_scale

    def value_and_grad(self, X):
        """X (n, D) -> K (n, n), K_grad (n, D) with K_grad[i] = d/dx_i sum_j k(x_j, x_i)... in the
        reference's sign convention: ``2 * sum_j gamma * (x_i - x_j) * K[i, j]`` (kernels.py:26)."""
        return engine.svgd_kernel_value_and_grad(np.asarray(X, np.float64), self.length_scale)
