"""Observation store feeding ``fit`` (mirror of bore/data.py:4-48).  Host-side numpy only.

``MultiFidelityRecord`` (bore/data.py:51-261) feeds the LSTM plugin only and is out of scope
(SURVEY.md section 8f)."""
import numpy as np


class Record:
    """Append-only list of (x, y[, budget]) observations."""

    def __init__(self):
        self.features, self.targets, self.budgets = [], [], []

    def size(self):
        return len(self.targets)

    def append(self, x, y, b=None):
        self.features.append(x)
        self.targets.append(y)
        if b is not None:
            self.budgets.append(b)

    def load_feature_matrix(self):
        return np.vstack(self.features)

    def load_target_vector(self):
        return np.hstack(self.targets)

    def load_regression_data(self):
        return self.load_feature_matrix(), self.load_target_vector()

    def load_classification_data(self, gamma):
        """Quantile labelling (bore/data.py:31-35): ``z = y < quantile(y, gamma)`` -- linear
        interpolation quantile, STRICT inequality."""
        X, y = self.load_regression_data()
        return X, np.less(y, np.quantile(y, q=gamma))

    def is_duplicate(self, x, rtol=1e-5, atol=1e-8):
        """True if x is ``np.allclose`` to any stored feature vector (bore/data.py:42-48)."""
        for x_prev in self.features:
            if np.allclose(x_prev, x, rtol=rtol, atol=atol):
                return True
        return False
