"""Observation store feeding ``fit`` (mirror of bore/data.py:4-48), plus the device form of its
two computations: ``quantile_labels`` (bore/data.py:31-35 for many
problems at once) and ``UniqueFilter`` (bore/data.py:42-48 as the ``filter_fn`` of ``argmax``,
evaluated for every result in one launch instead of one Python callback per result).

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


def quantile_labels(y, gamma, net):
    """``z = y < np.quantile(y, gamma)`` per row of ``y`` (M, N), computed by the
    ``bore_quantile_labels`` kernel of ``net`` (a ``NativeMLP``); returns (z bool (M, N), tau (M,)).
    Bit-identical to bore/data.py:33-34 applied row by row."""
    y = np.ascontiguousarray(np.atleast_2d(np.asarray(y, np.float64)))
    z, tau = net.quantile_labels_dev(net.to_device(y, np.float64), gamma, want_tau=True)
    return z.cpu().numpy() != 0, tau.cpu().numpy()


class UniqueFilter:
    """``filter_fn`` for ``argmax`` that rejects results already in ``record``
    (bore/plugins/hpbandster/base.py:227-231).  Called as a function it is the reference's
    host predicate; ``MaximizableMixin.argmax`` recognises the type and evaluates it for all
    results at once on the device (``bore_is_duplicate``) -- same selection, no per-result
    callback."""

    def __init__(self, record, rtol=1e-5, atol=1e-8, logger=None):
        self.record, self.rtol, self.atol, self.logger = record, rtol, atol, logger

    def __call__(self, res):
        dup = self.record.is_duplicate(res.x, rtol=self.rtol, atol=self.atol)
        if dup and self.logger is not None:
            self.logger.warning("Duplicate detected! Skipping...")
        return not dup

    def report_dropped(self, keep_dev):
        """Device path: one warning per duplicate that the keep mask dropped -- the lines the
        reference logs from its per-result callback (plugins/hpbandster/base.py:227-231).  The
        mask is only read back when somebody listens."""
        if self.logger is None:
            return
        enabled = getattr(self.logger, "isEnabledFor", None)
        if enabled is not None and not enabled(30):  # logging.WARNING
            return
        for _ in range(int((keep_dev == 0).sum().item())):
            self.logger.warning("Duplicate detected! Skipping...")

    def stored(self):
        """(N, D) float64 matrix of the stored feature vectors (N may be 0)."""
        if not self.record.features:
            return np.zeros((0, 0), np.float64)
        return np.ascontiguousarray(self.record.load_feature_matrix(), np.float64)
