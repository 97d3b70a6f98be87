"""Observation store feeding ``fit`` (mirror of bore/data.py:4-48), plus the device form of its
two computations: ``quantile_labels`` (bore/data.py:31-35 for many
problems at once) and ``UniqueFilter`` (bore/data.py:42-48 as the ``filter_fn`` of ``argmax``,
evaluated for every result in one launch instead of one Python callback per result).

``MultiFidelityRecord`` (bore/data.py:51-261) is the store behind the LSTM multi-fidelity plugin
(SURVEY.md section 8f row 4): observations keyed by configuration and budget, per-rung quantile
thresholds, and the padded (inputs, targets) sequences ``bore_lstm_fit`` trains on."""
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


class MultiFidelityRecord:
    """Observations ``(x, y, budget)`` grouped by configuration (bore/data.py:51-261).

    ``_data[key(x)][budget] = y`` and ``_targets[budget] = [y, ...]`` as in the reference (both
    views are kept: the plugin and the reference's tests read them).  Rung t is the t-th smallest
    budget seen so far.  Labels use ``<=`` against the per-rung ``gamma`` quantile (bore/data.py:171-173,
    209 -- unlike ``Record``, whose inequality is strict)."""

    def __init__(self, gamma=None):
        self._data = {}
        self._targets = {}
        self.gamma = gamma

    @staticmethod
    def compute_key(x):
        return tuple(x.tolist())

    def append(self, x, y, b):
        # a repeated (x, b) overwrites the per-configuration value but is still appended to the
        # rung's target list, exactly as the reference does (bore/data.py:61-72)
        self._data.setdefault(self.compute_key(x), {})[b] = y
        self._targets.setdefault(b, []).append(y)

    # ---- rungs and budgets
    def budgets(self, reverse=False):
        return sorted(self._targets, reverse=reverse)

    def budget(self, t):
        return self.budgets()[t]

    def num_rungs(self):
        return len(self._targets)

    def _rung_size_from_budget(self, b):
        return len(self._targets[b])

    def rung_size(self, t):
        return self._rung_size_from_budget(self.budget(t))

    def rung_sizes(self):
        return [len(self._targets[b]) for b in self.budgets()]

    def size(self):
        return sum(self.rung_sizes())

    def highest_rung(self, min_size=1):
        """Index of the highest rung holding at least ``min_size`` evaluations, else None."""
        ok = [t for t, n in enumerate(self.rung_sizes()) if n >= min_size]
        return ok[-1] if ok else None

    # ---- features and targets
    def num_features(self):
        return len(self._data)

    def load_feature_matrix(self):
        return np.vstack(list(self._data))

    def _targets_from_budget(self, b):
        return self._targets[b]

    def targets(self, t):
        return self._targets[self.budget(t)]

    def _threshold_from_budget(self, b):
        return np.quantile(self._targets[b], q=self.gamma)

    def threshold(self, t):
        return self._threshold_from_budget(self.budget(t))

    def thresholds(self):
        return [self._threshold_from_budget(b) for b in self.budgets()]

    def _binary_labels_from_budget(self, b):
        return np.less_equal(self._targets[b], self._threshold_from_budget(b))

    def binary_labels(self, t):
        return self._binary_labels_from_budget(self.budget(t))

    # ---- sequences for the recurrent classifier
    def sequences_dict(self, pad_value=-1., binary=True, return_indices=False):
        """``{key: [label or value per rung]}`` with ``pad_value`` where the configuration was not
        evaluated at that rung (bore/data.py:183-222); with ``return_indices`` also the
        per-rung presence flags."""
        assert not binary or self.gamma is not None, \
            "Must instantiate with `gamma` specified for binary labels!"
        budgets = self.budgets()
        taus = [self._threshold_from_budget(b) for b in budgets]
        sequences, indices = {}, {}
        for key, by_budget in self._data.items():
            ys, present = [], []
            for b, tau in zip(budgets, taus):
                seen = b in by_budget
                present.append(seen)
                if not seen:
                    ys.append(pad_value)
                elif binary:
                    ys.append(int(by_budget[b] <= tau))
                else:
                    ys.append(by_budget[b])
            sequences[key], indices[key] = ys, present
        return (sequences, indices) if return_indices else sequences

    def sequences(self, pad_value=-1., binary=True):
        """Padded arrays ``inputs (N, T, D)`` float64 -- the configuration repeated at the rungs it
        was evaluated at, ``pad_value`` rows elsewhere -- and ``targets (N, T, 1)``
        (bore/data.py:224-251)."""
        seqs, present = self.sequences_dict(pad_value=pad_value, binary=binary, return_indices=True)
        inputs, targets = [], []
        for key, ys in seqs.items():
            rows = np.full((len(ys), len(key)), pad_value, dtype="float64")
            rows[present[key]] = np.array(key)
            inputs.append(rows)
            targets.append(np.expand_dims(ys, axis=-1))
        return np.stack(inputs, axis=0), np.stack(targets, axis=0)

    def is_duplicate(self, x, rtol=1e-5, atol=1e-8):
        """True if x is ``np.allclose`` to any stored configuration (bore/data.py:253-261)."""
        return any(np.allclose(np.array(k), x, rtol=rtol, atol=atol) for k in self._data)


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
        n = self.record.num_features() if hasattr(self.record, "num_features") else len(self.record.features)
        if n == 0:
            return np.zeros((0, 0), np.float64)
        return np.ascontiguousarray(self.record.load_feature_matrix(), np.float64)
