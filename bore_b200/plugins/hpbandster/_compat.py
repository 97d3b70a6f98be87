"""Stand-ins for the two host frameworks the plugin plugs into, used ONLY when the real packages
are not importable (they are absent from this image and there is no network):

* ``ConfigSpace`` 0.4.18 (reference pin, setup.py:17) -- ``ConfigurationSpace``, the three
  hyperparameter types ``DenseConfigurationSpace`` supports
  (bore/plugins/hpbandster/types.py:69-83) and ``Configuration``.
* ``hpbandster`` 0.7.4 (setup.py:18) -- ``base_config_generator`` and a synchronous,
  single-process ``HyperBand`` (successive halving over geometric budgets) so the config
  generator can be exercised end to end.  HpBandSter's real Master/Dispatcher/Pyro4 RPC is the
  reference's control plane and is out of scope.

Everything here is host-side bookkeeping written from knowledge of those releases
[HB-semantics]; the vector encoding is pinned by the reference's own golden vector
(tests/test_types.py:90).  When the real packages are installed they are used instead.
"""
import logging

import numpy as np

try:  # pragma: no cover - not installed in this image
    import ConfigSpace as CS
    HAVE_CONFIGSPACE = True
except ImportError:
    CS = None
    HAVE_CONFIGSPACE = False

try:  # pragma: no cover - not installed in this image
    from hpbandster.optimizers.hyperband import HyperBand
    from hpbandster.core.base_config_generator import base_config_generator
    HAVE_HPBANDSTER = True
except ImportError:
    HyperBand = None
    base_config_generator = None
    HAVE_HPBANDSTER = False


# ----------------------------------------------------------------------------- ConfigSpace shim
if not HAVE_CONFIGSPACE:

    class _Hyperparameter:
        def __init__(self, name):
            self.name = name

    class UniformFloatHyperparameter(_Hyperparameter):
        def __init__(self, name, lower, upper, default_value=None, log=False):
            super().__init__(name)
            if log:
                raise NotImplementedError("log-scale hyperparameters are not in the shim")
            self.lower, self.upper = float(lower), float(upper)

        def to_vector(self, value):
            return (float(value) - self.lower) / (self.upper - self.lower)

        def from_vector(self, v):
            return float(min(max(v, 0.0), 1.0) * (self.upper - self.lower) + self.lower)

        def sample_vector(self, rs):
            return rs.uniform(0.0, 1.0)

    class UniformIntegerHyperparameter(_Hyperparameter):
        """ConfigSpace represents an integer range through a float range widened by 0.49999 on
        both sides, so every integer owns an equal share of [0, 1]."""

        def __init__(self, name, lower, upper, default_value=None, log=False):
            super().__init__(name)
            if log:
                raise NotImplementedError("log-scale hyperparameters are not in the shim")
            self.lower, self.upper = int(lower), int(upper)
            self._lo = self.lower - 0.49999
            self._hi = self.upper + 0.49999

        def to_vector(self, value):
            return (int(value) - self._lo) / (self._hi - self._lo)

        def from_vector(self, v):
            x = min(max(v, 0.0), 1.0) * (self._hi - self._lo) + self._lo
            return int(min(max(int(np.rint(x)), self.lower), self.upper))

        def sample_vector(self, rs):
            return self.to_vector(self.from_vector(rs.uniform(0.0, 1.0)))

    class CategoricalHyperparameter(_Hyperparameter):
        def __init__(self, name, choices, default_value=None):
            super().__init__(name)
            self.choices = tuple(choices)
            self.num_choices = len(self.choices)

        def to_vector(self, value):
            return float(self.choices.index(value))

        def from_vector(self, v):
            return self.choices[int(np.rint(v))]

        def sample_vector(self, rs):
            return float(rs.randint(0, self.num_choices))

    class ConfigurationSpace:
        """Hyperparameters kept sorted by name (what ConfigSpace 0.4 does for unconditioned
        spaces; the reference's test relies on it: index 0 is ``activation_fn_1``,
        tests/test_types.py:67-70)."""

        def __init__(self, name=None, seed=None):
            self.name = name
            self.random = np.random.RandomState(seed)
            self._hps = {}

        def seed(self, seed):
            self.random = np.random.RandomState(seed)

        def add_hyperparameter(self, hp):
            if hp.name in self._hps:
                raise ValueError(f"hyperparameter {hp.name!r} already in the space")
            self._hps[hp.name] = hp
            self._hps = dict(sorted(self._hps.items()))
            return hp

        def add_hyperparameters(self, hps):
            for hp in hps:
                self.add_hyperparameter(hp)
            return hps

        def get_hyperparameters(self):
            return list(self._hps.values())

        def get_hyperparameter_names(self):
            return list(self._hps.keys())

        def get_hyperparameter(self, name):
            return self._hps[name]

        def get_hyperparameter_by_idx(self, idx):
            return list(self._hps.keys())[idx]

        def get_idx_by_hyperparameter_name(self, name):
            return list(self._hps.keys()).index(name)

        def _configuration_class(self):
            return Configuration

        def sample_configuration(self, size=1):
            cls = self._configuration_class()
            out = []
            for _ in range(size):
                vec = np.array([hp.sample_vector(self.random) for hp in self.get_hyperparameters()])
                out.append(cls(self, vector=vec))
            return out if size > 1 else out[0]

    class Configuration:
        def __init__(self, configuration_space, values=None, vector=None):
            if (values is None) == (vector is None):
                raise ValueError("exactly one of `values` and `vector` must be given")
            self.configuration_space = configuration_space
            hps = configuration_space.get_hyperparameters()
            if values is not None:
                unknown = set(values) - {hp.name for hp in hps}
                if unknown:
                    raise ValueError(f"unknown hyperparameters {sorted(unknown)}")
                self._vector = np.array([hp.to_vector(values[hp.name]) for hp in hps], np.float64)
            else:
                self._vector = np.asarray(vector, np.float64).copy()
                assert self._vector.shape == (len(hps),)

        def get_array(self):
            return self._vector

        def get_dictionary(self):
            hps = self.configuration_space.get_hyperparameters()
            return {hp.name: hp.from_vector(v) for hp, v in zip(hps, self._vector)}

        def __getitem__(self, name):
            return self.get_dictionary()[name]

        def __eq__(self, other):
            return isinstance(other, Configuration) and \
                self.get_dictionary() == other.get_dictionary()

        def __repr__(self):
            return f"Configuration({self.get_dictionary()})"

    class _CSNamespace:
        pass

    CS = _CSNamespace()
    CS.ConfigurationSpace = ConfigurationSpace
    CS.Configuration = Configuration
    CS.UniformFloatHyperparameter = UniformFloatHyperparameter
    CS.UniformIntegerHyperparameter = UniformIntegerHyperparameter
    CS.CategoricalHyperparameter = CategoricalHyperparameter


# ----------------------------------------------------------------------------- hpbandster shim
if not HAVE_HPBANDSTER:

    class base_config_generator:
        """hpbandster.core.base_config_generator duck type: a logger, ``get_config(budget)`` to
        be overridden, ``new_result(job)`` that reports crashed jobs."""

        def __init__(self, logger=None):
            self.logger = logging.getLogger("hpbandster") if logger is None else logger

        def get_config(self, budget):
            raise NotImplementedError("This function needs to be overwritten in %s." %
                                      self.__class__.__name__)

        def new_result(self, job, update_model=True):
            if job.exception is not None:
                self.logger.warning("job {} failed with exception\n{}".format(job.id, job.exception))

    class Job:
        """The slice of hpbandster.core.dispatcher.Job the config generator reads
        (plugins/hpbandster/base.py:276-288)."""

        def __init__(self, id, config, budget):
            self.id = id
            self.kwargs = dict(config=config, budget=budget)
            self.result = None
            self.exception = None

    class _Master:
        """Grandparent initializer the reference calls directly
        (``super(HyperBand, self).__init__(config_generator=cg, **kwargs)``,
        plugins/hpbandster/base.py:56): keeps the generator, a logger and a config dict."""

        def __init__(self, config_generator, run_id="bore_b200", logger=None, **kwargs):
            self.config_generator = config_generator
            self.run_id = run_id
            self.logger = logging.getLogger("hpbandster") if logger is None else logger
            self.config_generator.logger = self.logger
            self.config = {}
            self.extra = kwargs

    class HyperBand(_Master):
        """Synchronous single-process Hyperband: for each bracket, sample ``n0`` configs through
        ``config_generator.get_config(budget)``, evaluate them with ``compute_fn(config, budget)
        -> loss`` and promote the best 1/eta to the next budget.  Mirrors hpbandster's bracket
        arithmetic; there is no worker pool, dispatcher or RPC."""

        def __init__(self, configspace=None, eta=3, min_budget=0.01, max_budget=1, **kwargs):
            raise NotImplementedError("use bore_b200.plugins.hpbandster.BORE")

        def run(self, n_iterations, compute_fn):
            results = []
            job_id = 0
            for it in range(n_iterations):
                s = self.max_SH_iter - 1 - (it % self.max_SH_iter)
                n0 = int(np.floor(self.max_SH_iter / (s + 1)) * self.eta ** s)
                ns = [max(int(n0 * (self.eta ** (-i))), 1) for i in range(s + 1)]
                budgets = self.budgets[(-s - 1):]
                survivors = None
                for n_i, budget in zip(ns, budgets):
                    if survivors is None:
                        configs = [self.config_generator.get_config(budget)[0] for _ in range(n_i)]
                    else:
                        configs = survivors[:n_i]
                    losses = []
                    for cfg in configs:
                        job = Job((it, 0, job_id), cfg, budget)
                        job_id += 1
                        job.result = {"loss": float(compute_fn(cfg, budget)), "info": {}}
                        self.config_generator.new_result(job)
                        losses.append(job.result["loss"])
                        results.append((cfg, budget, job.result["loss"]))
                    order = np.argsort(losses, kind="stable")
                    survivors = [configs[i] for i in order]
            return results
