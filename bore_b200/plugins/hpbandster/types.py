"""Dense one-hot encoding of hyperparameter configurations into the unit cube -- the boundary
helper of the HpBandSter plugin (same public names as bore/plugins/hpbandster/types.py:7-136:
``DenseConfigurationSpace``, ``DenseConfiguration``, ``array_from_dict``, ``dict_from_array``).

The classifier and its argmax live on ``[0, 1]^D_dense``.  A configuration's ConfigSpace vector
("sparse": one slot per hyperparameter) is laid out into "dense" coordinates by a list of
``_Slot`` records: a numeric hyperparameter occupies one dense coordinate and keeps its vector
value, a categorical with k choices occupies k coordinates holding its one-hot code.  Decoding
takes the arg-max of each categorical block, so a soft one-hot coming out of the optimiser is
legal (tests/test_types.py:98-106 of the reference).

``ConfigSpace`` is used when installed, else the stand-in of ``_compat`` (this image has
neither ConfigSpace nor hpbandster).  Host-side only; nothing here touches the GPU.
"""
from collections import namedtuple

import numpy as np
from scipy.optimize import Bounds

from ._compat import CS

# sparse index, first dense index, number of dense coordinates, is-categorical
_Slot = namedtuple("_Slot", "src trg width categorical")


def _layout(hyperparameters):
    slots, trg = [], 0
    for src, hp in enumerate(hyperparameters):
        if isinstance(hp, CS.CategoricalHyperparameter):
            slots.append(_Slot(src, trg, int(hp.num_choices), True))
        elif isinstance(hp, (CS.UniformIntegerHyperparameter, CS.UniformFloatHyperparameter)):
            slots.append(_Slot(src, trg, 1, False))
        else:
            raise NotImplementedError(
                "Only hyperparameters of types `CategoricalHyperparameter`, "
                "`UniformIntegerHyperparameter`, `UniformFloatHyperparameter` are supported!")
        trg += slots[-1].width
    return slots, trg


class DenseConfigurationSpace(CS.ConfigurationSpace):
    """A ConfigurationSpace holding ``other``'s hyperparameters only (conditions, forbidden
    clauses and the seed of ``other`` are not carried over) that also knows its dense layout."""

    def __init__(self, other, *args, **kwargs):
        super(DenseConfigurationSpace, self).__init__(*args, **kwargs)
        self.add_hyperparameters(other.get_hyperparameters())
        self._slots, self.size_dense = _layout(self.get_hyperparameters())
        self.size_sparse = len(self._slots)
        # the reference's attribute names, kept for callers that introspect the mapping
        self.nums = [(s.src, s.trg) for s in self._slots if not s.categorical]
        self.cats = [(s.src, s.trg, s.width) for s in self._slots if s.categorical]

    def _configuration_class(self):  # hook of the ConfigSpace stand-in
        return DenseConfiguration

    def get_dimensions(self, sparse=False):
        return self.size_sparse if sparse else self.size_dense

    def get_bounds(self):
        """The unit cube as ``scipy.optimize.Bounds`` (what ``argmax`` receives)."""
        return Bounds(np.zeros(self.size_dense), np.ones(self.size_dense))

    def sample_configuration(self, size=1):
        drawn = super(DenseConfigurationSpace, self).sample_configuration(size=size)
        many = size > 1
        out = [c if isinstance(c, DenseConfiguration)
               else DenseConfiguration(self, values=c.get_dictionary())
               for c in (drawn if many else [drawn])]
        return out if many else out[0]

    # -- the two directions of the encoding ------------------------------------------------
    def densify(self, sparse, dtype="float64"):
        dense = np.zeros(self.size_dense, dtype=dtype)
        for s in self._slots:
            if s.categorical:
                dense[s.trg + int(sparse[s.src])] = 1
            else:
                dense[s.trg] = sparse[s.src]
        return dense

    def sparsify(self, dense, dtype="float64"):
        dense = np.asarray(dense)
        sparse = np.empty(self.size_sparse, dtype=dtype)
        for s in self._slots:
            block = dense[s.trg:s.trg + s.width]
            sparse[s.src] = np.argmax(block) if s.categorical else block[0]
        return sparse


class DenseConfiguration(CS.Configuration):

    def __init__(self, configuration_space, *args, **kwargs):
        assert isinstance(configuration_space, DenseConfigurationSpace)
        super(DenseConfiguration, self).__init__(configuration_space, *args, **kwargs)

    @classmethod
    def from_array(cls, configuration_space, array_dense, dtype="float64"):
        assert isinstance(configuration_space, DenseConfigurationSpace)
        return cls(configuration_space=configuration_space,
                   vector=configuration_space.sparsify(array_dense, dtype))

    def to_array(self, dtype="float64"):
        return self.configuration_space.densify(self.get_array(), dtype)


def array_from_dict(config_space, dct):
    return DenseConfiguration(config_space, values=dct).to_array()


def dict_from_array(config_space, array):
    return DenseConfiguration.from_array(config_space, array_dense=array).get_dictionary()
