"""HpBandSter plugins of the BORE hot path (mirror of bore/plugins/hpbandster/__init__.py):
``BORE`` (MLP classifier) and ``BOREHyperband`` (LSTM multi-fidelity classifier, SURVEY.md section 8f
row 4)."""
from .base import BORE, ClassifierConfigGenerator, TRANSFORMS  # noqa: F401
from .multi_fidelity import BOREHyperband, SequenceClassifierConfigGenerator  # noqa: F401
from .types import (DenseConfigurationSpace, DenseConfiguration,  # noqa: F401
                    array_from_dict, dict_from_array)
