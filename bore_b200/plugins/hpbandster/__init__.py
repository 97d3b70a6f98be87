"""HpBandSter plugin of the BORE-MLP hot path (mirror of bore/plugins/hpbandster/__init__.py).

``BOREHyperband`` (the LSTM multi-fidelity generator, bore/plugins/hpbandster/multi_fidelity.py)
is out of scope: a different model family, not BORE-MLP."""
from .base import BORE, ClassifierConfigGenerator, TRANSFORMS  # noqa: F401
from .types import (DenseConfigurationSpace, DenseConfiguration,  # noqa: F401
                    array_from_dict, dict_from_array)
