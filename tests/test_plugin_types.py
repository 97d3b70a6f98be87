"""Dense one-hot encoding of the HpBandSter plugin (CPU).  The known answers are the reference's
own (tests/test_types.py:35-106 of ltiao/bore), restated against the ConfigSpace stand-in when
ConfigSpace is not installed."""
import numpy as np
import pytest

from bore_b200.plugins.hpbandster._compat import CS
from bore_b200.plugins.hpbandster.types import (DenseConfigurationSpace, DenseConfiguration,
                                                array_from_dict, dict_from_array)

SEED = 8888


@pytest.fixture
def config_space():
    cs = CS.ConfigurationSpace(seed=SEED)
    cs.add_hyperparameter(CS.UniformIntegerHyperparameter("n_units_1", lower=0, upper=5))
    cs.add_hyperparameter(CS.UniformIntegerHyperparameter("n_units_2", lower=0, upper=5))
    cs.add_hyperparameter(CS.UniformFloatHyperparameter("dropout_1", lower=0, upper=0.9))
    cs.add_hyperparameter(CS.UniformFloatHyperparameter("dropout_2", lower=0, upper=0.9))
    cs.add_hyperparameter(CS.CategoricalHyperparameter("activation_fn_1", ["tanh", "relu"]))
    cs.add_hyperparameter(CS.CategoricalHyperparameter("activation_fn_2", ["tanh", "relu"]))
    cs.add_hyperparameter(CS.UniformIntegerHyperparameter("init_lr", lower=0, upper=5))
    cs.add_hyperparameter(CS.CategoricalHyperparameter("lr_schedule", ["cosine", "const"]))
    cs.add_hyperparameter(CS.UniformIntegerHyperparameter("batch_size", lower=0, upper=3))
    return cs


def test_shapes(config_space):
    cs_dense = DenseConfigurationSpace(config_space, seed=SEED)
    assert cs_dense.get_dimensions(sparse=True) == 9
    assert cs_dense.get_dimensions(sparse=False) == 12
    bounds = cs_dense.get_bounds()
    np.testing.assert_array_equal(bounds.lb, np.zeros(12))
    np.testing.assert_array_equal(bounds.ub, np.ones(12))
    assert isinstance(cs_dense.sample_configuration(), CS.Configuration)
    assert isinstance(cs_dense.sample_configuration(size=1), CS.Configuration)
    configs = cs_dense.sample_configuration(size=5)
    assert len(configs) == 5
    for config in configs:
        assert isinstance(config, CS.Configuration)
        a = config.to_array()
        assert a.shape == (12,) and np.all(a >= 0) and np.all(a <= 1)


def test_dense_encoding_golden_vector(config_space):
    cs_dense = DenseConfigurationSpace(config_space, seed=SEED)
    assert cs_dense.get_hyperparameter_by_idx(0) == "activation_fn_1"
    dct = {'activation_fn_1': 'relu', 'activation_fn_2': 'tanh', 'batch_size': 2,
           'dropout_1': 0.39803953082292726, 'dropout_2': 0.022039062686389176, 'init_lr': 0,
           'lr_schedule': 'cosine', 'n_units_1': 5, 'n_units_2': 1}
    array = DenseConfiguration(cs_dense, values=dct).to_array()
    assert np.less_equal(0., array).all() and np.less_equal(array, 1.).all()
    np.testing.assert_array_almost_equal(
        array, [0., 1., 1., 0., 0.62500063, 0.44226615, 0.02448785, 0.08333194, 1., 0.,
                0.91666806, 0.24999917])
    np.testing.assert_array_equal(array, array_from_dict(cs_dense, dct))
    # exact round trip
    assert DenseConfiguration.from_array(cs_dense, array).get_dictionary() == dct
    assert dict_from_array(cs_dense, array) == dct
    # soft one-hot: arg-max decides
    array[0], array[1] = 0.8, 0.6
    assert dict_from_array(cs_dense, array)["activation_fn_1"] == "tanh"


def test_unsupported_hyperparameter_type(config_space):
    class Odd:
        name = "odd"
    config_space._hps["odd"] = Odd() if hasattr(config_space, "_hps") else None
    if not hasattr(config_space, "_hps"):
        pytest.skip("real ConfigSpace validates types itself")
    with pytest.raises(NotImplementedError):
        DenseConfigurationSpace(config_space, seed=SEED)


def test_hyperband_budget_arithmetic():
    """BORE.__init__ copies Hyperband's bracket boilerplate (plugins/hpbandster/base.py:63-81)."""
    from bore_b200.plugins.hpbandster import BORE
    cs = CS.ConfigurationSpace(seed=0)
    cs.add_hyperparameter(CS.UniformFloatHyperparameter("x", lower=0, upper=1))
    opt = BORE(cs, eta=3, min_budget=1 / 9, max_budget=1, seed=0)
    assert opt.max_SH_iter == 3
    np.testing.assert_allclose(opt.budgets, [1 / 9, 1 / 3, 1.0])
    assert opt.config["gamma"] == pytest.approx(1 / 3)
    cg = opt.config_generator
    assert cg.input_dim == 1 and cg.num_starts == 5 and cg.num_samples == 1024
    # warm-up: random proposals until num_random_init results are in (no GPU touched)
    cg.random_rate = None
    cfg, info = cg.get_config(1.0)
    assert set(cfg) == {"x"} and info == {}
