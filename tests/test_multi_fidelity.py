"""``MultiFidelityRecord`` (mirror of bore/data.py:51-261) against the outputs of the reference's own
class on the same seeded streams (tests/golden/multi_fidelity_golden.npz, written by
tests/golden/make_golden.py --multi-fidelity from /root/reference), plus the host logic of the
multi-fidelity plugin that needs no GPU.  CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import multi_fidelity_cases  # noqa: E402  (the seeded streams; no reference import)

from bore_b200.data import MultiFidelityRecord, UniqueFilter  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "multi_fidelity_golden.npz"))
CASES = multi_fidelity_cases()


def _record(c):
    rec = MultiFidelityRecord(gamma=c["gamma"])
    for x, y, b in c["stream"]:
        rec.append(x=x, y=y, b=b)
    return rec


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_record_matches_reference(ci):
    rec, k = _record(CASES[ci]), f"c{ci}/"
    assert np.array_equal(rec.budgets(), GOLD[k + "budgets"])
    assert rec.budgets(reverse=True) == rec.budgets()[::-1]
    assert np.array_equal(rec.rung_sizes(), GOLD[k + "rung_sizes"])
    assert rec.size() == int(GOLD[k + "size"]) and rec.num_features() == int(GOLD[k + "num_features"])
    assert rec.num_rungs() == len(GOLD[k + "budgets"])
    assert np.array_equal(rec.thresholds(), GOLD[k + "thresholds"])          # bit-exact: np.quantile
    hr = [-1 if rec.highest_rung(m) is None else rec.highest_rung(m) for m in (1, 3, 5, 8, 100)]
    assert np.array_equal(hr, GOLD[k + "highest_rung"])
    for t in range(rec.num_rungs()):
        assert np.array_equal(rec.binary_labels(t), GOLD[k + f"labels{t}"])
        assert rec.rung_size(t) == GOLD[k + "rung_sizes"][t]
        assert rec.threshold(t) == GOLD[k + "thresholds"][t]
        assert rec.budget(t) == GOLD[k + "budgets"][t]
        assert rec.targets(t) is rec._targets[rec.budget(t)]
    for name, pad in (("m1", -1.0), ("tiny", 1e-9)):
        X, Y = rec.sequences(pad_value=pad, binary=True)
        assert X.dtype == GOLD[k + f"X_{name}"].dtype and np.array_equal(X, GOLD[k + f"X_{name}"])
        assert Y.dtype == GOLD[k + f"Y_{name}"].dtype and np.array_equal(Y, GOLD[k + f"Y_{name}"])
    _, Yv = rec.sequences(pad_value=-1.0, binary=False)
    assert np.array_equal(Yv, GOLD[k + "Y_values"])
    dup = np.array([rec.is_duplicate(x) for x in GOLD[k + "cand"]])
    assert np.array_equal(dup, GOLD[k + "dup"]) and dup.any() and not dup.all()
    assert rec.load_feature_matrix().shape == (rec.num_features(), CASES[ci]["D"])


def test_record_edge_cases():
    rec = MultiFidelityRecord(gamma=0.25)
    assert rec.num_rungs() == 0 and rec.size() == 0 and rec.highest_rung() is None
    assert rec.budgets() == [] and rec.rung_sizes() == [] and not rec.is_duplicate(np.zeros(2))
    rec.append(np.array([0.1, 0.2]), 1.0, 3.0)   # a configuration first seen at a HIGH budget
    rec.append(np.array([0.3, 0.4]), 2.0, 1.0)
    X, Y = rec.sequences(pad_value=-1.0)
    assert X.shape == (2, 2, 2) and Y.shape == (2, 2, 1)
    assert np.array_equal(X[0], [[-1, -1], [0.1, 0.2]]) and np.array_equal(X[1], [[0.3, 0.4], [-1, -1]])
    assert Y[0, 0, 0] == -1 and Y[1, 1, 0] == -1 and Y[0, 1, 0] == 1 and Y[1, 0, 0] == 1  # y <= tau, one value
    d, ind = rec.sequences_dict(return_indices=True)
    assert list(ind.values()) == [[False, True], [True, False]]
    with pytest.raises(AssertionError):
        MultiFidelityRecord().sequences_dict(binary=True)
    assert MultiFidelityRecord.compute_key(np.array([1.0, 2.0])) == (1.0, 2.0)
    f = UniqueFilter(rec)
    assert f.stored().shape == (2, 2)
    assert UniqueFilter(MultiFidelityRecord(0.5)).stored().shape == (0, 0)


def test_generator_host_logic_without_gpu():
    """Constructor checks, the warm-up gate and the record plumbing run before any CUDA call."""
    import logging
    from bore_b200.plugins.hpbandster import BOREHyperband, SequenceClassifierConfigGenerator
    from bore_b200.plugins.hpbandster._compat import CS, Job
    cs = CS.ConfigurationSpace(seed=0)
    cs.add_hyperparameter(CS.UniformFloatHyperparameter("a", lower=0.0, upper=2.0))
    cs.add_hyperparameter(CS.CategoricalHyperparameter("c", ["u", "v", "w"]))
    opt = BOREHyperband(cs, eta=3, min_budget=1 / 9, max_budget=1, seed=0, num_random_init=4,
                        random_rate=None, logger=logging.getLogger("bore-mf-test"))
    cg = opt.config_generator
    assert isinstance(cg, SequenceClassifierConfigGenerator)
    assert cg.input_dim == 4 and cg.mask_value == -1.0 and cg.gamma == 1 / 3
    assert opt.max_SH_iter == 3 and np.allclose(opt.budgets, [1 / 9, 1 / 3, 1.0])
    assert len(cg.model_factory.cells) == 2 and cg.model_factory.cells[0].activation == "elu"
    cfg, info = cg.get_config(1.0)          # no rung has 4 observations: random candidate, no CUDA
    assert set(cfg) == {"a", "c"} and info == {}
    for j in range(3):
        job = Job((0, 0, j), cg.config_space.sample_configuration().get_dictionary(), 1 / 9)
        job.result = {"loss": float(j)}
        cg.new_result(job)
    assert cg.record.rung_sizes() == [3] and cg.record.highest_rung(4) is None
    cfg, _ = cg.get_config(1 / 9)
    assert set(cfg) == {"a", "c"}
    with pytest.raises(NotImplementedError):
        BOREHyperband(cs, retrain=True)
    with pytest.raises(AssertionError):
        BOREHyperband(cs, gamma=1.5)


def test_factory_surface_and_limits_without_gpu():
    """bore/models.py:48-104: constructor asserts, the limits csrc/lstm.cu states, and -- like every other
    entry point -- no CPU fallback: touching the weights without a CUDA device raises."""
    import torch
    from bore_b200 import _lib
    from bore_b200.layers import BinaryCrossentropy
    from bore_b200.models import StackedRecurrentFactory
    fac = StackedRecurrentFactory(4, 1, num_layers=2, num_units=8, layer_kws=dict(activation="elu"))
    assert fac.input_dim == 4 and len(fac.cells) == 2 and fac.final_layer.units == 1
    assert fac._l2() == [0.0] * 8
    with pytest.raises(AssertionError):
        StackedRecurrentFactory(4, 1, layer_kws=dict(return_sequences=True))
    with pytest.raises(AssertionError):
        StackedRecurrentFactory(4, 1, final_layer_kws=dict(activation="sigmoid"))
    for kw in (dict(input_dim=64), dict(num_units=64), dict(num_layers=5), dict(output_dim=2)):
        args = dict(input_dim=4, output_dim=1, num_layers=2, num_units=8)
        args.update(kw)
        with pytest.raises(NotImplementedError):
            StackedRecurrentFactory(**args)
    with pytest.raises(NotImplementedError):
        fac.build_one_to_one(9)
    net = fac.build_many_to_many(mask_value=-1.0)
    with pytest.raises(RuntimeError, match="compile"):
        net.fit(np.zeros((3, 2, 4)), np.zeros((3, 2, 1)))
    with pytest.raises(NotImplementedError):
        net.compile(optimizer="sgd", loss=BinaryCrossentropy(from_logits=True))
    with pytest.raises(NotImplementedError):
        net.compile(optimizer="adam", loss=BinaryCrossentropy(from_logits=False))
    net.compile(optimizer="adam", loss=BinaryCrossentropy(from_logits=True), metrics=["accuracy"])
    with pytest.raises(NotImplementedError):
        net.fit(np.zeros((3, 9, 4)), np.zeros((3, 9, 1)))          # more than 8 rungs
    with pytest.raises(NotImplementedError):
        net.fit(np.zeros((3, 2, 4)), np.zeros((3, 2, 1)), batch_size=128)
    lines = []
    net.summary(print_fn=lines.append)
    assert any("LSTMCell" in l for l in lines) and any("mask_value=-1.0" in l for l in lines)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.BoreNativeError):
            net.fit(np.zeros((3, 2, 4)), np.zeros((3, 2, 1)), verbose=0)
        with pytest.raises(_lib.BoreNativeError):
            fac.build_one_to_one(2).predict(np.zeros((1, 4)))
