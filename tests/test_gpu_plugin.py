"""HpBandSter config generator end to end on the GPU (BASELINE.json configs[4]): synthetic 8-D
hyperparameter benchmark, BOHB-style brackets, pooled and per-budget classifiers."""
import logging

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _space():
    from bore_b200.plugins.hpbandster._compat import CS
    cs = CS.ConfigurationSpace(seed=3)
    for i in range(5):
        cs.add_hyperparameter(CS.UniformFloatHyperparameter(f"x{i}", lower=-1.0, upper=1.0))
    cs.add_hyperparameter(CS.UniformIntegerHyperparameter("k", lower=1, upper=8))
    cs.add_hyperparameter(CS.CategoricalHyperparameter("act", ["a", "b"]))
    return cs  # dense dimension 5 + 1 + 2 = 8


def _loss(cfg, budget):
    x = np.array([cfg[f"x{i}"] for i in range(5)])
    base = np.sum((x - 0.3) ** 2) + 0.05 * abs(cfg["k"] - 5) + (0.2 if cfg["act"] == "a" else 0.0)
    return base + 0.05 / budget * np.sin(37.0 * np.sum(x))  # budget-dependent "noise"


@pytest.mark.parametrize("per_budget", [False, True])
def test_bore_hyperband_improves_over_random(per_budget):
    from bore_b200.plugins.hpbandster import BORE
    opt = BORE(_space(), eta=3, min_budget=1 / 9, max_budget=1, seed=0, num_random_init=8,
               num_steps_per_iter=200, num_starts=4, num_samples=256, per_budget=per_budget,
               logger=logging.getLogger("bore-test"))
    cg = opt.config_generator
    assert cg.input_dim == 8
    results = opt.run(n_iterations=6, compute_fn=_loss)
    assert len(results) >= 40
    full = [l for _, b, l in results if b == 1.0]
    first = np.mean([l for _, _, l in results[:8]])
    assert min(full) < first
    # the generator did train and maximise a classifier on the device
    logit = cg.logit if not per_budget else next(v for v in cg._budget_logits.values() if v is not None)
    assert logit is not None and logit._net is not None
    assert logit._last_stats["evals"] > 0
    # proposals decode to legal configurations
    for cfg, _, _ in results:
        assert 1 <= cfg["k"] <= 8 and cfg["act"] in ("a", "b")
        assert all(-1.0 <= cfg[f"x{i}"] <= 1.0 for i in range(5))


def test_duplicate_filter_and_fallback():
    """filter_fn=_is_unique drops maxima already in the record; None -> random candidate."""
    from bore_b200.plugins.hpbandster import BORE
    opt = BORE(_space(), eta=3, min_budget=1 / 9, max_budget=1, seed=1, num_random_init=6,
               num_steps_per_iter=100, num_starts=2, num_samples=64, random_rate=None)
    cg = opt.config_generator
    rs = np.random.RandomState(0)
    from bore_b200.plugins.hpbandster._compat import Job
    for j in range(8):
        cfg = cg.config_space.sample_configuration().get_dictionary()
        job = Job((0, 0, j), cfg, 1.0)
        job.result = {"loss": float(_loss(cfg, 1.0))}
        cg.new_result(job)
    cfg, _ = cg.get_config(1.0)
    assert cg.record.size() == 8 and set(cfg) == {"x0", "x1", "x2", "x3", "x4", "k", "act"}
    # force every candidate to be a "duplicate": the generator must fall back to a random config
    cg.record.is_duplicate = lambda x, **kw: True
    cfg2, _ = cg.get_config(1.0)
    assert set(cfg2) == set(cfg)
