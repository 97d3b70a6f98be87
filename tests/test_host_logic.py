"""Host-side mirror of the reference surface, checked against known answers produced by the
reference's own modules (tests/golden/host_golden.json).  CPU only: no kernels are launched."""
import doctest
import json
import os

import numpy as np
import pytest
from scipy.optimize import Bounds

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "host_golden.json")) as f:
        return json.load(f)


def test_math_against_reference(gold):
    import bore_b200.math as bm
    for n, b, want in gold["steps_per_epoch"]:
        assert bm.steps_per_epoch(n, b) == want
        assert isinstance(bm.steps_per_epoch(n, b), int)
    for a, b, want in gold["ceil_divide"]:
        assert int(bm.ceil_divide(a, b)) == want
    assert doctest.testmod(bm).failed == 0  # same function as the doctests of bore/math.py:15-27


def test_from_bounds_against_reference(gold):
    from bore_b200.optimizers.utils import from_bounds
    for case in gold["from_bounds"]:
        (lo, hi), d = from_bounds(Bounds(np.array(case["lo"]), np.array(case["hi"])))
        assert [list(map(float, lo)), list(map(float, hi)), d] == case["bounds_obj"]
        (lo, hi), d = from_bounds(list(zip(case["lo"], case["hi"])))
        assert [list(map(float, lo)), list(map(float, hi)), d] == case["pairs"]
        assert isinstance(lo, tuple)


def test_record_against_reference(gold):
    from bore_b200.data import Record
    for case in gold["record"]:
        rec = Record()
        for xi, yi in zip(case["X"], case["y"]):
            rec.append(x=np.array(xi), y=yi, b=1.0)
        assert rec.size() == case["size"] == len(rec.budgets)
        X, z = rec.load_classification_data(case["gamma"])
        assert X.shape == (case["n"], 3)
        assert [bool(v) for v in z] == case["z"]
        got = [rec.is_duplicate(np.array(p)) for p in case["probes"]]
        assert got == case["is_duplicate"]


class _FakeModel:
    """Stands in for a native model: records what the tracer asked the kernel to do."""

    def __init__(self, D):
        self.D, self.calls = D, []

    def __call__(self, x):
        from bore_b200 import ops
        assert isinstance(x, ops.Tracer)
        return ops.Expr(self, x.shape, tuple(x.shape[:-1]) + (1,))

    def _native_value_and_grad(self, X, transform, negate):
        self.calls.append((X.shape, transform, negate))
        return np.arange(X.shape[0], dtype=np.float32), np.ones((X.shape[0], self.D), np.float32)


def test_convert_shape_algebra_and_tracing():
    """bore/base.py:35-38: (D,) -> [(), (D,)] as a list; batches (S, D) -> [(S,), (S, D)];
    transform(-u) is traced into (transform, negate)."""
    from bore_b200 import convert, ops
    m = _FakeModel(4)
    for name, fn in ops.TRANSFORMS.items():
        f = convert(m, transform=lambda u, fn=fn: fn(-u))
        out = f(np.zeros(4))
        assert isinstance(out, list) and out[0].shape == () and out[1].shape == (4,)
        assert out[1].dtype == np.float64 and out[0].dtype == np.float32
        assert m.calls[-1] == ((1, 4), name, True)
        out = f(np.zeros((7, 4), np.float32))
        assert out[0].shape == (7,) and out[1].shape == (7, 4) and out[1].dtype == np.float32
    f = convert(m)  # the `_func_max` form: identity, no negation
    f(np.zeros(4))
    assert m.calls[-1] == ((1, 4), "identity", False)
    with pytest.raises(NotImplementedError):
        convert(m, transform=lambda u: -ops.sigmoid(u))(np.zeros(4))
    with pytest.raises(TypeError):
        convert(m, transform=lambda u: 3.0)(np.zeros(4))


def test_decorators_on_arrays():
    from bore_b200 import decorators as d, ops
    assert d.stack(lambda a: a.sum())(1.0, 2.0, 3.0) == 6.0
    assert d.unstack(lambda a, b: a - b)(np.array([[5.0, 2.0], [1.0, 1.0]])).tolist() == [3.0, 0.0]
    assert d.squeeze(axis=-1)(lambda a: a)(np.zeros((3, 1))).shape == (3,)
    assert d.unbatch(lambda a: a * 2)(np.ones(3)).shape == (3,)
    assert np.allclose(ops.sigmoid(np.array([0.0])), 0.5) and np.allclose(ops.exp(np.array([0.0])), 1.0)


def test_model_spec_without_gpu():
    """Layer bookkeeping, compile-time validation and the DenseSequential quirk need no device."""
    from bore_b200.models import MaximizableDenseSequential, MaximizableSequential
    from bore_b200.layers import Dense, BinaryCrossentropy, l2
    m = MaximizableDenseSequential(input_dim=8, output_dim=1, num_layers=2, num_units=32,
                                   layer_kws=dict(activation="elu", kernel_regularizer=l2(1e-3),
                                                  bias_regularizer=l2(1e-3)))
    assert [l.units for l in m.layers] == [32, 32, 32, 1]
    assert m.layers[0].input_dim == 8 and m.count_params() == 2433
    assert m._l2() == [1e-3] * 6 + [0.0, 0.0]  # hidden layers only, like the plugin
    m.compile(optimizer="adam", metrics=["accuracy"], loss=BinaryCrossentropy(from_logits=True))
    lines = []
    m.summary(print_fn=lines.append)
    assert any("2433" in s for s in lines)
    r = MaximizableSequential()
    r.add(Dense(16, activation="relu")); r.add(Dense(1, activation="sigmoid"))
    r.compile(optimizer="adam", loss="binary_crossentropy")
    with pytest.raises(NotImplementedError):
        r.compile(optimizer="sgd", loss="binary_crossentropy")
    with pytest.raises(NotImplementedError):
        r.compile(optimizer="adam", loss=BinaryCrossentropy(from_logits=True))  # sigmoid head
    with pytest.raises(NotImplementedError):
        r.add("not a layer")
    with pytest.raises(ValueError):
        Dense(4, activation="swish")
    from bore_b200 import ops
    assert r._min_transform_name() == "identity"
    assert MaximizableSequential(transform=ops.sigmoid)._min_transform_name() == "sigmoid"


def test_maybe_distort():
    from bore_b200.base import maybe_distort
    loc = np.array([0.2, 0.9])
    assert maybe_distort(loc) is loc
    b = Bounds(np.zeros(2), np.ones(2))
    msgs = []
    out = maybe_distort(loc, 0.05, b, np.random.RandomState(0), print_fn=msgs.append)
    assert out.shape == (2,) and np.all(out >= 0) and np.all(out <= 1) and len(msgs) == 1
    with pytest.raises(AssertionError):
        maybe_distort(loc, 0.05)


# ---------------------------------------------------------------- host-side containers (no GPU)
def test_batched_results_is_a_lazy_sequence_of_optimize_results():
    from bore_b200.batched import BatchedResults
    dim = 3
    #        x0   x1   x2   fun  nit nfev status task key
    rec = np.array([[.1, .2, .3, -0.5, 7, 9, 0, 1, 5.0],
                    [.4, .5, .6, -0.1, 3, 4, 1, 3, 9.0],
                    [.0, .0, .0, 0.0, 0, 0, 2, 0, 0.0]])      # key 0: no start qualified
    res = BatchedResults(rec, dim)
    assert len(res) == 3 and list(res.found) == [True, True, False]
    r0, r1, r2 = res[0], res[1], res[-1]
    assert r2 is None and res[2] is None
    assert np.array_equal(r0.x, [.1, .2, .3]) and r0.fun == np.float32(-0.5) and r0.nit == 7 and r0.nfev == 9
    assert r0.success and r0.status == 0 and not r1.success and r1.status == 1
    assert isinstance(r0.message, str) and r0.x is not res.x          # a copy, not a view
    assert [r is None for r in res] == [False, False, True]
    assert len(res[0:2]) == 2
    with pytest.raises(IndexError):
        res[3]


def test_lazy_loss_behaves_like_its_array():
    import torch
    from bore_b200.batched import LazyLoss
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    l = LazyLoss(torch.from_numpy(a.copy()))
    assert l.shape == (2, 3) and len(l) == 2
    np.testing.assert_array_equal(l, a)
    np.testing.assert_array_equal(l[:, 0], a[:, 0])
    assert np.all(np.isfinite(l)) and np.asarray(l, np.float64).dtype == np.float64
    assert [row.tolist() for row in l] == a.tolist()


def test_unique_filter_host_predicate_is_the_reference_rule():
    """UniqueFilter called as a function = not Record.is_duplicate (bore/plugins/hpbandster/base.py:227-231)."""
    from scipy.optimize import OptimizeResult
    from bore_b200.data import Record, UniqueFilter
    rec = Record()
    f = UniqueFilter(rec)
    assert f.stored().shape[0] == 0 and f(OptimizeResult(x=np.array([.5, .5])))
    rec.append(np.array([.5, .5]), 1.0)
    rec.append(np.array([.1, .9]), 2.0)
    assert f.stored().shape == (2, 2)
    assert not f(OptimizeResult(x=np.array([.5, .5 + 1e-9])))
    assert f(OptimizeResult(x=np.array([.5, .5 + 1e-3])))


def test_svgd_host_pieces_match_the_reference_definitions():
    """rank / distortions (bore/optimizers/svgd/base.py:11-64) are host helpers; the SVGD arithmetic
    itself is device-only (tests/test_svgd.py)."""
    from bore_b200.optimizers.svgd.base import (SVGD, DistortionConstant, DistortionExpDecay, rank)
    from bore_b200.optimizers.svgd.kernels import RadialBasis
    from oracle import svgd as osv
    a = np.random.RandomState(0).rand(17)
    a[3] = a[5]
    np.testing.assert_array_equal(rank(a), osv.rank(a))
    assert DistortionConstant(2.5)(rank(a)) == 2.5
    np.testing.assert_array_equal(DistortionExpDecay(0.7)(rank(a)), np.power(rank(a), -0.7))
    s = SVGD(kernel=RadialBasis(length_scale=None), distortion=DistortionExpDecay(2.0))
    o = s._device_options()
    assert np.isnan(o["length_scale"]) and o["lambd"] == 2.0
    o = SVGD(kernel=RadialBasis(0.5), distortion=DistortionConstant(3.0))._device_options()
    assert o["length_scale"] == 0.5 and np.isnan(o["lambd"]) and o["zeta_c"] == 3.0
    with pytest.raises(NotImplementedError):
        SVGD(kernel=object())._device_options()
