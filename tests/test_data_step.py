"""Data step either side of the path: quantile labelling
(bore/data.py:31-35) and the duplicate filter (bore/data.py:42-48).

CPU tests pin the oracle (and the spelled-out lerp formula the kernel implements) to the outputs of
the reference's own ``bore.data.Record`` (tests/golden/data_step_golden.npz); GPU tests hold the
kernels to those fixtures and to the oracle, BIT-EXACT (fp64 / boolean work).
"""
import os

import numpy as np
import pytest

from oracle import data_step as ods

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_step_golden.npz"))
CASES = [int(c) for c in GOLD["cases"]]


def case(ci):
    return {k: GOLD[f"c{ci}/{k}"] for k in ("X", "y", "gamma", "z", "cand", "dup")}


@pytest.mark.parametrize("ci", CASES)
def test_oracle_matches_reference_record(ci):
    c = case(ci)
    z, tau = ods.quantile_labels(c["y"], float(c["gamma"]))
    np.testing.assert_array_equal(z, c["z"])
    for row, t in zip(c["y"], tau):  # the kernel's formula is np.quantile's, bit for bit
        assert ods.lerp_quantile(row, float(c["gamma"])) == t
    np.testing.assert_array_equal(ods.is_duplicate(c["cand"], c["X"]), c["dup"])


def test_lerp_formula_random():
    rs = np.random.RandomState(0)
    for _ in range(300):
        n = rs.randint(1, 70)
        y = np.round(rs.normal(size=n), rs.randint(0, 6))
        q = rs.choice([0.0, 0.25, 1 / 3, 0.5, 0.9, 1.0, rs.uniform()])
        assert ods.lerp_quantile(y, q) == np.quantile(y, q)
    y = np.array([1.0, np.nan, 0.0])
    assert np.isnan(ods.lerp_quantile(y, 0.25)) and np.isnan(np.quantile(y, 0.25))


# ------------------------------------------------------------------------------------ GPU
def _net():
    from bore_b200.engine import NativeMLP
    return NativeMLP([2, 4, 1], ["relu", "sigmoid"])


@pytest.mark.gpu
@pytest.mark.parametrize("ci", CASES)
def test_kernels_match_reference_record(ci):
    from bore_b200.data import quantile_labels
    c = case(ci)
    net = _net()
    z, tau = quantile_labels(c["y"], float(c["gamma"]), net)
    np.testing.assert_array_equal(z, c["z"])
    np.testing.assert_array_equal(tau, ods.quantile_labels(c["y"], float(c["gamma"]))[1])
    keep = net.keep_unique_dev(net.to_device(c["cand"], np.float64), net.to_device(c["X"], np.float64))
    np.testing.assert_array_equal(keep.cpu().numpy() == 0, c["dup"])


@pytest.mark.gpu
def test_quantile_labels_random_sizes_bit_exact():
    from bore_b200.data import quantile_labels
    net = _net()
    rs = np.random.RandomState(3)
    for N in (1, 2, 3, 31, 32, 33, 255, 1024, 1025, 4096, 5000, 16384):
        for q in (0.0, 0.25, 1 / 3, 0.5, 0.77, 1.0):
            y = rs.normal(size=(3, N))
            y[1] = np.round(y[1], 1)
            z, tau = quantile_labels(y, q, net)
            z0, tau0 = ods.quantile_labels(y, q)
            np.testing.assert_array_equal(tau, tau0)
            np.testing.assert_array_equal(z, z0)
    y = rs.normal(size=(2, 50)); y[1, 7] = np.nan          # a NaN poisons its own problem only
    z, tau = quantile_labels(y, 0.25, net)
    z0, tau0 = ods.quantile_labels(y, 0.25)
    assert np.isnan(tau[1]) and tau[0] == tau0[0]
    np.testing.assert_array_equal(z, z0)
    with pytest.raises(Exception, match="range"):
        quantile_labels(y, 1.5, net)
    with pytest.raises(Exception, match="exceeds"):
        quantile_labels(np.zeros((1, 16385)), 0.5, net)


@pytest.mark.gpu
def test_duplicate_edge_cases():
    net = _net()
    prev = np.array([[[0.0, 1.0, np.inf], [1e-9, -2.0, 5.0], [np.nan, 0.0, 0.0]]])
    cand = np.array([[[0.0, 1.0, np.inf],        # inf == inf counts (x == y)
                      [0.0, -2.0, 5.0],          # 1e-9 vs 0: |a-b| <= atol=1e-8 -> duplicate
                      [np.nan, 0.0, 0.0],        # NaN never close (equal_nan=False)
                      [0.0, 1.0 + 2e-5, np.inf], # outside rtol
                      [-0.0, 1.0, np.inf]]])     # -0 == +0
    keep = net.keep_unique_dev(net.to_device(cand, np.float64), net.to_device(prev, np.float64))
    np.testing.assert_array_equal(keep.cpu().numpy() == 0, ods.is_duplicate(cand, prev))


@pytest.mark.gpu
def test_argmax_with_unique_filter_matches_host_filter():
    """argmax(filter_fn=UniqueFilter) (device mask) picks what the reference's scan with the
    host predicate picks (bore/mixins.py:80-87 + bore/plugins/hpbandster/base.py:227-231)."""
    from bore_b200.data import Record, UniqueFilter
    from bore_b200.layers import Dense
    from bore_b200.models import MaximizableSequential
    rs = np.random.RandomState(1)
    m = MaximizableSequential(seed=3)
    m.add(Dense(16, activation="tanh", input_dim=3)); m.add(Dense(1, activation="sigmoid"))
    m.compile(optimizer="adam", loss="binary_crossentropy")
    X = rs.uniform(size=(60, 3)); y = np.sum((X - 0.3) ** 2, axis=1)
    m.fit(X, y < np.quantile(y, 0.25), epochs=20, batch_size=64, verbose=0)
    bounds = [(0.0, 1.0)] * 3
    plain = m.argmax(bounds, num_starts=8, num_samples=64, print_fn=None, random_state=5)
    rec = Record()
    for xi, yi in zip(X, y):
        rec.append(xi, yi)
    rec.append(plain.x, 0.0)                                   # the winner is now a duplicate
    filt = UniqueFilter(rec)
    dev = m.argmax(bounds, filter_fn=filt, num_starts=8, num_samples=64, print_fn=None, random_state=5)
    host = m.argmax(bounds, filter_fn=lambda r: filt(r), num_starts=8, num_samples=64, print_fn=None,
                    random_state=5)
    assert (dev is None) == (host is None)
    if dev is not None:
        np.testing.assert_array_equal(dev.x, host.x)
        assert dev.fun == host.fun and not np.allclose(dev.x, plain.x, rtol=1e-5, atol=1e-8)


@pytest.mark.gpu
def test_batched_fit_from_raw_targets_and_exclude():
    from bore_b200.batched import BatchedMaximizableSequential
    from bore_b200.layers import Dense
    M, N, D = 6, 90, 4
    rs = np.random.RandomState(2)
    X = rs.uniform(size=(M, N, D)); y = np.sum((X - 0.5) ** 2, axis=2) + 0.05 * rs.normal(size=(M, N))
    perms = np.stack([rs.permutation(N) for _ in range(15)])

    def make():
        b = BatchedMaximizableSequential([Dense(16, activation="relu", input_dim=D), Dense(1, activation="sigmoid")],
                                         n_problems=M, seed=11)
        b.compile()
        return b
    a, b = make(), make()
    z = np.stack([row < np.quantile(row, 0.25) for row in y])
    la = a.fit(X, z, batch_size=32, epochs=15, permutations=perms)
    lb = b.fit(X, y, batch_size=32, epochs=15, permutations=perms, gamma=0.25)   # labelled on device
    np.testing.assert_array_equal(la, lb)
    bounds = [(0.0, 1.0)] * D
    Xi = rs.uniform(size=(M, 64, D))
    r0 = a.argmax(bounds, num_starts=4, num_samples=64, X_init=Xi)
    excl = np.concatenate([X, np.stack([r.x for r in r0])[:, None, :]], axis=1)
    r1 = a.argmax(bounds, num_starts=4, num_samples=64, X_init=Xi, exclude=excl)
    for p in range(M):
        assert r1[p] is None or not np.allclose(r1[p].x, r0[p].x, rtol=1e-5, atol=1e-8)
        assert r1[p] is None or r1[p].fun >= r0[p].fun


@pytest.mark.gpu
@pytest.mark.parametrize("D,distortion", [(2, 0.05), (8, 0.01), (50, 0.2)])
def test_truncnorm_distort_matches_scipy_with_the_same_random_state(D, distortion):
    """maybe_distort on the device (bore/base.py:45-64): the reference draws through scipy's
    truncnorm.rvs from the caller's random_state; the batch version consumes the same variates
    in the same order and reproduces the values (1e-9)."""
    from scipy.optimize import Bounds
    from bore_b200.base import maybe_distort, maybe_distort_batch
    from bore_b200.engine import NativeMLP
    net = NativeMLP([D, 4, 1], ["relu", "sigmoid"])
    rs = np.random.RandomState(3)
    P = 37
    loc = rs.uniform(size=(P, D))
    loc[0] = 0.0; loc[1] = 1.0; loc[2, ::2] = 0.0; loc[3] = 1e-12     # suggestions on the faces of the box
    bounds = Bounds(np.zeros(D), np.ones(D))
    rs_ref, rs_dev = np.random.RandomState(11), np.random.RandomState(11)
    ref = np.stack([maybe_distort(loc[p], distortion, bounds, rs_ref, print_fn=lambda s: None) for p in range(P)])
    got = maybe_distort_batch(loc, distortion, bounds, rs_dev, net=net)
    assert got.shape == (P, D)
    assert np.all(got >= 0.0) and np.all(got <= 1.0)
    assert np.abs(got - ref).max() <= 1e-9, np.abs(got - ref).max()
    assert rs_ref.uniform() == rs_dev.uniform()                      # both consumed the same stream
    assert np.array_equal(maybe_distort_batch(loc, None), loc)       # distortion=None: untouched
