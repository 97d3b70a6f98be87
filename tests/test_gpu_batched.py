"""Batched BO problems (BASELINE.json configs[3]): M independent classifiers trained and maximised
in the same launches must each behave like the reference's single-problem fit + argmax."""
import numpy as np
import pytest

from oracle import keras_mlp as km, argmax as am
from helpers import synthetic_targets

pytestmark = pytest.mark.gpu


def _problems(M, N, D, seed):
    rs = np.random.RandomState(seed)
    X = rs.uniform(size=(M, N, D))
    y = np.stack([synthetic_targets(X[p]) + 0.3 * p * X[p, :, 0] for p in range(M)])
    z = np.stack([y[p] < np.quantile(y[p], 1 / 3) for p in range(M)])
    return X, z


def test_batched_fit_and_argmax_match_the_oracle_per_problem():
    from scipy.optimize import Bounds
    from bore_b200 import BatchedMaximizableSequential, Dense, BinaryCrossentropy, sigmoid
    M, N, D, E = 6, 150, 6, 8
    dims, acts = [D, 32, 32, 32, 1], ["elu", "elu", "elu", "linear"]  # the plugin's network
    X, z = _problems(M, N, D, seed=0)
    layers = [Dense(32, activation="elu", input_dim=D), Dense(32, activation="elu"),
              Dense(32, activation="elu"), Dense(1)]
    model = BatchedMaximizableSequential(layers, n_problems=M, transform=sigmoid, seed=1)
    model.compile(optimizer="adam", loss=BinaryCrossentropy(from_logits=True))
    w0 = [km.init_weights(dims, 10 + p) for p in range(M)]
    model.set_weights(w0)
    perms = np.stack([np.random.RandomState(3).permutation(N) for _ in range(E)])
    hist = model.fit(X, z, batch_size=64, epochs=E, permutations=perms)
    assert hist.shape == (M, E)
    w_ref = []
    for p in range(M):
        w = [a.copy() for a in w0[p]]
        h_ref, _ = km.fit(w, acts, X[p], z[p], E, 64, perms)
        assert np.abs(hist[p] - h_ref).max() <= 1e-4, p
        w_ref.append(w)
    got_w = model.get_weights()
    for p in range(M):
        for a, b in zip(got_w[p], w_ref[p]):
            np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-5)

    # predict per problem
    Xq = np.random.RandomState(4).uniform(size=(M, 40, D))
    pred = model.predict(Xq)
    for p in range(M):
        np.testing.assert_allclose(pred[p], km.predict(got_w[p], acts, Xq[p]), rtol=1e-5, atol=1e-6)

    # argmax per problem, the screening samples drawn problem after problem from one stream
    bounds = Bounds(np.zeros(D), np.ones(D))
    res = model.argmax(bounds, num_starts=4, num_samples=128, random_state=np.random.RandomState(5))
    rs = np.random.RandomState(5)
    assert len(res) == M
    for p in range(M):
        ref = am.argmax(got_w[p], acts, bounds, num_starts=4, num_samples=128,
                        print_fn=lambda s: None, random_state=rs, transform="sigmoid")
        assert (res[p] is None) == (ref is None)
        if ref is not None:
            assert abs(float(res[p].fun) - float(ref.fun)) <= 1e-4, (p, res[p].fun, ref.fun)
            assert np.all(res[p].x >= 0) and np.all(res[p].x <= 1)
    assert model._last_stats["evals"] >= M * 4


def test_batched_all_samples_are_starts_and_many_problems():
    """num_starts == num_samples path, and more problems than SMs."""
    from bore_b200 import BatchedMaximizableSequential, Dense, identity
    M, D = 300, 3
    layers = [Dense(16, activation="relu", input_dim=D), Dense(16, activation="relu"),
              Dense(1, activation="sigmoid")]
    model = BatchedMaximizableSequential(layers, n_problems=M, transform=identity, seed=2)
    model.compile(optimizer="adam", loss="binary_crossentropy")
    X, z = _problems(M, 64, D, seed=1)
    hist = model.fit(X, z, batch_size=32, epochs=3)
    assert hist.shape == (M, 3) and np.all(np.isfinite(hist))
    res = model.argmax([(0.0, 1.0)] * D, num_starts=3, num_samples=3, random_state=0)
    assert len(res) == M
    ws = model.get_weights()
    acts = ["relu", "relu", "sigmoid"]
    ok = 0
    for p in range(0, M, 37):
        r = res[p]
        if r is None:
            continue
        f_chk, _ = km.value_and_input_grad(ws[p], acts, r.x[None], "identity", True, np.float32)
        assert abs(float(f_chk[0]) - float(r.fun)) <= 1e-5
        ok += 1
    assert ok > 0


def test_batched_argmax_distortion_is_maybe_distort_per_problem():
    """``distortion=`` resamples every problem's winner exactly like the plugin's host call
    (bore/plugins/hpbandster/base.py:266 -> bore/base.py:51-64) with the same random_state."""
    from scipy.optimize import Bounds
    from bore_b200 import BatchedMaximizableSequential, Dense, maybe_distort
    M, D = 5, 6
    dims = [D, 32, 32, 1]
    layers = [Dense(32, activation="relu", input_dim=D), Dense(32, activation="relu"), Dense(1, activation="sigmoid")]
    model = BatchedMaximizableSequential(layers, n_problems=M, seed=2)
    model.compile(optimizer="adam", loss="binary_crossentropy")
    model.set_weights([km.init_weights(dims, 40 + p) for p in range(M)])
    X_init = np.random.RandomState(1).uniform(size=(M, 64, D))
    bounds = [(0.0, 1.0)] * D
    plain = model.argmax(bounds, num_starts=4, num_samples=64, X_init=X_init)
    dist = model.argmax(bounds, num_starts=4, num_samples=64, X_init=X_init, distortion=0.05,
                        random_state=np.random.RandomState(9))
    rs = np.random.RandomState(9)
    b = Bounds(np.zeros(D), np.ones(D))
    for p in range(M):
        want = maybe_distort(plain[p].x, 0.05, b, rs, print_fn=lambda s: None)
        assert np.abs(dist[p].x - want).max() <= 1e-9
        assert np.all(dist[p].x >= 0) and np.all(dist[p].x <= 1)
        assert dist[p].fun == plain[p].fun      # the record still describes the optimum found
