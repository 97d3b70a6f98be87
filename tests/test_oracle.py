"""The oracle itself: cross-checked against torch-CPU autograd (a second, independent
implementation of the same maths) and against the committed golden fixtures.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import keras_mlp as km, argmax as am
from helpers import NETS, trained_weights

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _torch_forward(w, acts, X):
    h = X
    for l, a in enumerate(acts):
        h = h @ torch.tensor(w[2 * l], dtype=torch.float64) + torch.tensor(w[2 * l + 1], dtype=torch.float64)
        h = {"linear": lambda t: t, "relu": torch.relu, "elu": torch.nn.functional.elu,
             "sigmoid": torch.sigmoid, "tanh": torch.tanh}[a](h)
    return h


@pytest.mark.parametrize("name", list(NETS))
def test_value_and_grad_vs_torch_autograd(name):
    dims, acts, transform = NETS[name]
    w = trained_weights(dims, acts, seed=1, N=120, epochs=8)
    X = np.random.RandomState(0).uniform(-0.2, 1.2, size=(40, dims[0]))
    Xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    u = _torch_forward(w, acts, Xt)
    T = {"identity": lambda t: t, "sigmoid": torch.sigmoid, "exp": torch.exp}[transform]
    f = T(-u)[:, 0]
    f.sum().backward()
    f64, g64 = km.value_and_input_grad(w, acts, X, transform, True, np.float64)
    assert np.allclose(f64, f.detach().numpy(), rtol=1e-12, atol=1e-14)
    assert np.allclose(g64, Xt.grad.numpy(), rtol=1e-10, atol=1e-13)
    f32, g32 = km.value_and_input_grad(w, acts, X, transform, True, np.float32)
    assert f32.dtype == np.float32
    assert np.allclose(f32, f64, rtol=2e-5, atol=2e-6)
    assert np.abs(g32 - g64).max() <= 2e-5 * max(1.0, np.abs(g64).max())


@pytest.mark.parametrize("name,l2", [("cfg2_hartmann6", 0.0), ("cfg5_plugin8", 1e-3)])
def test_weight_grads_and_adam_vs_torch(name, l2):
    dims, acts, _ = NETS[name]
    w = km.init_weights(dims, 4, np.float64)
    rs = np.random.RandomState(1)
    X = rs.uniform(size=(50, dims[0])); z = (rs.uniform(size=50) < 0.3)
    loss, grads = km.loss_and_weight_grads(w, acts, X, z, l2, np.float64)
    wt = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in w]
    h = torch.tensor(X)
    acts_l = list(acts[:-1]) + ["linear"]
    for l, a in enumerate(acts_l):
        h = h @ wt[2 * l] + wt[2 * l + 1]
        h = {"linear": lambda t: t, "relu": torch.relu, "elu": torch.nn.functional.elu}[a](h)
    lt = torch.nn.functional.binary_cross_entropy_with_logits(h[:, 0], torch.tensor(z, dtype=torch.float64))
    lt = lt + l2 * sum((p ** 2).sum() for p in wt)
    lt.backward()
    assert abs(loss - lt.item()) <= 1e-12
    for g, p in zip(grads, wt):
        assert np.allclose(g, p.grad.numpy(), rtol=1e-9, atol=1e-12)
    # one Keras-Adam step by hand: t=1 -> alpha = lr*sqrt(1-b2)/(1-b1), m=(1-b1)g, v=(1-b2)g^2
    st = km.AdamState(w)
    w2 = [a.copy() for a in w]
    km.adam_apply(w2, grads, st, np.float64)
    alpha = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    for a, b, g in zip(w, w2, grads):
        m, v = 0.1 * g, 0.001 * g * g
        assert np.allclose(b, a - alpha * m / (np.sqrt(v) + 1e-7), rtol=1e-12, atol=1e-15)
    assert st.t == 1


def test_fit_semantics():
    """Ragged last batch kept; epoch loss is the sample-weighted mean of pre-update losses."""
    dims, acts, _ = NETS["cfg1_branin"]
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(100, 2)); z = rs.uniform(size=100) < 0.25
    w = km.init_weights(dims, 0)
    w_before = [a.copy() for a in w]
    perms = np.stack([rs.permutation(100)])
    hist, adam = km.fit(w, acts, X, z, 1, 64, perms)
    assert adam.t == 2  # ceil(100/64)
    l1, g1 = km.loss_and_weight_grads(w_before, acts, X[perms[0][:64]].astype(np.float32),
                                      z[perms[0][:64]].astype(np.float32))
    w_mid = [a.copy() for a in w_before]
    km.adam_apply(w_mid, g1, km.AdamState(w_mid))
    l2_, _ = km.loss_and_weight_grads(w_mid, acts, X[perms[0][64:]].astype(np.float32),
                                      z[perms[0][64:]].astype(np.float32))
    assert abs(hist[0] - (l1 * 64 + l2_ * 36) / 100) <= 1e-6


def test_dense_sequential_layer_quirk():
    assert km.dense_sequential_dims(8, 1, 2, 32) == [8, 32, 32, 32, 1]
    assert km.dense_sequential_dims(2, 1, 1, 16) == [2, 16, 16, 1]
    assert km.dense_sequential_dims(2, 1, 0, 16) == [2, 1]


def test_lockstep_driver_is_scipy_minimize():
    """The setulb lock-step driver reproduces scipy.optimize.minimize bit for bit."""
    from scipy.optimize import Bounds
    dims, acts, transform = NETS["cfg2_hartmann6"]
    w = trained_weights(dims, acts, seed=0, N=120, epochs=8)
    X0 = np.random.RandomState(3).uniform(size=(10, 6))
    ref = am.minimize_starts(w, acts, X0, Bounds(np.zeros(6), np.ones(6)), transform=transform)
    ls = am.LockstepLBFGSB(X0, np.zeros(6), np.ones(6))
    while ls.pending.any():
        pts = ls.X[ls.pending]
        fg = [km.value_and_input_grad(w, acts, p[None], transform) for p in pts]
        ls.feed([a[0][0] for a in fg], [a[1][0] for a in fg])
    got = ls.result()
    for k in ("x", "fun", "nit", "nfev", "status"):
        assert np.array_equal(got[k], ref[k]), k


def test_oracle_reproduces_lbfgsb_golden():
    """Golden vectors minted from SciPy (tests/golden/make_golden.py) still reproduce: guards
    both the oracle MLP and the installed SciPy against drift."""
    import scipy
    from scipy.optimize import Bounds
    g = np.load(os.path.join(GOLD, "lbfgsb_golden.npz"))
    for name in ("cfg1_branin", "cfg5_plugin8"):
        dims, acts, transform = NETS[name]
        w = [g[f"{name}/w{i}"] for i in range(2 * len(acts))]
        X0 = g[name + "/X0"]
        f0, g0 = km.value_and_input_grad(w, acts, X0, transform, True, np.float32)
        assert np.allclose(f0, g[name + "/f0"], rtol=1e-6, atol=1e-7)
        assert np.allclose(g0, g[name + "/g0"], rtol=1e-5, atol=1e-7)
        if str(g["scipy_version"]) == scipy.__version__:
            n = dims[0]
            r = am.minimize_starts(w, acts, X0[:8], Bounds(np.zeros(n), np.ones(n)), transform=transform)
            assert np.abs(r["fun"] - g[name + "/fun"][:8]).max() <= 1e-6


def test_oracle_reproduces_fit_golden():
    g = np.load(os.path.join(GOLD, "fit_golden.npz"))
    for name in ("cfg1_branin", "cfg5_plugin8"):
        dims, acts, _ = NETS[name]
        w, o = [], 0
        flat = g[name + "/w0"]
        for fi, fo in zip(dims[:-1], dims[1:]):
            w.append(flat[o:o + fi * fo].reshape(fi, fo).copy()); o += fi * fo
            w.append(flat[o:o + fo].copy()); o += fo
        perms = g[name + "/perms"]
        hist, _ = km.fit(w, acts, g[name + "/X"], g[name + "/z"], len(perms), 64, perms,
                         l2=float(g[name + "/l2"]))
        assert np.abs(hist - g[name + "/loss"]).max() <= 1e-5


# ---------------------------------------------------------------------------- the whole fit loop
def _torch_fit(w0, acts, X, z, epochs, batch, perms, l2=0.0, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """An independent implementation of what ``oracle.keras_mlp.fit`` restates: gradients by
    autograd (not the hand-written backward pass), torch's own binary_cross_entropy_with_logits,
    Adam in the textbook form  m = b1 m + (1 - b1) g,  v = b2 v + (1 - b2) g^2,
    w -= lr_t m / (sqrt(v) + eps)  with  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)  (Keras: eps is
    OUTSIDE the bias correction), epoch loss = sum(batch loss * batch size) / N.  fp64."""
    F = torch.nn.functional
    act_fn = {"linear": lambda t: t, "relu": torch.relu, "elu": F.elu, "sigmoid": torch.sigmoid,
              "tanh": torch.tanh}
    ws = [torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in w0]
    m = [torch.zeros_like(a) for a in ws]
    v = [torch.zeros_like(a) for a in ws]
    Xt, zt = torch.tensor(np.asarray(X, np.float64)), torch.tensor(np.asarray(z, np.float64)).reshape(-1, 1)
    N, t, hist = Xt.shape[0], 0, []
    for e in range(epochs):
        tot = 0.0
        perm = torch.as_tensor(np.asarray(perms[e], np.int64))
        for s in range(0, N, batch):
            idx = perm[s:s + batch]
            h = Xt[idx]
            for l, a in enumerate(acts):
                h = h @ ws[2 * l] + ws[2 * l + 1]
                if l < len(acts) - 1:
                    h = act_fn[a](h)
            loss = F.binary_cross_entropy_with_logits(h, zt[idx], reduction="mean")  # on the logits
            if l2:
                loss = loss + sum(l2 * (a * a).sum() for a in ws)
            grads = torch.autograd.grad(loss, ws)
            tot += float(loss) * len(idx)
            t += 1
            lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            with torch.no_grad():
                for i, g in enumerate(grads):
                    m[i] = b1 * m[i] + (1.0 - b1) * g
                    v[i] = b2 * v[i] + (1.0 - b2) * g * g
                    ws[i] -= lr_t * m[i] / (torch.sqrt(v[i]) + eps)
        hist.append(tot / N)
    return np.array(hist), [a.detach().numpy() for a in ws], t


@pytest.mark.parametrize("name,l2", [("cfg2_hartmann6", 0.0), ("cfg5_plugin8", 0.0), ("cfg5_plugin8", 1e-4),
                                     ("tanh_exp", 0.0), ("cfg1_branin", 0.0)])
def test_fit_loop_vs_independent_torch_implementation(name, l2):
    """>= 300 Adam steps in fp64: loss trajectory and final weights to 1e-10.  Narrows the
    unpinned half of the oracle to the [TF-semantics] RULES (which formula Keras uses), away from
    their implementation here."""
    dims, acts, _ = NETS[name]
    acts = list(acts[:-1]) + ["sigmoid" if acts[-1] == "sigmoid" else "linear"]
    rs = np.random.RandomState(12)
    N, E, B = 150, 100, 64                      # 3 steps per epoch (the last batch is short): 300 steps
    X = rs.uniform(size=(N, dims[0]))
    y = np.sum((X - 0.45) ** 2, axis=1)
    z = y < np.quantile(y, 0.3)
    perms = np.stack([rs.permutation(N) for _ in range(E)])
    w0 = km.init_weights(dims, 3, np.float64)
    w = [a.copy() for a in w0]
    hist, adam = km.fit(w, acts, X, z, E, B, perms, l2=l2, dtype=np.float64)
    hist_t, w_t, t = _torch_fit(w0, acts, X, z, E, B, perms, l2=l2)
    assert adam.t == t == 300
    assert np.abs(hist - hist_t).max() <= 1e-10, np.abs(hist - hist_t).max()
    assert max(np.abs(a - b).max() for a, b in zip(w, w_t)) <= 1e-10
    assert hist[-1] < hist[0]
