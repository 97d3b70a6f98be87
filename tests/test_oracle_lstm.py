"""The LSTM half of the oracle (oracle/keras_lstm.py) against an independent torch-CPU autograd
implementation of the same stacked-LSTM classifier: forward, input gradient of the one-to-one
network, weight gradients of the masked many-to-many loss, and the whole fit loop.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import keras_lstm as kl, keras_mlp as km

ACT = {"tanh": torch.tanh, "elu": torch.nn.functional.elu, "relu": torch.relu, "sigmoid": torch.sigmoid,
       "linear": lambda t: t}


def _torch_forward(ws, activation, X, mask):
    """X (B, T, D) torch fp64, mask (B, T) bool -> logits (B, T)."""
    act = ACT[activation]
    L = (len(ws) - 2) // 3
    U = ws[1].shape[0]
    B, T, _ = X.shape
    seq = X
    for l in range(L):
        K, R, b = ws[3 * l], ws[3 * l + 1], ws[3 * l + 2]
        h = torch.zeros(B, U, dtype=X.dtype)
        c = torch.zeros(B, U, dtype=X.dtype)
        outs = []
        for t in range(T):
            z = seq[:, t] @ K + h @ R + b
            i, f, g, o = torch.sigmoid(z[:, :U]), torch.sigmoid(z[:, U:2 * U]), act(z[:, 2 * U:3 * U]), torch.sigmoid(z[:, 3 * U:])
            c_new = f * c + i * g
            h_new = o * act(c_new)
            m = mask[:, t][:, None]
            h = torch.where(m, h_new, h)
            c = torch.where(m, c_new, c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    return (seq @ ws[-2] + ws[-1])[..., 0]


def _problem(seed, N=48, T=4, D=5, mask_value=1e-9):
    rs = np.random.RandomState(seed)
    X = rs.uniform(size=(N, T, D))
    Y = (rs.uniform(size=(N, T, 1)) < 0.4).astype(np.float64)
    # ragged sequences: a sample is observed on a prefix of the rungs, some with a gap
    for n in range(N):
        keep = rs.randint(1, T + 1)
        X[n, keep:] = mask_value
        Y[n, keep:] = mask_value
        if n % 7 == 3 and keep > 2:
            X[n, 1] = mask_value
    return X, Y


@pytest.mark.parametrize("activation", ["tanh", "elu"])
def test_forward_and_weight_gradients_vs_torch(activation):
    D, U, L, mv = 5, 8, 2, 1e-9
    X, Y = _problem(0, D=D, mask_value=mv)
    w = kl.init_weights(D, U, L, seed=1, dtype=np.float64)
    mask = kl.compute_mask(X, mv)
    assert not mask.all() and mask.any()
    loss, grads = kl.loss_and_weight_grads(w, activation, X, Y, mask, l2=None, dtype=np.float64)
    wt = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in w]
    u = _torch_forward(wt, activation, torch.tensor(X), torch.tensor(mask))
    y = torch.tensor(Y[..., 0])
    per = torch.nn.functional.binary_cross_entropy_with_logits(u, y, reduction="none")
    loss_t = (per * torch.tensor(mask, dtype=torch.float64)).sum() / u.numel()
    gt = torch.autograd.grad(loss_t, wt)
    assert abs(float(loss) - float(loss_t)) <= 1e-12
    assert np.allclose(kl.forward(w, activation, X, mask, np.float64), u.detach().numpy(), rtol=1e-12, atol=1e-14)
    for a, b in zip(grads, gt):
        assert np.abs(a - b.numpy()).max() <= 1e-12 * max(1.0, np.abs(b.numpy()).max())


@pytest.mark.parametrize("activation,transform", [("tanh", "identity"), ("elu", "sigmoid"), ("elu", "exp")])
def test_one_to_one_value_and_input_gradient_vs_torch(activation, transform):
    D, U, L, T = 4, 8, 2, 3
    w = kl.init_weights(D, U, L, seed=2, dtype=np.float64)
    X = np.random.RandomState(3).uniform(size=(20, D))
    f, g = kl.value_and_input_grad(w, activation, X, T, transform, True, np.float64)
    Xt = torch.tensor(X, requires_grad=True)
    wt = [torch.tensor(a) for a in w]
    Xr = Xt[:, None, :].repeat(1, T, 1)
    u = _torch_forward(wt, activation, Xr, torch.ones(20, T, dtype=torch.bool))[:, -1]
    Tf = {"identity": lambda t: t, "sigmoid": torch.sigmoid, "exp": torch.exp}[transform]
    ft = Tf(-u)
    ft.sum().backward()
    assert np.allclose(f, ft.detach().numpy(), rtol=1e-12, atol=1e-14)
    assert np.allclose(g, Xt.grad.numpy(), rtol=1e-10, atol=1e-13)
    # the last step of the many-to-many network IS the one-to-one network (tests/test_models.py:96)
    Xr_np = np.repeat(X[:, None, :], T, axis=1)
    assert np.array_equal(kl.forward(w, activation, Xr_np, None, np.float64)[:, -1:],
                          kl.predict_one_to_one(w, activation, X, T, np.float64))


def test_fit_loop_vs_independent_torch_implementation():
    D, U, L, mv = 5, 8, 2, 1e-9
    X, Y = _problem(4, N=40, D=D, mask_value=mv)
    E, B = 40, 16                           # 3 steps per epoch (last batch of 8): 120 Adam steps
    rs = np.random.RandomState(5)
    perms = np.stack([rs.permutation(40) for _ in range(E)])
    w0 = kl.init_weights(D, U, L, seed=6, dtype=np.float64)
    w = [a.copy() for a in w0]
    l2 = [1e-4, 0, 1e-4, 1e-4, 0, 1e-4, 0, 0]   # cells: input kernel and bias (the plugin's regularisers)
    hist, adam = kl.fit(w, "elu", X, Y, E, B, perms, mv, l2=l2, dtype=np.float64)
    # independent loop: autograd + textbook Adam
    wt = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in w0]
    m = [torch.zeros_like(a) for a in wt]; v = [torch.zeros_like(a) for a in wt]
    Xt, Yt = torch.tensor(X), torch.tensor(Y[..., 0])
    mask = torch.tensor(kl.compute_mask(X, mv))
    t, hist_t = 0, []
    for e in range(E):
        tot = 0.0
        for s in range(0, 40, B):
            idx = torch.as_tensor(perms[e][s:s + B])
            u = _torch_forward(wt, "elu", Xt[idx], mask[idx])
            per = torch.nn.functional.binary_cross_entropy_with_logits(u, Yt[idx], reduction="none")
            loss = (per * mask[idx].double()).sum() / u.numel()
            loss = loss + sum(c * (a * a).sum() for c, a in zip(l2, wt) if c)
            gr = torch.autograd.grad(loss, wt)
            tot += float(loss) * len(idx)
            t += 1
            lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
            with torch.no_grad():
                for i, g in enumerate(gr):
                    m[i] = 0.9 * m[i] + 0.1 * g
                    v[i] = 0.999 * v[i] + 0.001 * g * g
                    wt[i] -= lr_t * m[i] / (torch.sqrt(v[i]) + 1e-7)
        hist_t.append(tot / 40)
    assert adam.t == t == 120
    assert np.abs(hist - np.array(hist_t)).max() <= 1e-10
    assert max(np.abs(a - b.detach().numpy()).max() for a, b in zip(w, wt)) <= 1e-10
    assert hist[-1] < hist[0]
