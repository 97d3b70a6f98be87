"""NumPy restatement of the Keras arithmetic of the reference's LSTM multi-fidelity classifier.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**, like oracle/keras_mlp.py:
TensorFlow 2.5.0 cannot run here; every rule is written from knowledge of that release and flagged
[TF-semantics]; forward values, input gradients, weight gradients and the whole fit loop are
cross-checked against an independent torch-CPU autograd implementation (tests/test_oracle_lstm.py).

Reference call sites:

* ``StackedRecurrentFactory`` (bore/models.py:48-104): ``num_layers`` ``LSTMCell(num_units, **layer_kws)``
  + ``Dense(output_dim)``; ``build_many_to_many`` = Masking -> RNN(cell, return_sequences=True) per
  cell -> TimeDistributed(Dense); ``build_one_to_one(num_steps)`` = RepeatVector(num_steps) -> the
  same cells (last one return_sequences=False) -> the same Dense, as a ``MaximizableSequential``.
* training: ``SequenceClassifierConfigGenerator._update_classifier``
  (bore/plugins/hpbandster/multi_fidelity.py:198-233): ``fit(inputs, targets, epochs, batch_size)``
  with adam + ``BinaryCrossentropy(from_logits=True)`` on sequences padded with ``mask_value``
  (bore/data.py:183-251).

[TF-semantics] restated (tensorflow v2.5.0 source paths, from knowledge of that tree):

* LSTMCell  keras/layers/recurrent_v2.py / recurrent.py ``LSTMCell.call``: z = x K + h R + b with
            K (in, 4U), R (U, 4U), b (4U), gate order i, f, c, o; i, f, o = sigmoid
            (``recurrent_activation``), c~ = activation(z_c); c = f c_prev + i c~; h = o activation(c);
            ``unit_forget_bias=True``: bias initialised to 0 with the f slice = 1; kernel
            glorot_uniform, recurrent kernel orthogonal.
* Masking   keras/layers/core.py ``Masking``: step t of sample b is masked when ALL its features equal
            mask_value; keras/backend.py ``rnn`` (mask branch, zero_output_for_mask=False): at a
            masked step the states are carried over and the output is the previous output (zeros
            before the first step); the mask travels through return_sequences=True layers and
            TimeDistributed to the loss.
* loss      keras/engine/compile_utils.py ``LossesContainer.__call__``: the mask becomes the sample
            weight; losses.py BinaryCrossentropy (mean over the last axis, here of size 1) with
            ReductionV2.SUM_OVER_BATCH_SIZE divides the weighted sum by the NUMBER OF ELEMENTS
            (batch x steps), masked ones included.
* Adam, fit loop, regularisers: as oracle/keras_mlp.py (the plugin passes kernel_regularizer and
  bias_regularizer to the cells: the input kernel and the bias, not the recurrent kernel).
"""
import numpy as np

from . import keras_mlp as km


def _act(name, a):
    return km._act("linear" if name is None else name, a)


def _act_grad_from_output(name, h):
    return km._act_grad_from_output("linear" if name is None else name, h)


def init_weights(input_dim, units, num_layers, seed, output_dim=1, dtype=np.float32):
    """Keras-ordered weights ``[K_0, R_0, b_0, K_1, R_1, b_1, ..., W_dense, b_dense]`` with the
    default initialisers (glorot_uniform / orthogonal / zeros + unit forget bias)."""
    rs = np.random.RandomState(seed)
    ws = []
    fan_in = input_dim
    for _ in range(num_layers):
        ws.append(km.glorot_uniform(rs, fan_in, 4 * units, dtype))
        q, r = np.linalg.qr(rs.normal(size=(4 * units, units)))
        q = q * np.sign(np.diag(r))
        ws.append(np.ascontiguousarray(q.T).astype(dtype))          # (units, 4 units), orthonormal rows
        b = np.zeros(4 * units, dtype)
        b[units:2 * units] = 1
        ws.append(b)
        fan_in = units
    ws.append(km.glorot_uniform(rs, units, output_dim, dtype))
    ws.append(np.zeros(output_dim, dtype))
    return ws


def compute_mask(X, mask_value):
    """(B, T) bool: True where the step is NOT masked (Masking: any(x != mask_value, axis=-1))."""
    return np.any(np.asarray(X) != mask_value, axis=-1)


def forward(weights, activation, X, mask=None, dtype=np.float32, keep=False):
    """Stacked LSTM + Dense over sequences X (B, T, D) -> logits (B, T) (output_dim 1).
    ``mask`` (B, T) bool or None.  With ``keep``: also the per-layer, per-step activations."""
    f = dtype
    X = np.asarray(X).astype(f)
    B, T, _ = X.shape
    L = (len(weights) - 2) // 3
    U = weights[1].shape[0]
    if mask is None:
        mask = np.ones((B, T), bool)
    seq = X
    cache = []
    for l in range(L):
        K, R, b = (weights[3 * l + i].astype(f, copy=False) for i in range(3))
        h = np.zeros((B, U), f)
        c = np.zeros((B, U), f)
        out = np.zeros((B, T, U), f)
        lay = []
        for t in range(T):
            z = seq[:, t] @ K + h @ R + b
            i = km._sigmoid(z[:, :U]); fg = km._sigmoid(z[:, U:2 * U])
            g = _act(activation, z[:, 2 * U:3 * U]); o = km._sigmoid(z[:, 3 * U:])
            c_new = fg * c + i * g
            ac = _act(activation, c_new)
            h_new = o * ac
            m = mask[:, t][:, None]
            lay.append(dict(x=seq[:, t].copy(), h_prev=h, c_prev=c, i=i, f=fg, g=g, o=o, c=c_new, ac=ac))
            h = np.where(m, h_new, h)        # masked step: states carried, output = previous output
            c = np.where(m, c_new, c)
            out[:, t] = h
        cache.append(lay)
        seq = out
    Wd, bd = weights[-2].astype(f, copy=False), weights[-1].astype(f, copy=False)
    u = (seq @ Wd + bd)[..., 0]
    if keep:
        return u, seq, cache
    return u


def predict_one_to_one(weights, activation, X, num_steps, dtype=np.float32):
    """``build_one_to_one(num_steps)``: RepeatVector -> LSTMs -> Dense of the LAST step: (S, 1)."""
    X = np.asarray(X)
    Xr = np.repeat(X[:, None, :], num_steps, axis=1)
    return forward(weights, activation, Xr, None, dtype)[:, -1:].copy()


def _backward(weights, activation, cache, mask, d_out, dtype):
    """BPTT.  d_out (B, T, U): dLoss/d(top layer output) per step.  Returns the weight gradients
    of the recurrent layers (Keras order) and dLoss/dX (B, T, D)."""
    f = dtype
    L = len(cache)
    B, T, U = d_out.shape
    grads = []
    d_seq = d_out
    for l in range(L - 1, -1, -1):
        K, R = weights[3 * l].astype(f, copy=False), weights[3 * l + 1].astype(f, copy=False)
        dK, dR, db = np.zeros_like(K), np.zeros_like(R), np.zeros(4 * U, f)
        d_in = np.zeros((B, T, K.shape[0]), f)
        dh_next = np.zeros((B, U), f)
        dc_next = np.zeros((B, U), f)
        for t in range(T - 1, -1, -1):
            s = cache[l][t]
            m = mask[:, t][:, None]
            dh = d_seq[:, t] + dh_next       # output_t = h_t, masked or not
            do = dh * s["ac"] * s["o"] * (1 - s["o"])
            dc = dh * s["o"] * _act_grad_from_output(activation, s["ac"]) + dc_next
            di = dc * s["g"] * s["i"] * (1 - s["i"])
            df = dc * s["c_prev"] * s["f"] * (1 - s["f"])
            dg = dc * s["i"] * _act_grad_from_output(activation, s["g"])
            dz = np.where(m, np.concatenate([di, df, dg, do], axis=1), f(0))  # a masked step computes nothing
            dK += s["x"].T @ dz
            dR += s["h_prev"].T @ dz
            db += dz.sum(axis=0)
            d_in[:, t] = dz @ K.T
            dh_next = np.where(m, dz @ R.T, dh)                   # masked: h_t = h_{t-1}
            dc_next = np.where(m, dc * s["f"], dc_next)           # masked: c_t = c_{t-1}
        grads = [dK, dR, db] + grads
        d_seq = d_in
    return grads, d_seq


def value_and_input_grad(weights, activation, X, num_steps, transform="identity", negate=True,
                         dtype=np.float32):
    """The ``convert`` closure (bore/base.py:35-42) on the one-to-one network: f = T(-u(x)) and
    df/dx for every row of X (S, D) -- x is repeated over the steps, so the gradient sums over them."""
    f = dtype
    X = np.asarray(X)
    S = X.shape[0]
    Xr = np.repeat(X[:, None, :], num_steps, axis=1)
    u_all, top, cache = forward(weights, activation, Xr, None, dtype, keep=True)
    u = u_all[:, -1]
    sgn = f(-1) if negate else f(1)
    v = sgn * u
    if transform == "sigmoid":
        val = km._sigmoid(v); dT = val * (1 - val)
    elif transform == "exp":
        val = np.exp(v); dT = val
    else:
        val = v; dT = np.ones_like(v)
    du = (dT * sgn).astype(f)
    U = weights[1].shape[0]
    d_out = np.zeros((S, num_steps, U), f)
    d_out[:, -1] = du[:, None] * weights[-2].astype(f, copy=False)[:, 0][None, :]
    _, dX = _backward(weights, activation, cache, np.ones((S, num_steps), bool), d_out, dtype)
    return val.astype(f), dX.sum(axis=1).astype(f)


def loss_and_weight_grads(weights, activation, Xb, Yb, mask, l2=None, dtype=np.float32):
    """Masked BCE-with-logits over a batch of sequences, divided by batch x steps, + l2 terms;
    gradient wrt every weight (Keras order)."""
    f = dtype
    B, T, _ = Xb.shape
    u, top, cache = forward(weights, activation, Xb, mask, dtype, keep=True)
    y = Yb.astype(f).reshape(B, T)
    w = mask.astype(f)
    n = f(B * T)
    loss = f(np.sum(km.bce_with_logits(u, y) * w, dtype=f) / n)
    du = (km._sigmoid(u) - y) * w / n
    Wd = weights[-2].astype(f, copy=False)
    d_out = du[:, :, None] * Wd[:, 0][None, None, :]
    dWd = (top.reshape(B * T, -1).T @ du.reshape(B * T, 1)).astype(f)
    dbd = np.array([du.sum()], f)
    grads, _ = _backward(weights, activation, cache, mask, d_out.astype(f), dtype)
    grads = grads + [dWd, dbd]
    if l2 is not None:
        l2v = km._l2_vector(l2, len(weights))
        reg = f(0)
        for i, wgt in enumerate(weights):
            if l2v[i]:
                reg += f(l2v[i]) * f(np.sum(wgt * wgt, dtype=f))
                grads[i] = grads[i] + f(2 * l2v[i]) * wgt
        loss = f(loss + reg)
    return loss, grads


def fit(weights, activation, X, Y, epochs, batch_size, permutations, mask_value, adam=None, l2=None,
        dtype=np.float32):
    """Keras ``fit`` on padded sequences X (N, T, D), Y (N, T, 1): per epoch the samples
    ``permutations[e]`` in consecutive batches, last short batch kept; epoch loss = batch losses
    weighted by batch size.  ``weights`` / ``adam`` updated in place."""
    X = np.asarray(X).astype(dtype)
    Y = np.asarray(Y).astype(dtype)
    mask = compute_mask(X, dtype(mask_value))
    N = X.shape[0]
    if adam is None:
        adam = km.AdamState(weights)
    hist = np.zeros(epochs, dtype)
    for e in range(epochs):
        perm = np.asarray(permutations[e])
        tot = dtype(0)
        for s in range(0, N, batch_size):
            idx = perm[s:s + batch_size]
            loss, grads = loss_and_weight_grads(weights, activation, X[idx], Y[idx], mask[idx], l2, dtype)
            tot += loss * dtype(len(idx))
            km.adam_apply(weights, grads, adam, dtype)
        hist[e] = tot / dtype(N)
    return hist, adam


def evaluate(weights, activation, X, Y, mask_value, l2=None, dtype=np.float32):
    """``evaluate`` -> [loss, accuracy]; the loss carries the l2 penalties (Keras adds ``model.losses``
    to the compiled loss in test_step too); the accuracy thresholds the LOGIT at 0.5 (from_logits
    quirk, logging only) and averages over the unmasked steps (metric sample weights = the mask)."""
    X = np.asarray(X).astype(dtype)
    mask = compute_mask(X, dtype(mask_value))
    u = forward(weights, activation, X, mask, dtype)
    y = np.asarray(Y).astype(dtype).reshape(u.shape)
    w = mask.astype(dtype)
    loss = float(np.sum(km.bce_with_logits(u, y) * w) / u.size)
    if l2 is not None:
        l2v = km._l2_vector(l2, len(weights))
        loss += float(sum(l2v[i] * np.sum(np.square(wgt, dtype=np.float64)) for i, wgt in enumerate(weights)))
    acc = float(np.sum(((u > 0.5).astype(dtype) == y) * w) / max(w.sum(), 1))
    return [loss, acc]
