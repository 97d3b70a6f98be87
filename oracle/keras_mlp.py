"""NumPy restatement of the Keras arithmetic on the BORE-MLP hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** for this file:
TensorFlow 2.5.0 (reference pin, /root/reference/setup.py:42) is absent from the image,
so every rule below is written from knowledge of that release and flagged
[TF-semantics].  What narrows the risk to the RULES themselves (tests/test_oracle.py):
value and gradients agree with torch-CPU autograd, and the whole ``fit`` loop agrees with an
independent torch implementation (autograd gradients, torch's own BCE-with-logits, the Adam
update written in the textbook beta*m + (1-beta)*g form) to 1e-10 over hundreds of fp64 steps.

The TF 2.5 sources each [TF-semantics] rule restates (paths inside the tensorflow repository at
tag v2.5.0; cited from knowledge of that tree -- it cannot be opened here):

* Dense                 tensorflow/python/keras/layers/core.py ``Dense.call`` ->
                        keras/layers/ops/core.py ``dense``: MatMul(inputs, kernel), bias_add,
                        activation; kernel shape (in, out)
* glorot_uniform        keras/initializers/initializers_v2.py ``GlorotUniform`` =
                        VarianceScaling(scale=1, mode="fan_avg", distribution="uniform"):
                        limit = sqrt(3 * scale / ((fan_in + fan_out) / 2))
* relu / elu grads      tensorflow/core/kernels/relu_op_functor.h ``ReluGrad`` (gradients *
                        (features > 0)) and ``EluGrad`` ((activations < 0).select((activations + 1)
                        * gradients, gradients)) -- both through the layer OUTPUT;
                        sigmoid / tanh: core/kernels/cwise_ops_gradients.h (y (1 - y) dy, (1 - y^2) dy)
* "binary_crossentropy" keras/backend.py ``binary_crossentropy``: for a graph tensor produced by a
  on a sigmoid output   ``Sigmoid`` op the logits are taken from ``output.op.inputs[0]`` and
                        from_logits is set ("we use logits from the sigmoid function directly");
                        otherwise clip to [eps, 1 - eps] and the log form.  Then
                        python/ops/nn_impl.py ``sigmoid_cross_entropy_with_logits``:
                        relu(x) - x z + log1p(exp(-|x|))
* loss reduction        keras/losses.py ``LossFunctionWrapper`` with
                        ReductionV2.SUM_OVER_BATCH_SIZE; regularisers added by
                        keras/engine/compile_utils.py ``LossesContainer.__call__``
* Adam                  keras/optimizer_v2/adam.py ``Adam._resource_apply_dense`` ->
                        ResourceApplyAdam, core/kernels/training_ops.cc ``ApplyAdam``:
                        alpha = lr sqrt(1 - beta2^t) / (1 - beta1^t); m += (g - m)(1 - beta1);
                        v += (g^2 - v)(1 - beta2); var -= (m alpha) / (sqrt(v) + epsilon);
                        epsilon = backend_config.epsilon() = 1e-7; t = iterations + 1
* fit / batches         keras/engine/training.py ``Model.fit`` with
                        keras/engine/data_adapter.py ``TensorLikeDataAdapter``: per epoch
                        ``random_ops.random_shuffle(range(N))``, ``slice_batch_indices``: full
                        batches + one partial batch; epoch loss = keras/metrics.py ``Mean`` with
                        sample_weight = batch size (compile_utils.py ``LossesContainer``)
* accuracy on logits    keras/metrics.py ``binary_accuracy(y_true, y_pred, threshold=0.5)`` applied
                        to the raw output (plugin quirk, logging only)

Reference call sites each function follows:

* ``dense_sequential_dims``     -> bore/models.py:9-21  (layer-count quirk)
* ``forward`` / ``predict``     -> Keras ``Sequential.__call__`` / ``predict`` as used at
                                   bore/mixins.py:50, bore/base.py:40
* ``value_and_input_grad``      -> bore/base.py:35-42 + bore/decorators.py:24-79 with the
                                   ``transform(-u)`` closure of bore/mixins.py:20
* ``fit`` / ``evaluate``        -> Keras ``fit``/``evaluate`` as called at README.rst:93 and
                                   bore/plugins/hpbandster/base.py:184-186, compiled with
                                   adam + binary cross-entropy (README.rst:66;
                                   plugins/hpbandster/base.py:156-157)

All arithmetic is fp32 by default (``dtype=np.float32``), matching Keras' autocast of the
fp64 inputs SciPy hands over.  ``dtype=np.float64`` gives the
"truth" variant used to bound fp32 rounding noise in the tests.
"""
import numpy as np

ACTIVATIONS = ("linear", "relu", "elu", "sigmoid", "tanh")
TRANSFORMS = ("identity", "sigmoid", "exp")  # plugins/hpbandster/base.py:18


# --------------------------------------------------------------------------- model spec
def dense_sequential_dims(input_dim, output_dim, num_layers, num_units):
    """Layer widths built by ``DenseSequential.__init__`` (bore/models.py:14-21).

    The ``if not i`` branch has no ``else``, so the first hidden layer is added twice:
    ``num_layers=2`` yields THREE hidden layers.  Reproduced on purpose.
    """
    dims = [input_dim]
    for i in range(num_layers):
        if not i:
            dims.append(num_units)
        dims.append(num_units)
    dims.append(output_dim)
    return dims


def glorot_uniform(rs, fan_in, fan_out, dtype=np.float32):
    """Keras default ``kernel_initializer`` [TF-semantics]: U(+-sqrt(6/(fan_in+fan_out)))."""
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rs.uniform(-limit, limit, size=(fan_in, fan_out)).astype(dtype)


def init_weights(dims, seed, dtype=np.float32):
    """Keras-ordered weight list ``[W0 (in,out), b0 (out,), W1, b1, ...]`` with zero biases."""
    rs = np.random.RandomState(seed)
    ws = []
    for fi, fo in zip(dims[:-1], dims[1:]):
        ws.append(glorot_uniform(rs, fi, fo, dtype))
        ws.append(np.zeros(fo, dtype))
    return ws


# --------------------------------------------------------------------------- activations
def _act(name, a):
    if name == "linear":
        return a
    if name == "relu":
        return np.maximum(a, a.dtype.type(0))
    if name == "elu":  # alpha = 1  [TF-semantics]
        return np.where(a > 0, a, np.expm1(np.minimum(a, a.dtype.type(0))))
    if name == "sigmoid":
        return _sigmoid(a)
    if name == "tanh":
        return np.tanh(a)
    raise ValueError(name)


def _act_grad_from_output(name, h):
    """d act / d pre-activation, expressed through the layer OUTPUT h (what TF's
    ReluGrad / EluGrad / SigmoidGrad / TanhGrad kernels take) [TF-semantics]."""
    one = h.dtype.type(1)
    if name == "linear":
        return np.ones_like(h)
    if name == "relu":
        return (h > 0).astype(h.dtype)
    if name == "elu":
        return np.where(h > 0, one, h + one)
    if name == "sigmoid":
        return h * (one - h)
    if name == "tanh":
        return one - h * h
    raise ValueError(name)


def _sigmoid(a):
    # stable in both tails; Eigen's logistic is 1/(1+exp(-x)) with clamping
    out = np.empty_like(a)
    pos = a >= 0
    out[pos] = 1 / (1 + np.exp(-a[pos]))
    e = np.exp(a[~pos])
    out[~pos] = e / (1 + e)
    return out


# --------------------------------------------------------------------------- forward
def forward(weights, acts, X, dtype=np.float32, keep=False):
    """``y = act(x @ W + b)`` per Dense layer [TF-semantics].  X: (S, D)."""
    h = np.asarray(X).astype(dtype)  # Keras autocast f64 -> f32 at model entry
    hs = [h]
    for l, act in enumerate(acts):
        W = weights[2 * l].astype(dtype, copy=False)
        b = weights[2 * l + 1].astype(dtype, copy=False)
        h = _act(act, h @ W + b)
        hs.append(h)
    return (h, hs) if keep else h


def predict(weights, acts, X, dtype=np.float32):
    """Keras ``predict`` (bore/mixins.py:50): returns (S, out_dim)."""
    return forward(weights, acts, X, dtype)


def _transform(name, v):
    if name == "identity":
        return v, np.ones_like(v)
    if name == "sigmoid":
        s = _sigmoid(v)
        return s, s * (1 - s)
    if name == "exp":
        e = np.exp(v)
        return e, e
    raise ValueError(name)


def value_and_input_grad(weights, acts, X, transform="identity", negate=True,
                         dtype=np.float32):
    """f = T(-u(x)) (``negate=True``, the ``_func_min`` of bore/mixins.py:20) or T(u(x))
    (``_func_max``, bore/mixins.py:97) and g = df/dx, for every row of X.

    Mirrors bore/decorators.py:48-65: the tape watches only x.  Returns (f (S,), g (S, D))
    in ``dtype`` -- the reference then hands them to SciPy as fp64-typed values
    (decorators.py:54-56), which the callers of this function do themselves.
    """
    X = np.atleast_2d(np.asarray(X))
    u, hs = forward(weights, acts, X, dtype, keep=True)
    assert u.shape[1] == 1, "output dimension must be 1 (bore/base.py:19-21)"
    sgn = dtype(-1) if negate else dtype(1)
    f, dT = _transform(transform, sgn * u)
    delta = dT * sgn  # d f / d u
    for l in range(len(acts) - 1, -1, -1):
        delta = delta * _act_grad_from_output(acts[l], hs[l + 1])
        delta = delta @ weights[2 * l].astype(dtype, copy=False).T
    return f[:, 0], delta


# --------------------------------------------------------------------------- training
def bce_with_logits(u, z):
    """``tf.nn.sigmoid_cross_entropy_with_logits`` [TF-semantics]:
    max(u,0) - u*z + log1p(exp(-|u|)).  Keras uses this form both for
    ``BinaryCrossentropy(from_logits=True)`` (plugins/hpbandster/base.py:157) and, inside
    the compiled train function, for ``"binary_crossentropy"`` on a Sigmoid output
    (README.rst:64-66), where it reaches through to the Sigmoid op's input."""
    return np.maximum(u, 0) - u * z + np.log1p(np.exp(-np.abs(u)))


class AdamState:
    """Keras Adam [TF-semantics]: lr 1e-3, beta1 .9, beta2 .999, eps 1e-7 outside the
    bias correction; ``iterations`` persists across ``fit`` calls on the same model."""

    def __init__(self, weights, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.m = [np.zeros_like(w) for w in weights]
        self.v = [np.zeros_like(w) for w in weights]
        self.t = 0
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps


def adam_apply(weights, grads, st, dtype=np.float32):
    """ResourceApplyAdam [TF-semantics]:
    alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    w -= alpha*m/(sqrt(v)+eps)."""
    f = dtype
    st.t += 1
    b1, b2 = f(st.beta1), f(st.beta2)
    b1p = f(np.power(b1, f(st.t)))
    b2p = f(np.power(b2, f(st.t)))
    alpha = f(f(st.lr) * np.sqrt(f(1) - b2p) / (f(1) - b1p))
    for i, g in enumerate(grads):
        st.m[i] += (g - st.m[i]) * (f(1) - b1)
        st.v[i] += (g * g - st.v[i]) * (f(1) - b2)
        weights[i] -= (st.m[i] * alpha) / (np.sqrt(st.v[i]) + f(st.eps))


def loss_and_weight_grads(weights, acts, Xb, zb, l2=0.0, dtype=np.float32):
    """Mean BCE-with-logits over the batch (+ l2*sum(w^2) over kernels AND biases,
    plugins/hpbandster/base.py:113-116) and its gradient wrt every weight.

    ``acts[-1]`` may be "sigmoid" (README form) or "linear" (plugin form); either way the
    loss is taken on the final pre-activation (see ``bce_with_logits``)."""
    f = dtype
    final = acts[-1]
    assert final in ("sigmoid", "linear")
    acts_l = list(acts[:-1]) + ["linear"]
    u, hs = forward(weights, acts_l, Xb, dtype, keep=True)
    n = Xb.shape[0]
    z = zb.astype(dtype).reshape(n, 1)
    loss = f(np.mean(bce_with_logits(u, z), dtype=dtype))
    delta = (_sigmoid(u) - z) / f(n)
    grads = [None] * len(weights)
    for l in range(len(acts) - 1, -1, -1):
        if l < len(acts) - 1:
            delta = delta * _act_grad_from_output(acts_l[l], hs[l + 1])
        grads[2 * l] = hs[l].T @ delta
        grads[2 * l + 1] = delta.sum(axis=0)
        if l > 0:
            delta = delta @ weights[2 * l].T
    l2v = _l2_vector(l2, len(weights))
    if any(l2v):
        reg = f(0)
        for i, w in enumerate(weights):
            if l2v[i]:
                reg += f(l2v[i]) * f(np.sum(w * w, dtype=dtype))
                grads[i] = grads[i] + f(2 * l2v[i]) * w
        loss = f(loss + reg)
    return loss, grads


def _l2_vector(l2, n):
    """``l2`` is one factor for every kernel and bias, or a per-array sequence in Keras weight
    order [k0, b0, k1, b1, ...] (the plugin regularises the hidden layers only:
    plugins/hpbandster/base.py:147-155 passes the regularisers through ``layer_kws``, not
    ``final_layer_kws``)."""
    if np.isscalar(l2):
        return [float(l2)] * n
    l2 = [float(v) for v in l2]
    assert len(l2) == n
    return l2


def fit(weights, acts, X, z, epochs, batch_size, permutations, adam=None, l2=0.0,
        dtype=np.float32):
    """Keras ``fit(shuffle=True)`` [TF-semantics] with the per-epoch permutations made
    explicit (a build extension needed for parity; Keras draws them internally).

    Per epoch: samples ``permutations[e]`` are cut into consecutive batches of
    ``batch_size``; the last short batch is kept (bore/math.py:8-29) and averaged over its
    own size.  Reported epoch loss = sum(batch_loss * batch_n) / N, each batch_loss taken
    BEFORE that step's update.  ``weights`` and ``adam`` are updated in place.

    Returns (history_loss (epochs,), adam_state).
    """
    X = np.asarray(X).astype(dtype)
    z = np.asarray(z).astype(dtype)
    N = X.shape[0]
    if adam is None:
        adam = AdamState(weights)
    hist = np.zeros(epochs, dtype)
    for e in range(epochs):
        perm = np.asarray(permutations[e])
        tot = dtype(0)
        for s in range(0, N, batch_size):
            idx = perm[s:s + batch_size]
            loss, grads = loss_and_weight_grads(weights, acts, X[idx], z[idx], l2, dtype)
            tot += loss * dtype(len(idx))
            adam_apply(weights, grads, adam, dtype)
        hist[e] = tot / dtype(N)
    return hist, adam


def evaluate(weights, acts, X, z, l2=0.0, from_logits=True, dtype=np.float32):
    """Keras ``evaluate`` -> [loss, accuracy] (plugins/hpbandster/base.py:186).
    Quirk kept [TF-semantics]: ``accuracy`` thresholds the model OUTPUT at 0.5, so with a
    linear (logit) output it thresholds the logit, not the probability."""
    X = np.asarray(X).astype(dtype)
    zz = np.asarray(z).astype(dtype)
    loss, _ = loss_and_weight_grads(weights, acts, X, zz, l2, dtype)
    out = forward(weights, acts, X, dtype)[:, 0]
    acc = np.mean((out > 0.5).astype(dtype) == zz)
    return [float(loss), float(acc)]
