"""The reference-facing Python surface on the GPU: ports of the reference's own hot-path test
(tests/test_models.py:12-50) and of the README loop, plus argmax parity with the oracle's
restatement of bore/mixins.py."""
import numpy as np
import pytest
from scipy.optimize import Bounds, minimize

from oracle import keras_mlp as km, argmax as am
from helpers import NETS, trained_weights, synthetic_targets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 42, 8888])
@pytest.mark.parametrize("activation", [None, "relu", "elu"])
def test_maximizable_dense_sequential(seed, activation):
    """Reference tests/test_models.py:12-50 (activation=None is the reference's exact case:
    default-init, linear hidden layers); relu/elu widen it as SURVEY.md section 4 asks."""
    from bore_b200.models import MaximizableDenseSequential
    random_state = np.random.RandomState(seed)
    input_dim, output_dim, n_layers, n_units = 2, 1, 2, 32
    n_starts, n_samples = 5, 1024
    bounds = Bounds(lb=np.zeros(input_dim), ub=np.ones(input_dim))
    layer_kws = {} if activation is None else dict(activation=activation)
    model = MaximizableDenseSequential(input_dim=input_dim, output_dim=output_dim,
                                       num_layers=n_layers, num_units=n_units,
                                       layer_kws=layer_kws, seed=seed)
    assert len(model.layers) == 4  # the layer-count quirk: 3 hidden + output
    X_test = random_state.uniform(low=bounds.lb, high=bounds.ub, size=(n_samples, input_dim))
    y_test = model.predict(X_test)
    assert y_test.shape == (n_samples, output_dim)
    opt = model.argmax(bounds=bounds, num_starts=n_starts, num_samples=n_samples,
                       method="L-BFGS-B", options=dict(maxiter=1000, ftol=1e-9),
                       print_fn=lambda x: None, random_state=random_state)
    assert opt.x.shape == (input_dim,)
    X_opt = np.expand_dims(opt.x, axis=0)
    if activation is None:
        assert np.greater_equal(model.predict(X_opt), y_test).all()
    else:  # nonlinear nets: a local optimiser need not beat every random sample; it does beat its start
        assert model.predict(X_opt)[0, 0] >= np.sort(y_test[:, 0])[-n_starts] - 1e-6


def test_readme_loop():
    """README.rst:56-103 with a synthetic black box."""
    from bore_b200.models import MaximizableSequential
    from bore_b200.layers import Dense
    rs = np.random.RandomState(0)
    classifier = MaximizableSequential(seed=0)
    classifier.add(Dense(16, activation="relu"))
    classifier.add(Dense(16, activation="relu"))
    classifier.add(Dense(1, activation="sigmoid"))
    classifier.compile(optimizer="adam", loss="binary_crossentropy")
    bounds = [(0.0, 1.0), (0.0, 1.0)]
    features = list(rs.uniform(size=(10, 2)))
    targets = list(synthetic_targets(np.vstack(features)))
    for i in range(4):
        X = np.vstack(features)
        y = np.hstack(targets)
        z = np.less(y, np.quantile(y, q=0.25))
        hist = classifier.fit(X, z, epochs=50, batch_size=64, verbose=0)
        assert len(hist.history["loss"]) == 50
        res = classifier.argmax(method="L-BFGS-B", num_start_points=3, bounds=bounds,
                                print_fn=None, random_state=rs)
        assert res is not None and res.x.shape == (2,)
        assert np.all(res.x >= 0) and np.all(res.x <= 1)
        features.append(res.x)
        targets.append(synthetic_targets(res.x[None])[0])
    # Adam state persisted across the fits: 4 * 50 epochs * 1 step
    assert classifier.get_optimizer_state()[2] == 200


@pytest.mark.parametrize("name", ["cfg5_plugin8", "cfg2_hartmann6"])
def test_argmax_matches_oracle_restatement(name):
    """Same X_init (same RandomState), same top-k set, same winner as the oracle's literal
    restatement of maxima/argmax driving SciPy."""
    from bore_b200 import ops
    from bore_b200.models import MaximizableSequential
    from bore_b200.layers import Dense
    dims, acts, transform = NETS[name]
    w = trained_weights(dims, acts, seed=21)
    model = MaximizableSequential(transform=ops.TRANSFORMS[transform])
    for i, (u, a) in enumerate(zip(dims[1:], acts)):
        model.add(Dense(u, activation=a, input_dim=dims[0] if i == 0 else None))
    model.set_weights(w)
    n = dims[0]
    bounds = Bounds(np.zeros(n), np.ones(n))
    lines = []
    got = model.maxima(bounds, num_starts=8, num_samples=512, print_fn=lines.append,
                       random_state=np.random.RandomState(5))
    ref = am.maxima(w, acts, bounds, num_starts=8, num_samples=512, print_fn=lambda s: None,
                    random_state=np.random.RandomState(5), transform=transform)
    assert len(got) == len(ref) == 8 and len(lines) == 8
    assert lines[0].startswith("[Maximum 01: value=")
    # same SET of optima (argpartition order is unspecified in the reference)
    gf = np.sort([r.fun for r in got]); rf = np.sort([r.fun for r in ref])
    agree = np.abs(gf - rf) <= 1e-4
    assert agree.mean() >= (1.0 if "relu" not in acts else 0.75), (gf, rf)
    best = model.argmax(bounds, num_starts=8, num_samples=512, print_fn=None,
                        random_state=np.random.RandomState(5))
    best_ref = am.argmax(w, acts, bounds, num_starts=8, num_samples=512, print_fn=lambda s: None,
                         random_state=np.random.RandomState(5), transform=transform)
    assert abs(best.fun - best_ref.fun) <= 1e-4
    # filter_fn path: reject the winner -> next best qualifying result
    second = model.argmax(bounds, filter_fn=lambda r: abs(r.fun - best.fun) > 1e-7, num_starts=8,
                          num_samples=512, print_fn=None, random_state=np.random.RandomState(5))
    assert second is None or second.fun >= best.fun
    none = model.argmax(bounds, filter_fn=lambda r: False, num_starts=8, num_samples=512,
                        print_fn=None, random_state=np.random.RandomState(5))
    assert none is None
    zero = model.argmax(bounds, num_starts=0, num_samples=512, print_fn=None,
                        random_state=np.random.RandomState(5))
    X_init = np.random.RandomState(5).uniform(size=(512, n))
    assert np.array_equal(zero.x, X_init[np.argmin(-km.predict(w, acts, X_init)[:, 0])]) or \
        abs(zero.fun - (-km.predict(w, acts, X_init)[:, 0]).min()) <= 1e-6


def test_convert_drives_scipy_and_multi_start():
    """convert() returns [f, g] with the reference's shapes and dtypes; SciPy's own L-BFGS-B can
    drive it (K2-in-the-loop harness); minimize_multi_start returns a list of results."""
    from bore_b200 import convert, ops
    from bore_b200.models import MaximizableSequential
    from bore_b200.layers import Dense
    from bore_b200.optimizers import minimize_multi_start
    dims, acts, transform = NETS["cfg5_plugin8"]
    w = trained_weights(dims, acts, seed=2)
    model = MaximizableSequential(transform=ops.sigmoid)
    for i, (u, a) in enumerate(zip(dims[1:], acts)):
        model.add(Dense(u, activation=a, input_dim=dims[0] if i == 0 else None))
    model.set_weights(w)
    x = np.random.RandomState(0).uniform(size=8)
    out = model._func_min(x)
    assert isinstance(out, list) and len(out) == 2
    f, g = out
    assert f.shape == () and f.dtype == np.float32 and g.shape == (8,) and g.dtype == np.float64
    f_ref, g_ref = am.make_func_min(w, acts, transform)(x)
    assert abs(f - f_ref) <= 1e-6 and np.abs(g - g_ref).max() <= 1e-6
    fb, gb = model._func_min(np.stack([x, x * 0.5]))
    assert fb.shape == (2,) and gb.shape == (2, 8)
    # SciPy on the host driving GPU evaluations
    bounds = Bounds(np.zeros(8), np.ones(8))
    r = minimize(model._func_min, x0=x, method="L-BFGS-B", jac=True, bounds=bounds,
                 options=dict(maxiter=1000, ftol=1e-9))
    r_ref = minimize(am.make_func_min(w, acts, transform), x0=x, method="L-BFGS-B", jac=True,
                     bounds=bounds, options=dict(maxiter=1000, ftol=1e-9))
    assert abs(r.fun - r_ref.fun) <= 1e-5
    # the same start on the device
    res = minimize_multi_start(model._func_min, bounds, num_starts=4, num_samples=64,
                               random_state=np.random.RandomState(3), method="L-BFGS-B", jac=True,
                               options=dict(maxiter=1000, ftol=1e-9))
    assert len(res) == 4 and all(hasattr(q, "x") and q.x.shape == (8,) for q in res)
    # a plain Python objective goes through the reverse-communication stepper
    def quad(xx):
        xx = np.asarray(xx)
        if xx.ndim == 2:
            return np.sum((xx - 0.3) ** 2, axis=1), 2 * (xx - 0.3)
        return np.sum((xx - 0.3) ** 2), 2 * (xx - 0.3)
    res = minimize_multi_start(quad, [(0.0, 1.0)] * 3, num_starts=3, num_samples=16,
                               random_state=np.random.RandomState(0))
    for q in res:
        assert q.success and np.abs(q.x - 0.3).max() <= 1e-5
    with pytest.raises(NotImplementedError):
        model.argmax(bounds, method="BFGS", print_fn=None)


def test_allreduce_maxloc_through_the_c_abi_with_a_raw_nccl_communicator():
    """bore_allreduce_maxloc: the collective is reachable without torch -- a
    communicator made with NCCL's own ncclCommInitAll (one rank here; tools/nccl_maxloc_2gpu.py
    runs two), the packed key reduced in place."""
    import ctypes as C
    import torch
    from bore_b200 import _lib
    lib = _lib.require_cuda()
    try:
        nccl = C.CDLL("libnccl.so.2")
    except OSError:
        pytest.skip("no libnccl.so.2 on the loader path")
    comm = C.c_void_p()
    devs = (C.c_int * 1)(0)
    assert nccl.ncclCommInitAll(C.byref(comm), 1, devs) == 0
    try:
        key = torch.tensor([0x3F80000012345678], dtype=torch.int64, device="cuda:0")
        stream = torch.cuda.current_stream(0).cuda_stream
        _lib.check(lib.bore_allreduce_maxloc(comm, C.c_void_p(key.data_ptr()), C.c_void_p(stream)))
        torch.cuda.synchronize()
        assert int(key.item()) == 0x3F80000012345678
    finally:
        nccl.ncclCommDestroy(comm)
    assert lib.bore_allreduce_maxloc(None, None, None) != 0      # loud on bad arguments
