"""The two argmax paths -- the fused persistent kernel (lbfgsb_fused.cu) and the lock-step rounds
of K2 + stepper launches (lbfgsb.cu) -- run the same L-BFGS-B core (lbfgsb_core.h) on MLP values
and gradients that are summed in the same order, so their results must be EQUAL, bit for bit: x,
fun, nit, nfev, status, task of every start.  Parity with SciPy is then tested once
(test_gpu_lbfgsb.py, test_gpu_fullsize.py) and holds for both."""
import numpy as np
import pytest

from helpers import NETS, trained_weights

pytestmark = pytest.mark.gpu

KEYS = ("x", "fun", "nit", "nfev", "status", "task")


def _set_mode(mode):
    from bore_b200 import _lib
    lib = _lib.require_cuda()
    _lib.check(lib.bore_lbfgsb_set_mode(mode))


@pytest.fixture(autouse=True)
def _restore_mode():
    yield
    _set_mode(0)


def _run(net, X0, transform, mode, **kw):
    import torch
    _set_mode(mode)
    net._work = None
    r = net.lbfgsb_dev(torch.from_numpy(X0).cuda(), 0.0, 1.0, transform=transform, **kw)
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in r.items()}


@pytest.mark.parametrize("name", sorted(NETS))
def test_fused_equals_rounds(name):
    from bore_b200.engine import NativeMLP
    dims, acts, transform = NETS[name]
    net = NativeMLP(dims, acts)
    net.set_weights(trained_weights(dims, acts, seed=5))
    X0 = np.random.RandomState(7).uniform(size=(700, dims[0]))
    a = _run(net, X0, transform, 2)
    b = _run(net, X0, transform, 1)
    assert a["rounds"] == 1 and b["rounds"] > 1      # one launch vs many
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["evals"] == b["evals"]


@pytest.mark.parametrize("m", [3, 7])
def test_fused_equals_rounds_other_maxcor(m):
    from bore_b200.engine import NativeMLP
    dims, acts, transform = NETS["cfg5_plugin8"]
    net = NativeMLP(dims, acts)
    net.set_weights(trained_weights(dims, acts, seed=2))
    X0 = np.random.RandomState(3).uniform(size=(300, dims[0]))
    a = _run(net, X0, transform, 2, m=m)
    b = _run(net, X0, transform, 1, m=m)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_default_rule_hands_the_tail_over(monkeypatch):
    """Above 16,384 starts (4,096 for more than 16 dimensions) the default rule runs lock-step rounds and finishes the last few
    thousand starts with the fused kernel in resume mode: same results as rounds alone."""
    from bore_b200.engine import NativeMLP
    dims, acts, transform = NETS["cfg5_plugin8"]
    net = NativeMLP(dims, acts)
    net.set_weights(trained_weights(dims, acts, seed=4))
    X0 = np.random.RandomState(11).uniform(size=(20000, dims[0]))
    a = _run(net, X0, transform, 0)
    b = _run(net, X0, transform, 1)
    assert 1 < a["rounds"] < b["rounds"]              # the tail was taken over
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["evals"] == b["evals"]


def test_batched_problems_fused_equals_rounds():
    """One CTA per model (BASELINE.json configs[3]): a few starts for each of many models."""
    import torch
    from bore_b200.engine import NativeMLP
    from oracle import keras_mlp as km
    dims, acts = [6, 32, 32, 1], ["relu", "relu", "sigmoid"]
    M, K = 37, 5
    net = NativeMLP(dims, acts, n_models=M)
    for p in range(M):
        net.set_weights(trained_weights(dims, acts, seed=100 + p, N=120, epochs=6), model=p)
    X0 = np.random.RandomState(5).uniform(size=(M, K, 6))
    out = []
    for mode in (2, 1):
        _set_mode(mode)
        net._work = None
        r = net.lbfgsb_multi_dev(torch.from_numpy(X0).cuda(), 0.0, 1.0)
        out.append({k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in r.items()})
    for k in KEYS:
        assert np.array_equal(out[0][k], out[1][k]), k
