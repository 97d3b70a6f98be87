"""K0/K2 parity: CUDA forward and value+input-gradient vs the NumPy oracle (fp32 and fp64)."""
import numpy as np
import pytest

from oracle import keras_mlp as km
from helpers import NETS, trained_weights

pytestmark = pytest.mark.gpu

# north_star: "MLP value and input-gradient must match TF within 1e-5 relative (fp32)".
# Relative to the magnitude of the quantity (|f| resp. max|g| of that point), with an
# absolute floor of a few fp32 ulps of the activations for near-zero values.
RTOL = 1e-5
ATOL_F = 2e-6
ATOL_G = 2e-6


def _engine(dims, acts, w):
    from bore_b200.engine import NativeMLP
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    return net


@pytest.mark.parametrize("name", list(NETS))
@pytest.mark.parametrize("S", [1, 3, 16, 17, 1024, 5000])
def test_value_and_grad_matches_oracle(name, S):
    dims, acts, transform = NETS[name]
    w = trained_weights(dims, acts, seed=1)
    rs = np.random.RandomState(S)
    X = rs.uniform(-0.2, 1.2, size=(S, dims[0]))
    net = _engine(dims, acts, w)
    for negate in (True, False):
        f, g = net.value_and_grad(X, transform, negate)
        f32, g32 = km.value_and_input_grad(w, acts, X, transform, negate, np.float32)
        f64, g64 = km.value_and_input_grad(w, acts, X, transform, negate, np.float64)
        gscale = np.abs(g64).max(axis=1, keepdims=True)
        assert np.all(np.abs(f - f32) <= RTOL * np.abs(f32) + ATOL_F), np.abs(f - f32).max()
        assert np.all(np.abs(g - g32) <= RTOL * gscale + ATOL_G), (np.abs(g - g32) / (gscale + 1e-30)).max()
        # and neither fp32 implementation is further from the fp64 truth than the other by much
        err_gpu = np.abs(g - g64).max()
        err_orc = np.abs(g32 - g64).max()
        assert err_gpu <= 4 * err_orc + ATOL_G


@pytest.mark.parametrize("name", list(NETS))
def test_predict_matches_oracle(name):
    dims, acts, _ = NETS[name]
    w = trained_weights(dims, acts, seed=2)
    X = np.random.RandomState(0).uniform(size=(1024, dims[0]))
    net = _engine(dims, acts, w)
    y = net.predict(X)
    assert y.shape == (1024, 1) and y.dtype == np.float32
    y32 = km.predict(w, acts, X)
    assert np.all(np.abs(y - y32) <= RTOL * np.abs(y32) + ATOL_F)


def test_weights_roundtrip_and_multi_model():
    from bore_b200.engine import NativeMLP
    dims, acts, _ = NETS["cfg2_hartmann6"]
    net = NativeMLP(dims, acts, n_models=3)
    ws = [km.init_weights(dims, s) for s in range(3)]
    for i, w in enumerate(ws):
        net.set_weights(w, model=i)
    X = np.random.RandomState(0).uniform(size=(64, 6))
    for i, w in enumerate(ws):
        got = net.get_weights(model=i)
        for a, b in zip(got, w):
            assert np.array_equal(a, b)
        y = net.predict(X, model=i)
        assert np.allclose(y, km.predict(w, acts, X), rtol=1e-5, atol=2e-6)


def test_empty_and_bad_inputs():
    from bore_b200.engine import NativeMLP
    from bore_b200._lib import BoreNativeError
    dims, acts, _ = NETS["cfg1_branin"]
    net = NativeMLP(dims, acts)
    assert net.predict(np.zeros((0, 2))).shape == (0, 1)
    with pytest.raises(BoreNativeError):
        NativeMLP([2, 16, 3], ["relu", "linear"])  # output dim must be 1
    with pytest.raises(BoreNativeError):
        NativeMLP([2, 1000, 1], ["relu", "linear"])  # hidden width limit
