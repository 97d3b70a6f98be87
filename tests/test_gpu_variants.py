"""Kernel variants that are selected once per process (environment): the run-time-shape build of the unit-split
cluster fit, and the scalar-FFMA build of K2 that the packed FFMA2 build must equal to the bit.  Each variant runs
in a process of its own."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FIT_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from oracle import keras_mlp as km
from helpers import NETS, synthetic_targets
from bore_b200.engine import NativeMLP
out = {{}}
for name, N, E in (("cfg3_ackley50", 500, 6), ("cfg2_hartmann6", 300, 10), ("cfg5_plugin8", 333, 8)):
    dims, acts, _ = NETS[name]
    rs = np.random.RandomState(5)
    X = rs.uniform(size=(N, dims[0])); y = synthetic_targets(X); z = y < np.quantile(y, 0.25)
    perms = np.stack([rs.permutation(N) for _ in range(E)])
    w0 = km.init_weights(dims, 3)
    w_ref = [w.copy() for w in w0]
    h_ref, _ = km.fit(w_ref, acts, X, z, E, 64, perms)
    net = NativeMLP(dims, acts); net.set_fit_mode(4); net.set_weights(w0)
    h = net.fit(X, z, E, 64, perms)
    out[name] = float(np.abs(h - h_ref).max())
    assert out[name] <= 1e-4, (name, out[name])
    for a, b in zip(net.get_weights(), w_ref):
        assert np.abs(a - b).max() <= 2e-3
print("ok", out)
"""

K2_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from oracle import keras_mlp as km
from helpers import NETS
from bore_b200.engine import NativeMLP
res = []
for name in ("cfg3_ackley50", "cfg2_hartmann6", "cfg5_plugin8", "tanh_exp"):
    dims, acts, transform = NETS[name]
    net = NativeMLP(dims, acts); net.set_weights(km.init_weights(dims, 7))
    X = np.random.RandomState(11).uniform(-1, 2, size=(1037, dims[0]))
    f, g = net.value_and_grad(X, transform, True)
    res += [np.asarray(f), np.asarray(g), np.asarray(net.predict(X))]
np.savez({out!r}, *res)
"""


def _run(script, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_unit_split_fit_runtime_shape_build_matches_oracle():
    """csrc/fit_unit.cu is compiled for run-time shapes and again with the shape as template constants; the latter is
    what tests/test_gpu_fit.py reaches for the BASELINE nets, this is the former on the same nets."""
    out = _run(FIT_SCRIPT.format(root=ROOT), {"BORE_FIT_UNIT_GENERIC": "1"})
    assert out.strip().startswith("ok")


def test_k2_packed_ffma2_equals_scalar_ffma_to_the_bit(tmp_path):
    """fma.rn.f32x2 is two IEEE FMAs: values, input gradients and predictions of the packed build (default) and of the
    scalar build (BORE_K2_FFMA2=0) are EQUAL -- which is what keeps K2 interchangeable with the fused kernel's own
    evaluation in the tail handover (tests/test_gpu_fused.py)."""
    a, b = str(tmp_path / "packed.npz"), str(tmp_path / "scalar.npz")
    _run(K2_SCRIPT.format(root=ROOT, out=a), {"BORE_K2_FFMA2": "1"})
    _run(K2_SCRIPT.format(root=ROOT, out=b), {"BORE_K2_FFMA2": "0"})
    pa, pb = np.load(a), np.load(b)
    assert len(pa.files) == len(pb.files) == 12
    for k in pa.files:
        assert np.array_equal(pa[k], pb[k]), k
