"""The on-device L-BFGS-B algorithm (bore_b200/csrc/lbfgsb_core.h) compiled for the host and
pinned against SciPy's ``setulb`` request by request with identical f,g.  CPU only.

This is the strictest parity level: on a smooth (ELU) objective the two
implementations request bit-comparable points until convergence; on piecewise-linear (ReLU)
objectives rounding-level differences eventually flip a kink, so the gate there is the
north_star rate on the final objective value."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import keras_mlp as km, argmax as am
from helpers import NETS, trained_weights

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "hostsim", "lbfgsb_hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(HERE, "..", "bore_b200", "csrc", f)
                    for f in ("lbfgsb_core.h", "lbfgsb_types.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.hs_create.restype = C.c_void_p
    lib.hs_create.argtypes = [C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.hs_start.argtypes = [C.c_void_p, dp, dp]
    lib.hs_step.argtypes = [C.c_void_p, C.c_double, dp, dp]
    lib.hs_result.argtypes = [C.c_void_p, dp, dp] + [C.POINTER(C.c_int)] * 4
    lib.hs_destroy.argtypes = [C.c_void_p]
    lib.hs_set_split.argtypes = [C.c_void_p, C.c_int]
    return lib


def _P(a):
    return a.ctypes.data_as(dp)


def _run(hs, w, acts, transform, X0, lo, hi, maxiter=1000, split=False):
    S, n = X0.shape
    keys = ("x", "fun", "nit", "nfev", "status")
    got = dict(x=np.zeros((S, n)), fun=np.zeros(S), nit=np.zeros(S, int), nfev=np.zeros(S, int),
               status=np.zeros(S, int))
    ref = {k: v.copy() for k, v in got.items()}
    max_req_dx = np.zeros(S)
    for s in range(S):
        h = hs.hs_create(n, 10, _P(lo), _P(hi), 1e-9, 1e-5, maxiter, 15000, 20)
        hs.hs_set_split(h, 1 if split else 0)
        ls = am.LockstepLBFGSB(X0[s:s + 1], lo, hi, maxiter=maxiter)
        xr = np.zeros(n)
        pend = hs.hs_start(h, _P(np.ascontiguousarray(X0[s])), _P(xr))
        together = True
        while pend or ls.pending.any():
            if pend and ls.pending.any() and together:
                dx = np.abs(ls.X[0] - xr).max()
                if dx > 1e-9:
                    together = False
                else:
                    max_req_dx[s] = max(max_req_dx[s], dx)
            if ls.pending.any():
                fr, gr = km.value_and_input_grad(w, acts, ls.X[0][None], transform)
                ls.feed(fr, gr)
            if pend:
                f, g = km.value_and_input_grad(w, acts, xr[None], transform)
                pend = hs.hs_step(h, float(f[0]), _P(np.ascontiguousarray(g[0].astype(np.float64))), _P(xr))
        x = np.zeros(n); f = C.c_double(); ii = [C.c_int() for _ in range(4)]
        hs.hs_result(h, _P(x), C.byref(f), *[C.byref(v) for v in ii])
        hs.hs_destroy(h)
        got["x"][s], got["fun"][s] = x, f.value
        got["nit"][s], got["nfev"][s], got["status"][s] = ii[0].value, ii[1].value, ii[2].value
        r = ls.result()
        for k in keys:
            ref[k][s] = r[k][0]
    return got, ref


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("name", ["cfg5_plugin8", "tanh_exp"])
def test_smooth_objectives_track_setulb(hs, name, split):
    """`split` runs every step as the LIGHT stage + (if requested) the HEAVY stage, the way the
    thread-per-start kernel does."""
    dims, acts, transform = NETS[name]
    n = dims[0]
    w = trained_weights(dims, acts, seed=1)
    X0 = np.random.RandomState(2).uniform(size=(40, n))
    got, ref = _run(hs, w, acts, transform, X0, np.zeros(n), np.ones(n), split=split)
    assert np.array_equal(got["status"], ref["status"])
    assert np.abs(got["fun"] - ref["fun"]).max() <= 1e-7
    assert np.abs(got["x"] - ref["x"]).max() <= 1e-6
    assert np.mean(got["nit"] == ref["nit"]) >= 0.95
    assert np.mean(got["nfev"] == ref["nfev"]) >= 0.95


@pytest.mark.parametrize("name", ["cfg1_branin", "cfg2_hartmann6", "cfg3_ackley50"])
def test_relu_objectives_agree_at_the_north_star_rate(hs, name):
    dims, acts, transform = NETS[name]
    n = dims[0]
    w = trained_weights(dims, acts, seed=1)
    S = 64 if n < 50 else 24
    X0 = np.random.RandomState(2).uniform(size=(S, n))
    got, ref = _run(hs, w, acts, transform, X0, np.zeros(n), np.ones(n), split=(n == 6))
    agree = np.abs(got["fun"] - ref["fun"]) <= 1e-4
    assert agree.mean() >= 0.90, agree.mean()
    assert np.mean(got["status"] == ref["status"]) >= 0.90


def test_unbounded_half_bounded_and_fixed(hs):
    """nbd 0/1/2/3 and a fixed variable (lo == hi)."""
    dims, acts, transform = NETS["tanh_exp"]
    w = trained_weights(dims, acts, seed=3)
    lo = np.array([-np.inf, 0.0, -np.inf, 0.3, -1.0])
    hi = np.array([np.inf, np.inf, 1.0, 0.3, 2.0])
    X0 = np.random.RandomState(4).uniform(size=(24, 5))
    got, ref = _run(hs, w, acts, transform, X0, lo, hi, maxiter=300)
    rel = np.abs(got["fun"] - ref["fun"]) / np.maximum(1.0, np.abs(ref["fun"]))
    assert np.mean(rel <= 1e-6) >= 0.9
    assert np.all(got["x"][:, 3] == 0.3)


def test_maxiter_status(hs):
    dims, acts, transform = NETS["cfg3_ackley50"]
    w = trained_weights(dims, acts, seed=1)
    X0 = np.random.RandomState(2).uniform(size=(6, 50))
    got, ref = _run(hs, w, acts, transform, X0, np.zeros(50), np.ones(50), maxiter=2)
    assert np.array_equal(got["status"], ref["status"]) and np.array_equal(got["nit"], ref["nit"])
    assert np.any(got["status"] == 1)


def test_scipy_bound_violation_regression(hs):
    """SciPy's own known-answer for this path (scipy/optimize/tests/test_lbfgsb_setulb.py:70-113):
    7 setulb steps on a tabulated objective that used to step outside [0,1] by rounding.  Our
    stepper must request the same points as setulb, all of them feasible."""
    from scipy.optimize import _lbfgsb
    from scipy.optimize.tests.test_lbfgsb_setulb import objfun
    n, m = 5, 10
    lo, hi = np.zeros(n), np.ones(n)
    x0 = np.array([0.8750000000000278, 0.7500000000000153, 0.9499999999999722,
                   0.8214285714285992, 0.6363636363636085])
    # setulb side
    idt = np.int32
    x = x0.copy()
    nbd = np.full(n, 2, idt)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
    iwa, task, ln_task = np.zeros(3 * n, idt), np.zeros(2, idt), np.zeros(2, idt)
    lsave, isave, dsave = np.zeros(4, idt), np.zeros(44, idt), np.zeros(29)
    ref_requests = []
    for _ in range(7):
        f, g = objfun(x)
        try:
            _lbfgsb.setulb(m, x, lo, hi, nbd, f, g, 1e7, 1e-5, wa, iwa, task, lsave, isave, dsave,
                           20, ln_task)
        except TypeError:
            pytest.skip("this SciPy's setulb has a different signature")
        if task[0] == 3:
            ref_requests.append(x.copy())
    # our side: same factr*epsmch, identical tabulated f,g
    h = hs.hs_create(n, m, _P(lo), _P(hi), 1e7 * np.finfo(float).eps, 1e-5, 1000, 15000, 20)
    xr = np.zeros(n)
    hs.hs_start(h, _P(x0.copy()), _P(xr))
    ours = [xr.copy()]
    for _ in range(len(ref_requests) - 1):
        f, g = objfun(xr)
        assert hs.hs_step(h, float(f), _P(np.ascontiguousarray(g, np.float64)), _P(xr)) == 1
        ours.append(xr.copy())
    hs.hs_destroy(h)
    for a, b in zip(ours, ref_requests):
        assert np.all(a >= lo) and np.all(a <= hi)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)
