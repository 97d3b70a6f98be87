"""K3 parity: batched on-device L-BFGS-B vs SciPy (the reference's optimiser).

Two levels:
  * stepper pinned to SciPy's ``setulb`` with IDENTICAL f,g (the oracle MLP evaluates both
    sides' requests): isolates the on-device algorithm from MLP rounding;
  * end to end (CUDA MLP + CUDA L-BFGS-B) vs ``scipy.optimize.minimize`` on the oracle MLP:
    the north_star criterion -- each start reaches the same local optimum within 1e-4 in
    objective value; target >= 95 % of starts.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import keras_mlp as km, argmax as am
from helpers import NETS, trained_weights, reference_self_agreement

pytestmark = pytest.mark.gpu

FUN_TOL = 1e-4  # north_star


def _stepper_vs_setulb(dims, acts, transform, S, seed, lo=0.0, hi=1.0):
    import torch
    from bore_b200 import _lib
    lib = _lib.require_cuda()
    n = dims[0]
    w = trained_weights(dims, acts, seed=seed)
    rs = np.random.RandomState(seed + 100)
    X0 = rs.uniform(size=(S, n))
    lo_a = np.full(n, lo, np.float64)
    hi_a = np.full(n, hi, np.float64)
    dev = torch.device("cuda", 0)
    nbytes = lib.bore_lbfgsb_workspace_bytes(S, n, 10)
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    xreq = torch.empty(S, n, dtype=torch.float64, device=dev)
    pend = torch.empty(S, dtype=torch.int32, device=dev)
    X0d = torch.from_numpy(X0).to(dev)
    P = lambda t: C.c_void_p(t.data_ptr())
    NP = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.bore_lbfgsb_init(P(X0d), S, n, NP(lo_a), NP(hi_a), 10, 1e-9, 1e-5, 1000, 15000,
                                    20, P(work), nbytes, P(xreq), P(pend), 0, None))
    ref = am.LockstepLBFGSB(X0, lo_a, hi_a)
    fd = torch.zeros(S, dtype=torch.float64, device=dev)
    gd = torch.zeros(S, n, dtype=torch.float64, device=dev)
    pending = C.c_int(S)
    rounds = 0
    while pending.value > 0 or ref.pending.any():
        if pending.value > 0:
            xr = xreq.cpu().numpy()
            f, g = km.value_and_input_grad(w, acts, xr, transform, True, np.float32)
            fd.copy_(torch.from_numpy(f.astype(np.float64)))
            gd.copy_(torch.from_numpy(g.astype(np.float64)))
            _lib.check(lib.bore_lbfgsb_step(P(fd), P(gd), 1, S, n, P(work), P(xreq), P(pend),
                                            C.byref(pending), 0, None))
        if ref.pending.any():
            fr, gr = km.value_and_input_grad(w, acts, ref.X[ref.pending], transform, True, np.float32)
            ref.feed(fr, gr)
        rounds += 1
        assert rounds < 20000
    x = torch.empty(S, n, dtype=torch.float64, device=dev)
    fun = torch.empty(S, dtype=torch.float64, device=dev)
    ints = torch.empty(4, S, dtype=torch.int32, device=dev)
    _lib.check(lib.bore_lbfgsb_results(S, n, P(work), P(x), P(fun), P(ints[0]), P(ints[1]),
                                       P(ints[2]), P(ints[3]), 0, None))
    got = dict(x=x.cpu().numpy(), fun=fun.cpu().numpy(), nit=ints[0].cpu().numpy(),
               nfev=ints[1].cpu().numpy(), status=ints[2].cpu().numpy(), task=ints[3].cpu().numpy())
    return got, ref.result()


def test_stepper_smooth_objective_tracks_setulb():
    """ELU net (plugin default): trajectories stay together -> same x, same nit/nfev for nearly
    every start (reduction order differs between a warp butterfly and setulb's serial loops, so
    a start may take one more/less step along a flat direction: fun agrees, x may not)."""
    dims, acts, transform = NETS["cfg5_plugin8"]
    got, ref = _stepper_vs_setulb(dims, acts, transform, S=96, seed=3)
    assert np.array_equal(got["status"], ref["status"])
    ref_task = np.where(ref["task"][:, 0] == 8, 0, ref["task"][:, 1])
    assert np.mean(got["task"] == ref_task) >= 0.95
    assert np.abs(got["fun"] - ref["fun"]).max() <= 1e-7
    dx = np.abs(got["x"] - ref["x"]).max(axis=1)
    print("stepper vs setulb: max|dfun|", np.abs(got["fun"] - ref["fun"]).max(), "dx quantiles",
          np.quantile(dx, [0.5, 0.95, 1.0]))
    assert np.mean(dx <= 1e-6) >= 0.90
    assert np.mean(dx <= 1e-4) >= 0.95
    assert dx.max() <= 1e-2
    # rounding-level differences (butterfly reductions, reciprocal multiplies, the complement
    # trick of formk) can flip the last convergence test of a start on this flat objective
    # (f is an fp32 value, so near the optimum the line search sees plateaus of equal f and an
    # ulp in x can cost or save an evaluation)
    print("nit equal", np.mean(got["nit"] == ref["nit"]), "nfev equal", np.mean(got["nfev"] == ref["nfev"]))
    assert np.mean(got["nit"] == ref["nit"]) >= 0.90
    assert np.mean(got["nfev"] == ref["nfev"]) >= 0.75
    assert np.mean(np.abs(got["nfev"] - ref["nfev"]) <= 3) >= 0.85
    assert np.abs(got["nit"] - ref["nit"]).max() <= 3


def _check_against_self_agreement(name, got, ref, w, acts, transform, X0, n):
    """ReLU objectives are piecewise linear: an ulp in f or g can flip a kink and send a start to
    another vertex, so the reference does not even agree with ITSELF under an fp32 re-association
    of the MLP.  The bar for us is that yardstick: agree with the reference (nearly) as often as
    it agrees with itself, and be statistically indistinguishable in the objective reached."""
    from scipy.optimize import Bounds
    agree = np.abs(got["fun"] - ref["fun"]) <= FUN_TOL
    one_sided = got["fun"] <= ref["fun"] + FUN_TOL
    r_self, _, alt = reference_self_agreement(w, acts, X0, Bounds(np.zeros(n), np.ones(n)),
                                              transform, FUN_TOL, ref=ref)
    d_ours = got["fun"] - ref["fun"]
    d_alt = alt["fun"] - ref["fun"]
    print(name, "agree", agree.mean(), "one-sided", one_sided.mean(), "reference self-agreement",
          r_self, "mean dfun ours", d_ours.mean(), "alt", d_alt.mean(),
          "status", np.bincount(got["status"], minlength=3), np.bincount(ref["status"], minlength=3))
    # (128 starts: the standard error of a rate near 0.9 is 2.7 points; the 2,048-start runs of
    # tests/test_gpu_parity_report.py hold the same comparison to 3 points)
    assert agree.mean() >= min(0.95, r_self - 0.07)
    # no systematic loss of solution quality: mean objective within 3 standard errors of the
    # spread the reference shows against itself
    se = max(np.std(d_alt), np.std(d_ours), 1e-6) / np.sqrt(len(d_ours))
    assert abs(d_ours.mean()) <= abs(d_alt.mean()) + 3.0 * se + 1e-5


@pytest.mark.parametrize("name", ["cfg1_branin", "cfg2_hartmann6", "cfg3_ackley50"])
def test_stepper_relu_objective_agreement(name):
    """Stepper alone (the oracle MLP answers both sides' requests) on ReLU nets."""
    dims, acts, transform = NETS[name]
    S = 128
    got, ref = _stepper_vs_setulb(dims, acts, transform, S=S, seed=5)
    w = trained_weights(dims, acts, seed=5)
    X0 = np.random.RandomState(5 + 100).uniform(size=(S, dims[0]))
    _check_against_self_agreement(name, got, ref, w, acts, transform, X0, dims[0])


@pytest.mark.parametrize("name", ["cfg5_plugin8", "cfg2_hartmann6", "cfg3_ackley50", "tanh_exp"])
def test_end_to_end_minimize_vs_scipy(name):
    from bore_b200.engine import NativeMLP
    from scipy.optimize import Bounds
    dims, acts, transform = NETS[name]
    n = dims[0]
    w = trained_weights(dims, acts, seed=7)
    S = 96
    X0 = np.random.RandomState(11).uniform(size=(S, n))
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    got = net.lbfgsb(X0, 0.0, 1.0, transform=transform)
    ref = am.minimize_starts(w, acts, X0, Bounds(np.zeros(n), np.ones(n)), transform=transform)
    agree = np.abs(got["fun"] - ref["fun"]) <= FUN_TOL
    print(name, "agree", agree.mean(), "rounds", got["rounds"],
          "evals", got["evals"], "nfev sum", got["nfev"].sum(), ref["nfev"].sum())
    if "relu" in acts:
        _check_against_self_agreement(name, got, ref, w, acts, transform, X0, n)
    else:
        assert agree.mean() >= 0.99  # smooth activations: the north_star bar, with margin
    # every returned point is feasible and its reported value is the model's value there
    assert np.all(got["x"] >= 0.0) and np.all(got["x"] <= 1.0)
    f_chk, _ = km.value_and_input_grad(w, acts, got["x"], transform, True, np.float32)
    assert np.abs(f_chk - got["fun"]).max() <= 1e-5
    assert set(np.unique(got["status"])) <= {0, 1, 2}


def test_unbounded_and_half_bounded():
    """nbd = 0 / 1 / 3 paths (SciPy Bounds with infinities)."""
    from bore_b200.engine import NativeMLP
    dims, acts, transform = NETS["tanh_exp"]
    n = dims[0]
    w = trained_weights(dims, acts, seed=9)
    X0 = np.random.RandomState(1).uniform(size=(32, n))
    lo = np.array([-np.inf, 0.0, -np.inf, 0.0, -1.0])
    hi = np.array([np.inf, np.inf, 1.0, 1.0, 2.0])
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    got = net.lbfgsb(X0, lo, hi, transform=transform, maxiter=200)
    from scipy.optimize import Bounds
    ref = am.minimize_starts(w, acts, X0, Bounds(lo, hi), options=dict(maxiter=200, ftol=1e-9),
                             transform=transform)
    rel = np.abs(got["fun"] - ref["fun"]) / np.maximum(1.0, np.abs(ref["fun"]))
    print("half-bounded rel diff", rel.max(), np.bincount(got["status"], minlength=3),
          np.bincount(ref["status"], minlength=3))
    assert np.mean(rel <= 1e-4) >= 0.9


def test_maxiter_limit_status():
    from bore_b200.engine import NativeMLP
    dims, acts, transform = NETS["cfg3_ackley50"]
    w = trained_weights(dims, acts, seed=4)
    X0 = np.random.RandomState(2).uniform(size=(16, 50))
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    got = net.lbfgsb(X0, 0.0, 1.0, transform=transform, maxiter=3)
    assert np.all(got["nit"] <= 3)
    assert np.all((got["status"] == 1) == (got["nit"] == 3) | (got["status"] != 1))
    assert np.any(got["status"] == 1)
    assert np.all(got["task"][got["status"] == 1] == 504)


def test_minimize_with_other_history_sizes_uses_the_generic_core():
    """maxcor != 10 runs the core compiled with a run-time m (maxcor == 10 runs the compile-time
    m = 10 build): both must follow SciPy with the same option."""
    from scipy.optimize import Bounds
    from bore_b200.engine import NativeMLP
    name = "cfg5_plugin8"
    dims, acts, transform = NETS[name]
    n = dims[0]
    w = trained_weights(dims, acts, seed=3)
    X0 = np.random.RandomState(9).uniform(size=(96, n))
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    for m in (3, 7, 10):
        got = net.lbfgsb(X0, 0.0, 1.0, transform=transform, m=m)
        ref = am.minimize_starts(w, acts, X0, Bounds(np.zeros(n), np.ones(n)), transform=transform,
                                 options=dict(maxiter=1000, ftol=1e-9, maxcor=m))
        agree = np.abs(got["fun"] - ref["fun"]) <= FUN_TOL
        print("maxcor", m, "agree", agree.mean(), "nit equal", np.mean(got["nit"] == ref["nit"]))
        assert agree.mean() >= 0.97, (m, agree.mean())
        assert np.mean(got["nit"] == ref["nit"]) >= 0.85
    # a ReLU net, where the history size changes the path of nearly every start (SciPy: nit differs
    # on 98 % of the starts between maxcor 3 and 10): the option must reach the device, and the
    # agreement is held to the reference's own self-agreement as everywhere on ReLU objectives
    name = "cfg2_hartmann6"
    dims, acts, transform = NETS[name]
    n = dims[0]
    w = trained_weights(dims, acts, seed=3)
    X0 = np.random.RandomState(9).uniform(size=(96, n))
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    nits = {}
    for m in (3, 10):
        got = net.lbfgsb(X0, 0.0, 1.0, transform=transform, m=m)
        ref = am.minimize_starts(w, acts, X0, Bounds(np.zeros(n), np.ones(n)), transform=transform,
                                 options=dict(maxiter=1000, ftol=1e-9, maxcor=m))
        nits[m] = got["nit"]
        _check_against_self_agreement(f"{name} maxcor={m}", got, ref, w, acts, transform, X0, n)
    assert np.mean(nits[3] != nits[10]) >= 0.5


@pytest.mark.parametrize("mode", [1, 2])
def test_non_finite_objective_ends_abnormal(mode):
    """transform = exp (bore/plugins/hpbandster/base.py:18) overflows fp32 where the logit is
    below -88: SciPy's line search degenerates on inf / NaN and the start ends ABNORMAL, which the
    reference's argmax drops (bore/mixins.py:83-85).  The device guard ends such a start with
    status 2 at once -- in both argmax paths -- and the finite starts are not disturbed."""
    from bore_b200 import _lib
    from bore_b200.engine import NativeMLP
    lib = _lib.require_cuda()
    dims, acts = [2, 8, 1], ["relu", "linear"]
    w = [np.array([[1.0, -1.0, 0.5, 0, 0, 0, 0, 0], [0.5, 1.0, -1.0, 0, 0, 0, 0, 0]], np.float32),
         np.zeros(8, np.float32),
         np.array([[-400.0], [-300.0], [-200.0], [0], [0], [0], [0], [0]], np.float32), np.zeros(1, np.float32)]
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    X0 = np.random.RandomState(0).uniform(size=(256, 2))
    f0, _ = km.value_and_input_grad(w, acts, X0, "exp", True, np.float32)
    bad0 = ~np.isfinite(f0)
    assert bad0.any() and (~bad0).any()
    _lib.check(lib.bore_lbfgsb_set_mode(mode))
    try:
        net._work = None
        got = net.lbfgsb(X0, 0.0, 1.0, transform="exp")
    finally:
        _lib.check(lib.bore_lbfgsb_set_mode(0))
    assert set(np.unique(got["status"])) <= {0, 1, 2}
    assert np.all(got["status"][bad0] == 2) and np.all(got["nit"][bad0] == 0) and np.all(got["nfev"][bad0] == 1)
    assert np.array_equal(got["x"][bad0], X0[bad0])
    good = got["status"] != 2
    assert good.any() and np.all(np.isfinite(got["fun"][good]))
    assert np.all(got["x"] >= 0.0) and np.all(got["x"] <= 1.0)


def test_nan_from_a_callers_objective_through_the_stepper():
    """bore_lbfgsb_step with f = NaN (or an inf in g) for some starts: those end with status 2 on
    the step that sees it, at the last good iterate; the others run to convergence."""
    import torch
    from bore_b200 import _lib
    lib = _lib.require_cuda()
    S, n = 64, 4
    rs = np.random.RandomState(3)
    X0 = rs.uniform(-1, 1, size=(S, n))
    c = rs.uniform(-0.5, 0.5, size=(S, n))
    lo_a, hi_a = np.full(n, -1.0), np.full(n, 1.0)
    dev = torch.device("cuda", 0)
    nbytes = lib.bore_lbfgsb_workspace_bytes(S, n, 10)
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    xreq = torch.empty(S, n, dtype=torch.float64, device=dev)
    pend = torch.empty(S, dtype=torch.int32, device=dev)
    X0d = torch.from_numpy(X0).to(dev)
    P = lambda t: C.c_void_p(t.data_ptr())
    NP = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.bore_lbfgsb_init(P(X0d), S, n, NP(lo_a), NP(hi_a), 10, 1e-9, 1e-5, 1000, 15000,
                                    20, P(work), nbytes, P(xreq), P(pend), 0, None))
    fd = torch.zeros(S, dtype=torch.float64, device=dev)
    gd = torch.zeros(S, n, dtype=torch.float64, device=dev)
    pending = C.c_int(S)
    poisoned_f, poisoned_g = np.arange(S) % 8 == 1, np.arange(S) % 8 == 5
    rounds = 0
    x_before = {}
    while pending.value > 0:
        xr = xreq.cpu().numpy()
        f = np.sum((xr - c) ** 4 + (xr - c) ** 2, axis=1)
        g = 4 * (xr - c) ** 3 + 2 * (xr - c)
        if rounds == 2:  # in mid line search / iteration
            f[poisoned_f] = np.nan
            g[poisoned_g, 1] = np.inf
        fd.copy_(torch.from_numpy(f)); gd.copy_(torch.from_numpy(g))
        _lib.check(lib.bore_lbfgsb_step(P(fd), P(gd), 1, S, n, P(work), P(xreq), P(pend),
                                        C.byref(pending), 0, None))
        rounds += 1
        assert rounds < 500
    x = torch.empty(S, n, dtype=torch.float64, device=dev)
    fun = torch.empty(S, dtype=torch.float64, device=dev)
    ints = torch.empty(4, S, dtype=torch.int32, device=dev)
    _lib.check(lib.bore_lbfgsb_results(S, n, P(work), P(x), P(fun), P(ints[0]), P(ints[1]),
                                       P(ints[2]), P(ints[3]), 0, None))
    status = ints[2].cpu().numpy()
    bad = poisoned_f | poisoned_g
    assert np.all(status[bad] == 2)
    assert np.all(status[~bad] == 0)
    assert np.all(np.isfinite(fun.cpu().numpy()))     # the value at the restored iterate
    xs = x.cpu().numpy()
    assert np.all(np.isfinite(xs)) and np.all(xs >= -1.0) and np.all(xs <= 1.0)
    assert np.abs(xs[~bad] - np.clip(c[~bad], -1, 1)).max() <= 1e-3
