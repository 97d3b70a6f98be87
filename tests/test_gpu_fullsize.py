"""Parity at BASELINE.json's FULL sizes (configs[2]: 50-D, Dense64x3, 2,000 observations, 65,536
starts) through size-independent properties -- the oracle cannot run 65,536 SciPy minimisations
in a test, so what is checked on all of them is what must hold for ANY correct run:

  * feasibility: every returned point lies in the box;
  * consistency: the reported value is the model's value at the returned point (K2 re-evaluated,
    and the NumPy oracle on a sample);
  * descent: no start ends above where it began;
  * stationarity: starts that stopped on the projected-gradient test (task 401) satisfy it;
  * idempotence: restarting from a converged point stops at once, at the same point;
  * accounting: sum(nfev) equals the evaluations the driver counted; statuses are SciPy's;
  * selection: the first-minimum key picks min(fun) over the eligible starts;
  * chunking invariance of K2 (one launch of 65,536 points == 16 launches of 4,096, bit for bit);
and on a random SAMPLE of the 65,536 starts the per-start agreement with SciPy on the oracle MLP
(north_star: objective within 1e-4), gated like tests/test_gpu_lbfgsb.py on the reference's own
self-agreement for ReLU nets.  The full-size fit (31 epochs x 32 steps) is compared with the
oracle's loss trajectory in both kernel mappings.
"""
import numpy as np
import pytest

from oracle import keras_mlp as km, argmax as am
from helpers import NETS, reference_self_agreement, permuted_units

pytestmark = pytest.mark.gpu

DIMS, ACTS, TRANSFORM = NETS["cfg3_ackley50"]
S_FULL, N_OBS, EPOCHS, BATCH = 65536, 2000, 31, 64
FUN_TOL = 1e-4   # north_star, objective value per start
LOSS_TOL = 1e-4  # north_star, training loss trajectory


def _ackley(X):
    u = -32.768 + 65.536 * X
    d = X.shape[1]
    return (-20.0 * np.exp(-0.2 * np.sqrt(np.sum(u * u, axis=1) / d))
            - np.exp(np.sum(np.cos(2 * np.pi * u), axis=1) / d) + 20.0 + np.e)


def _problem(seed=0):
    rs = np.random.RandomState(seed)
    X = rs.uniform(size=(N_OBS, DIMS[0]))
    y = _ackley(X)
    z = y < np.quantile(y, 0.25)
    perms = np.stack([rs.permutation(N_OBS) for _ in range(EPOCHS)]).astype(np.int32)
    return X, z, perms


@pytest.fixture(scope="module")
def trained():
    """cfg-3 classifier trained by the CUDA fit kernel at full size + the oracle's run."""
    from bore_b200.engine import NativeMLP
    X, z, perms = _problem()
    w0 = km.init_weights(DIMS, 0)
    net = NativeMLP(DIMS, ACTS)
    net.set_weights(w0)
    hist = net.fit(X, z, EPOCHS, BATCH, perms)
    w_ref = [w.copy() for w in w0]
    hist_ref, _ = km.fit(w_ref, ACTS, X, z, EPOCHS, BATCH, perms)
    return net, hist, hist_ref


@pytest.fixture(scope="module")
def full_run(trained):
    import torch
    net = trained[0]
    X0 = np.random.RandomState(1).uniform(size=(S_FULL, DIMS[0]))
    X0d = net.to_device(X0, np.float64)
    res = net.lbfgsb_dev(X0d, 0.0, 1.0, transform=TRANSFORM, m=10, ftol=1e-9, gtol=1e-5,
                         maxiter=1000, maxfun=15000, maxls=20)
    torch.cuda.synchronize()
    host = {k: res[k].cpu().numpy() for k in ("x", "fun", "nit", "nfev", "status", "task")}
    host["evals"], host["rounds"] = res["evals"], res["rounds"]
    return X0, res, host


@pytest.fixture(scope="module")
def oracle_fit():
    """The oracle's full-size run, and the same run with the hidden units permuted -- the
    identical computation up to the ORDER of every fp32 sum.  992 Adam steps through ReLU kinks
    amplify that rounding noise: the two oracle runs agree to 1e-7 for the first ~18 epochs and
    then drift apart by up to 7e-3 (measured; fp32 vs fp64 arithmetic: 9e-3).  That drift is the
    yardstick for "same trajectory" once it exceeds the 1e-4 bar itself."""
    X, z, perms = _problem()
    w0 = km.init_weights(DIMS, 0)
    w_ref = [w.copy() for w in w0]
    hist_ref, adam_ref = km.fit(w_ref, ACTS, X, z, EPOCHS, BATCH, perms)
    w_alt = permuted_units(w0, seed=1)
    hist_alt, _ = km.fit(w_alt, ACTS, X, z, EPOCHS, BATCH, perms)
    w64 = [w.astype(np.float64) for w in w0]
    hist64, _ = km.fit(w64, ACTS, X, z, EPOCHS, BATCH, perms, dtype=np.float64)
    return (hist_ref, adam_ref, np.abs(hist_ref - hist_alt), hist64,
            np.abs(hist_ref.astype(np.float64) - hist64))


@pytest.mark.parametrize("mode", [1, 2])
def test_full_size_fit_tracks_the_oracle(mode, oracle_fit):
    from bore_b200.engine import NativeMLP
    hist_ref, adam_ref, self_diff, hist64, or32_vs_64 = oracle_fit
    X, z, perms = _problem()
    net = NativeMLP(DIMS, ACTS)
    net.set_fit_mode(mode)
    net.set_weights(km.init_weights(DIMS, 0))
    hist = net.fit(X, z, EPOCHS, BATCH, perms)
    assert hist.shape == (EPOCHS,)
    diff = np.abs(hist - hist_ref)
    # (1) the north_star bar (1e-4) on every epoch before the reference's own re-association
    #     noise becomes visible
    #     (the onset of the divergence is itself a matter of rounding: allow it JITTER epochs
    #     of slack -- the noise grows from 1e-7 to 1e-3 within about three epochs)
    JITTER = 4
    onset = int(np.argmax(self_diff > 1e-5)) if (self_diff > 1e-5).any() else EPOCHS
    stable = onset - JITTER
    assert stable >= 10, "the oracle itself should be reproducible for the first epochs"
    assert diff[:stable].max() <= LOSS_TOL, (stable, diff[:stable].max())
    # (2) afterwards rounding noise has been amplified through the ReLU kinks and the question is
    #     which of two fp32 runs is "the" trajectory: judge both against the fp64 oracle -- the CUDA
    #     kernel may be at most twice as far from it as the fp32 oracle is (same slack on the time
    #     axis), i.e. it is one more fp32 realisation of the same computation
    #     (over the whole trajectory: WHEN the amplified noise shows up differs by an epoch or
    #     three between any two fp32 runs, so an epoch-by-epoch envelope only measures that jitter)
    d_gpu64 = np.abs(hist.astype(np.float64) - hist64)
    assert d_gpu64.max() <= 2.0 * or32_vs_64.max(), (d_gpu64.max(), or32_vs_64.max())
    print("full-size fit mode", mode, "stable epochs", stable, "max diff vs fp32 oracle", diff.max(),
          "vs fp64 oracle", d_gpu64.max(), "fp32 oracle vs fp64 oracle", or32_vs_64.max())
    assert net.get_adam_state()[2] == adam_ref.t == EPOCHS * (-(-N_OBS // BATCH)) == 992
    assert hist[-1] < hist[0]  # the classifier learned something


def test_full_size_tensor_pipe_fit_is_not_fp32(oracle_fit):
    """Fit mode 3 (3xTF32 mma.sync, csrc/fit_mma.cu) at full size.  Its products carry ~2^-21 relative error
    instead of fp32's 2^-24, so the amplified rounding noise of the 992-step run shows up EARLIER than in any
    fp32 run: measured, the trajectory leaves the 1e-4 band at epoch 6 (the FFMA kernels and the fp32 oracle
    stay inside it until epoch ~18) -- one of the two reasons the kernel is not the default (the other: it is
    slower).  This test pins that finding: exact start, same optimum, earlier drift."""
    from bore_b200.engine import NativeMLP
    hist_ref, adam_ref, self_diff, hist64, or32_vs_64 = oracle_fit
    X, z, perms = _problem()
    net = NativeMLP(DIMS, ACTS)
    net.set_fit_mode(3)
    net.set_weights(km.init_weights(DIMS, 0))
    hist = net.fit(X, z, EPOCHS, BATCH, perms)
    diff = np.abs(hist - hist_ref)
    assert diff[:4].max() <= LOSS_TOL, diff[:4].max()
    d_gpu64 = np.abs(hist.astype(np.float64) - hist64)
    assert d_gpu64.max() <= 5e-2 and hist[-1] < 0.5 * hist[0]
    print("full-size fit mode 3: first epoch outside 1e-4:", int(np.argmax(diff > LOSS_TOL)) if (diff > LOSS_TOL).any() else None,
          "max diff vs fp32 oracle", diff.max(), "vs fp64 oracle", d_gpu64.max(), "fp32 oracle vs fp64", or32_vs_64.max())
    assert net.get_adam_state()[2] == adam_ref.t


def test_full_size_feasible_consistent_descending(trained, full_run):
    net = trained[0]
    X0, res, h = full_run
    x, fun = h["x"], h["fun"]
    assert x.shape == (S_FULL, DIMS[0])
    assert np.all(x >= 0.0) and np.all(x <= 1.0)
    assert np.all(np.isfinite(fun))
    assert set(np.unique(h["status"])) <= {0, 1, 2}
    assert np.all(h["nit"] >= 0) and np.all(h["nfev"] >= 1) and np.all(h["nfev"] >= h["nit"])
    assert np.all(h["nit"] <= 1000) and np.all(h["nfev"] <= 15000 + 20)
    # accounting: the driver counts every evaluation K2 performed; nfev is SciPy's count, which
    # does not re-count a request that repeats the previous point (ScalarFunction memoises on x:
    # a line search restarted from the same iterate) -- so evals >= sum(nfev), and close to it
    assert int(h["nfev"].sum()) <= int(h["evals"]) <= int(1.02 * h["nfev"].sum())
    # consistency: K2 at the returned points reproduces the reported values (fp32 values)
    f_x, g_x = net.value_and_grad(x.astype(np.float32), TRANSFORM, True)
    assert np.abs(f_x - fun).max() <= 1e-6
    # ... and so does the NumPy oracle on a sample
    idx = np.random.RandomState(3).choice(S_FULL, 512, replace=False)
    f_o, _ = km.value_and_input_grad(net.get_weights(), ACTS, x[idx], TRANSFORM, True, np.float32)
    assert np.abs(f_o - fun[idx]).max() <= 1e-5
    # descent: no start ends above where it began (fp32 objective: allow its rounding)
    f_0, _ = net.value_and_grad(X0.astype(np.float32), TRANSFORM, True)
    assert np.all(fun <= f_0 + 1e-6)
    # stationarity: starts stopped by the projected-gradient test satisfy it at the returned point
    pg = g_x.astype(np.float64).copy()
    neg = pg < 0
    pg[neg] = np.maximum(x[neg] - 1.0, pg[neg])
    pg[~neg] = np.minimum(x[~neg] - 0.0, pg[~neg])
    sb = np.abs(pg).max(axis=1)
    by_pg = (h["status"] == 0) & (h["task"] == 401)
    assert by_pg.any()
    assert sb[by_pg].max() <= 1e-5 + 1e-7


def test_full_size_selection_is_the_first_minimum(trained, full_run):
    net = trained[0]
    _, res, h = full_run
    key = net.select_best(res["fun"], res["status"])
    ok = (h["status"] == 0) | (h["status"] == 1)
    best = np.flatnonzero(ok)[np.argmin(h["fun"][ok])]  # first minimum among the eligible
    from bore_b200 import distributed as bd
    gidx, rec = bd.global_winner(key, lambda i: res["fun"][i:i + 1], S_FULL, 1)
    assert int(gidx) == int(best)
    assert float(rec[0]) == float(h["fun"][best])


def test_full_size_restart_from_converged_points_is_idempotent(trained, full_run):
    net = trained[0]
    _, _, h = full_run
    conv = np.flatnonzero((h["status"] == 0) & (h["task"] == 401))[:4096]
    again = net.lbfgsb(h["x"][conv], 0.0, 1.0, transform=TRANSFORM)
    assert np.all(again["status"] == 0)
    assert np.all(again["nit"] == 0) and np.all(again["nfev"] == 1)
    assert np.array_equal(again["x"], h["x"][conv])
    assert np.abs(again["fun"] - h["fun"][conv]).max() <= 1e-7


def test_full_size_sample_agrees_with_scipy(trained, full_run):
    from helpers import parallel_minimize_starts
    net = trained[0]
    X0, _, h = full_run
    SAMPLE = 1024  # s.e. of an agreement rate near 0.96: 0.6 points
    idx = np.random.RandomState(7).choice(S_FULL, SAMPLE, replace=False)
    w = net.get_weights()
    ref = parallel_minimize_starts(w, ACTS, X0[idx], 0.0, 1.0, TRANSFORM)
    alt = parallel_minimize_starts(permuted_units(w), ACTS, X0[idx], 0.0, 1.0, TRANSFORM)
    self_rate = float(np.mean(np.abs(alt["fun"] - ref["fun"]) <= FUN_TOL))
    self_one = float(np.mean(alt["fun"] <= ref["fun"] + FUN_TOL))
    agree = np.abs(h["fun"][idx] - ref["fun"]) <= FUN_TOL
    one_sided = h["fun"][idx] <= ref["fun"] + FUN_TOL
    print("full-size sample: agree", agree.mean(), "one-sided", one_sided.mean(),
          "reference self-agreement", self_rate, "self one-sided", self_one)
    # ReLU objective: the yardstick is how well the reference agrees with itself under an fp32
    # re-association of the MLP -- within 3 points of it, and never below 0.93
    assert agree.mean() >= min(0.95, self_rate - 0.03)
    assert agree.mean() >= 0.93
    assert one_sided.mean() >= min(0.95, self_one - 0.03)
    ab_o, ab_r = np.mean(h["status"][idx] == 2), np.mean(ref["status"] == 2)
    assert abs(ab_o - ab_r) <= 0.03


def test_k2_chunking_invariance_at_full_size(trained):
    """65,536 x 50 points in one launch == the same points in 16 launches, bit for bit, and the
    oracle's values on a sample (<= 1e-5 relative, the north_star bar for value and gradient)."""
    net = trained[0]
    X = np.random.RandomState(5).uniform(size=(S_FULL, DIMS[0])).astype(np.float32)
    f_all, g_all = net.value_and_grad(X, TRANSFORM, True)
    for c in range(0, S_FULL, 4096 * 4):
        f_c, g_c = net.value_and_grad(X[c:c + 4096], TRANSFORM, True)
        assert np.array_equal(f_c, f_all[c:c + 4096])
        assert np.array_equal(g_c, g_all[c:c + 4096])
    idx = np.random.RandomState(6).choice(S_FULL, 1024, replace=False)
    f_o, g_o = km.value_and_input_grad(net.get_weights(), ACTS, X[idx].astype(np.float64), TRANSFORM,
                                       True, np.float32)
    assert np.abs(f_all[idx] - f_o).max() <= 1e-5 * max(1.0, np.abs(f_o).max())
    scale = max(np.abs(g_o).max(), 1e-30)
    assert np.abs(g_all[idx] - g_o).max() <= 1e-5 * scale + 1e-7
