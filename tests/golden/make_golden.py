"""Generates the committed fixtures under tests/golden/ -- run in the BUILD container only.

Two sources:
  1. the reference's own importable modules (bore.math, bore.data, bore.optimizers.utils under
     /root/reference; numpy/scipy only) -> `host_golden.json`: known answers for the host-side
     helpers of the path, produced by the reference code itself;
  2. the installed SciPy L-BFGS-B (the reference's optimiser, scipy/optimize/_lbfgsb_py.py)
     driven by the oracle MLP -> `lbfgsb_golden.npz`: per-start x*, fun, nit, nfev, status on
     seeded problems, with the SciPy/NumPy versions recorded.

/root/reference does not exist on the GPU box: tests only ever read the committed outputs.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def host_golden():
    sys.path.insert(0, "/root/reference")
    from bore.math import steps_per_epoch, ceil_divide          # noqa: E402
    from bore.data import Record                                 # noqa: E402
    from bore.optimizers.utils import from_bounds                # noqa: E402
    from scipy.optimize import Bounds

    out = {"source": "ltiao/bore v1.5.0 modules imported from /root/reference",
           "numpy": np.__version__, "scipy": scipy.__version__}
    out["steps_per_epoch"] = [[n, b, steps_per_epoch(n, b)] for n in
                              (1, 10, 32, 63, 64, 65, 100, 110, 127, 128, 500, 1000, 2000, 4096)
                              for b in (1, 32, 64, 100)]
    out["ceil_divide"] = [[a, b, int(ceil_divide(a, b))] for a in (-7, -1, 0, 1, 7, 64, 65)
                          for b in (1, 2, 64)]
    # from_bounds on both input kinds
    fb = []
    for lo, hi in (([0.0, 0.0], [1.0, 1.0]), ([-5.0, 0.0, 2.0], [10.0, 15.0, 3.0])):
        (l1, h1), d1 = from_bounds(Bounds(np.array(lo), np.array(hi)))
        (l2, h2), d2 = from_bounds(list(zip(lo, hi)))
        fb.append(dict(lo=lo, hi=hi, bounds_obj=[list(map(float, l1)), list(map(float, h1)), d1],
                       pairs=[list(map(float, l2)), list(map(float, h2)), d2]))
    out["from_bounds"] = fb
    # Record: quantile labelling and duplicate test
    recs = []
    for seed, n, gamma in ((0, 10, 0.25), (1, 37, 1 / 3), (2, 110, 0.25), (3, 8, 0.5)):
        rs = np.random.RandomState(seed)
        X = rs.uniform(size=(n, 3))
        y = rs.normal(size=n)
        y[:3] = y[0]  # ties around the threshold exercise the strict '<'
        rec = Record()
        for xi, yi in zip(X, y):
            rec.append(x=xi, y=yi, b=1.0)
        Xc, z = rec.load_classification_data(gamma)
        probes = [X[0], X[0] + 1e-9, X[0] + 1e-4, X[-1] * (1 + 5e-6), rs.uniform(size=3)]
        recs.append(dict(seed=seed, n=n, gamma=gamma, X=X.tolist(), y=y.tolist(),
                         z=[bool(v) for v in z], size=rec.size(),
                         probes=[p.tolist() for p in probes],
                         is_duplicate=[bool(rec.is_duplicate(p)) for p in probes]))
    out["record"] = recs
    with open(os.path.join(HERE, "host_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("host_golden.json written")


def lbfgsb_golden():
    from oracle import keras_mlp as km, argmax as am
    from helpers import NETS, trained_weights
    from scipy.optimize import Bounds
    out = dict(scipy_version=np.array(scipy.__version__), numpy_version=np.array(np.__version__))
    for name in ("cfg1_branin", "cfg2_hartmann6", "cfg5_plugin8", "cfg3_ackley50"):
        dims, acts, transform = NETS[name]
        n = dims[0]
        w = trained_weights(dims, acts, seed=31)
        S = 24
        X0 = np.random.RandomState(77).uniform(size=(S, n))
        r = am.minimize_starts(w, acts, X0, Bounds(np.zeros(n), np.ones(n)), transform=transform)
        f0, g0 = km.value_and_input_grad(w, acts, X0, transform, True, np.float32)
        out[name + "/X0"] = X0
        out[name + "/f0"] = f0
        out[name + "/g0"] = g0
        for i, wi in enumerate(w):
            out[f"{name}/w{i}"] = wi
        for k, v in r.items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "lbfgsb_golden.npz"), **out)
    print("lbfgsb_golden.npz written")


def fit_golden():
    """Oracle fit trajectories (restated Keras semantics) frozen so that a later change to the
    oracle cannot silently move the target the CUDA kernel is held to."""
    from oracle import keras_mlp as km
    from helpers import NETS, synthetic_targets
    out = {}
    for name, N, E, l2 in (("cfg1_branin", 110, 30, 0.0), ("cfg5_plugin8", 200, 12, 1e-3)):
        dims, acts, _ = NETS[name]
        rs = np.random.RandomState(5)
        X = rs.uniform(size=(N, dims[0]))
        y = synthetic_targets(X)
        z = y < np.quantile(y, 0.25)
        perms = np.stack([rs.permutation(N) for _ in range(E)])
        w = km.init_weights(dims, 9)
        out[name + "/w0"] = np.concatenate([a.ravel() for a in w])
        hist, adam = km.fit(w, acts, X, z, E, 64, perms, l2=l2)
        out[name + "/X"] = X
        out[name + "/z"] = z
        out[name + "/perms"] = perms
        out[name + "/loss"] = hist
        out[name + "/w_final"] = np.concatenate([a.ravel() for a in w])
        out[name + "/l2"] = np.array(l2)
    np.savez_compressed(os.path.join(HERE, "fit_golden.npz"), **out)
    print("fit_golden.npz written")


def data_step_golden():
    """Quantile labels and duplicate flags produced by the reference's own bore.data.Record
    (importable: numpy only) -- the known answers of the device data step (data.cu)."""
    sys.path.insert(0, "/root/reference")
    from bore.data import Record                                 # noqa: E402
    out = dict(numpy_version=np.array(np.__version__),
               source=np.array("bore.data.Record of ltiao/bore v1.5.0, imported from /root/reference"))
    cases = []
    for ci, (seed, M, N, D, gamma) in enumerate(((0, 5, 1, 2, 0.25), (1, 4, 2, 3, 1 / 3), (2, 6, 37, 6, 0.25),
                                                 (3, 3, 64, 50, 0.5), (4, 4, 500, 6, 0.25), (5, 2, 2000, 3, 1 / 3),
                                                 (6, 3, 110, 2, 0.0), (7, 3, 129, 2, 1.0), (8, 3, 1000, 4, 0.9))):
        rs = np.random.RandomState(seed)
        X = rs.uniform(size=(M, N, D))
        y = rs.normal(size=(M, N))
        if N > 4:
            y[:, :3] = y[:, :1]                       # ties
            y[0] = np.round(y[0], 1)                  # many repeated values around the threshold
            y[-1, 1] = -0.0; y[-1, 2] = 0.0
        K = 7
        cand = rs.uniform(size=(M, K, D))
        cand[:, 0] = X[:, 0]                           # exact duplicate
        cand[:, 1] = X[:, -1] * (1 + 5e-6)             # within rtol
        cand[:, 2] = X[:, N // 2] + 1e-9               # within atol-ish
        cand[:, 3] = X[:, 0] + 1e-4                    # outside
        cand[:, 4, 0] = X[:, 0, 0]                     # one coordinate equal only
        z = np.zeros((M, N), bool); dup = np.zeros((M, K), bool)
        for m in range(M):
            rec = Record()
            for xi, yi in zip(X[m], y[m]):
                rec.append(x=xi, y=yi)
            _, z[m] = rec.load_classification_data(gamma)
            dup[m] = [rec.is_duplicate(c) for c in cand[m]]
        for k, v in dict(X=X, y=y, gamma=np.array(gamma), z=z, cand=cand, dup=dup).items():
            out[f"c{ci}/{k}"] = v
        cases.append(ci)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "data_step_golden.npz"), **out)
    print("data_step_golden.npz written")


def svgd_golden():
    """Kernel values and SVGD trajectories produced by the reference's own
    bore.optimizers.svgd (importable: numpy / scipy / sklearn only)."""
    sys.path.insert(0, "/root/reference")
    from bore.optimizers.svgd.base import SVGD, DistortionConstant, DistortionExpDecay   # noqa: E402
    from bore.optimizers.svgd.kernels import RadialBasis                                   # noqa: E402
    from oracle import svgd as osv
    from helpers import NETS, trained_weights
    out = dict(numpy_version=np.array(np.__version__),
               source=np.array("bore.optimizers.svgd of ltiao/bore v1.5.0, imported from /root/reference"))
    # (1) the kernel alone -- the grid of the reference's tests/test_optimizers.py:76-127
    kc = 0
    for n in (1, 2, 4, 16, 33):
        for D in (1, 2, 64):
            for ls in (None, 1e-3, 0.5, 2.0):
                X = np.random.RandomState(42 + kc).rand(n, D)
                if n >= 4:
                    X[1] = X[0]                       # coincident particles: zeros off the diagonal
                K, Kg = RadialBasis(length_scale=ls).value_and_grad(X)
                out[f"k{kc}/X"] = X; out[f"k{kc}/K"] = K; out[f"k{kc}/Kg"] = Kg
                out[f"k{kc}/ls"] = np.array(np.nan if ls is None else ls)
                kc += 1
    out["n_kernel_cases"] = np.array(kc)
    # (2) SVGD on the reference test's own target (tests/test_optimizers.py:130-160): analytic func
    mu = np.array([-0.6871, 0.8010])
    precision = np.array([[0.2260, 0.1652], [0.1652, 0.6779]])

    def func(x):
        d = x - mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, precision, d), (mu - x) @ precision
    tc = 0
    for n_iter, n, ls, lambd, seed in ((50, 4, None, None, 42), (50, 16, 0.5, None, 8888), (200, 16, None, 0.5, 42),
                                       (500, 64, None, None, 42), (200, 7, 1.0, 2.0, 8888), (100, 2, None, None, 1),
                                       (100, 1, None, None, 1)):
        rs = np.random.RandomState(seed)
        x_init = rs.randn(n, 2)
        bounds = [(-3., 3.), (-2., 4.)]
        snaps = []
        dist = DistortionConstant() if lambd is None else DistortionExpDecay(lambd=lambd)
        svgd = SVGD(kernel=RadialBasis(length_scale=ls), n_iter=n_iter, step_size=1e-2, alpha=.9, eps=1e-6,
                    tau=1.0, distortion=dist)
        x = svgd.optimize_from_init(func, x_init, bounds=bounds, callback=lambda v: snaps.append(v.copy()))
        out[f"t{tc}/x_init"] = x_init; out[f"t{tc}/x"] = x
        keep = [i for i in (0, 1, 9, 49, 99, 199, 499) if i < n_iter]
        out[f"t{tc}/snap_iters"] = np.array(keep); out[f"t{tc}/snaps"] = np.stack([snaps[i] for i in keep])
        out[f"t{tc}/cfg"] = np.array([n_iter, n, np.nan if ls is None else ls, np.nan if lambd is None else lambd])
        tc += 1
    out["n_analytic_cases"] = np.array(tc)
    # (3) SVGD with the model closure as func (bore/mixins.py:98,115): the reference's SVGD class driven
    #     by the oracle MLP's value-and-gradient (fp32), trained weights
    mc = 0
    for name, n, n_iter, ls, lambd in (("cfg2_hartmann6", 16, 100, None, None), ("cfg5_plugin8", 8, 60, None, 1.0),
                                       ("cfg3_ackley50", 32, 40, None, None), ("cfg1_branin", 64, 200, 0.2, None)):
        dims, acts, transform = NETS[name]
        w = trained_weights(dims, acts, seed=31)
        D = dims[0]
        rs = np.random.RandomState(5)
        bounds = [(0., 1.)] * D
        x_init = rs.uniform(size=(n, D))
        snaps = []
        dist = DistortionConstant() if lambd is None else DistortionExpDecay(lambd=lambd)
        svgd = SVGD(kernel=RadialBasis(length_scale=ls), n_iter=n_iter, step_size=1e-3, alpha=.9, eps=1e-6,
                    tau=1.0, distortion=dist)
        x = svgd.optimize_from_init(osv.make_func_max(w, acts, transform), x_init, bounds=bounds,
                                    callback=lambda v: snaps.append(v.copy()))
        out[f"m{mc}/name"] = np.array(name); out[f"m{mc}/x_init"] = x_init; out[f"m{mc}/x"] = x
        keep = [i for i in (0, 1, 9, 39, 59, 99, 199) if i < n_iter]
        out[f"m{mc}/snap_iters"] = np.array(keep); out[f"m{mc}/snaps"] = np.stack([snaps[i] for i in keep])
        out[f"m{mc}/cfg"] = np.array([n_iter, n, np.nan if ls is None else ls, np.nan if lambd is None else lambd])
        for i, wi in enumerate(w):
            out[f"m{mc}/w{i}"] = wi
        mc += 1
    out["n_model_cases"] = np.array(mc)
    np.savez_compressed(os.path.join(HERE, "svgd_golden.npz"), **out)
    print("svgd_golden.npz written")


def multi_fidelity_cases():
    """Seeded (x, y, budget) streams shaped like Hyperband brackets: every bracket starts some fresh
    configurations at one of the budgets and promotes the best third upwards; a few (x, b) pairs are
    re-evaluated (the reference overwrites the per-configuration value and appends to the rung)."""
    cases = []
    for seed, D, budgets, n0 in ((0, 3, (1 / 9, 1 / 3, 1.0), 9), (1, 5, (1 / 27, 1 / 9, 1 / 3, 1.0), 12),
                                 (2, 2, (0.25, 1.0), 6)):
        rs = np.random.RandomState(seed)
        stream = []
        for bracket in range(len(budgets) + 1):
            start = bracket % len(budgets)
            xs = [np.round(rs.uniform(size=D), 6) for _ in range(max(n0 // (start + 1), 2))]
            for b in budgets[start:]:
                ys = [float(np.sum((x - 0.4) ** 2) + rs.normal() * 0.1 / b) for x in xs]
                for x, y in zip(xs, ys):
                    stream.append((x, y, b))
                keep = max(len(xs) // 3, 1)
                xs = [xs[i] for i in np.argsort(ys)[:keep]]
        x, y, b = stream[3]
        stream.append((x, y + 0.5, b))  # the same configuration at the same budget again
        cases.append(dict(seed=seed, D=D, gamma=(0.25 if seed != 1 else 1 / 3), stream=stream))
    return cases


def multi_fidelity_golden():
    """MultiFidelityRecord (bore/data.py:51-261) run by the reference itself on the streams above."""
    sys.path.insert(0, "/root/reference")
    from bore.data import MultiFidelityRecord
    out = dict(numpy_version=np.array(np.__version__),
               source=np.array("ltiao/bore v1.5.0 bore.data.MultiFidelityRecord imported from /root/reference"))
    cases = multi_fidelity_cases()
    out["cases"] = np.arange(len(cases))
    for ci, c in enumerate(cases):
        rec = MultiFidelityRecord(gamma=c["gamma"])
        for x, y, b in c["stream"]:
            rec.append(x=x, y=y, b=b)
        k = f"c{ci}/"
        out[k + "budgets"] = np.array(rec.budgets())
        out[k + "rung_sizes"] = np.array(rec.rung_sizes())
        out[k + "size"] = np.array(rec.size())
        out[k + "num_features"] = np.array(rec.num_features())
        out[k + "thresholds"] = np.array(rec.thresholds())
        out[k + "highest_rung"] = np.array([-1 if rec.highest_rung(m) is None else rec.highest_rung(m)
                                            for m in (1, 3, 5, 8, 100)])
        for t in range(rec.num_rungs()):
            out[k + f"labels{t}"] = np.asarray(rec.binary_labels(t))
        for name, pad in (("m1", -1.0), ("tiny", 1e-9)):
            X, Y = rec.sequences(pad_value=pad, binary=True)
            out[k + f"X_{name}"], out[k + f"Y_{name}"] = X, Y
        Xc, Yc = rec.sequences(pad_value=-1.0, binary=False)
        out[k + "Y_values"] = Yc
        rs = np.random.RandomState(100 + ci)
        F = np.vstack([np.array(key) for key in rec._data])
        cand = np.vstack([F[::4] * (1 + 3e-6), F[1::4] + 2e-5, rs.uniform(size=(4, c["D"]))])
        out[k + "cand"] = cand
        out[k + "dup"] = np.array([rec.is_duplicate(x) for x in cand])
    np.savez_compressed(os.path.join(HERE, "multi_fidelity_golden.npz"), **out)
    print("multi_fidelity_golden.npz written")


if __name__ == "__main__":
    if "--svgd" in sys.argv:
        svgd_golden()
        sys.exit(0)
    if "--multi-fidelity" in sys.argv:
        multi_fidelity_golden()
        sys.exit(0)
    if "--data-step" in sys.argv:
        data_step_golden()
        sys.exit(0)
    host_golden()
    data_step_golden()
    multi_fidelity_golden()
    svgd_golden()
    lbfgsb_golden()
    fit_golden()
