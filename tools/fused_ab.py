"""A/B of the two argmax paths on one GPU: fused persistent kernel vs lock-step rounds.
Part 1: same starts, same weights -> do the per-start results agree (objective within 1e-4)?
Part 2: timing on the bench workloads (CUDA events around bore_lbfgsb_minimize)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from bore_b200 import _lib
from bore_b200.engine import NativeMLP
from helpers import NETS, trained_weights

lib = _lib.require_cuda()
out = {}

def run(net, X0d, transform, mode, reps=1):
    _lib.check(lib.bore_lbfgsb_set_mode({0: 2, 1: 1, 2: 0}[mode]))
    net._work = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); e0.record()
        r = net.lbfgsb_dev(X0d, 0.0, 1.0, transform=transform)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in r.items()}, best

if "parity" in sys.argv or len(sys.argv) == 1:
    for name in ["cfg1_branin", "cfg2_hartmann6", "cfg3_ackley50", "cfg5_plugin8", "tanh_exp", "no_hidden", "ref_test_linear"]:
        dims, acts, transform = NETS[name]
        w = trained_weights(dims, acts, seed=5)
        net = NativeMLP(dims, acts); net.set_weights(w)
        S = 2048
        X0d = torch.from_numpy(np.random.RandomState(7).uniform(size=(S, dims[0]))).cuda()
        a, ta = run(net, X0d, transform, 0)
        b, tb = run(net, X0d, transform, 1)
        agree = np.mean(np.abs(a["fun"] - b["fun"]) <= 1e-4)
        rec = dict(agree=float(agree), status_eq=float(np.mean(a["status"] == b["status"])),
                   nit=(float(a["nit"].mean()), float(b["nit"].mean())), nfev=(float(a["nfev"].mean()), float(b["nfev"].mean())),
                   mean_fun=(float(a["fun"].mean()), float(b["fun"].mean())), ms=(ta, tb), evals=(int(a["evals"]), int(b["evals"])),
                   status_hist=(np.bincount(a["status"], minlength=3).tolist(), np.bincount(b["status"], minlength=3).tolist()),
                   feasible=bool((a["x"] >= 0).all() and (a["x"] <= 1).all()))
        out["parity_" + name] = rec
        print(name, rec, flush=True)

if "time" in sys.argv or len(sys.argv) == 1:
    for cfg, S in [("cfg3", 65536), ("cfg2", 1024), ("cfg5", 65536), ("cfg3", 8192)]:
        wl = bench.WORKLOADS[cfg]
        X, z, perms = bench.make_problem(wl, 0)
        net = NativeMLP(wl["dims"], wl["acts"])
        net.set_weights(bench.glorot_init(wl["dims"], 0))
        net.fit(X, z, wl["epochs"], wl["batch"], perms)
        X0d = torch.from_numpy(np.random.RandomState(1).uniform(size=(S, wl["dims"][0]))).cuda()
        tname = {"identity": "identity", "sigmoid": "sigmoid"}[wl["transform"]]
        a, ta = run(net, X0d, tname, 0, reps=3)
        b, tb = run(net, X0d, tname, 1, reps=2)
        rec = dict(S=S, fused_ms=ta, rounds_ms=tb, evals=(int(a["evals"]), int(b["evals"])), rounds=int(b["rounds"]),
                   agree=float(np.mean(np.abs(a["fun"] - b["fun"]) <= 1e-4)),
                   best=(float(a["fun"].min()), float(b["fun"].min())))
        out[f"time_{cfg}_{S}"] = rec
        print(cfg, rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_ab.json"), "w"), indent=1)

if "exact" in sys.argv:
    # the two paths run the same L-BFGS-B core on bit-identical f, g: results must be EQUAL
    for name in ["cfg1_branin", "cfg2_hartmann6", "cfg3_ackley50", "cfg5_plugin8", "tanh_exp", "no_hidden", "ref_test_linear"]:
        dims, acts, transform = NETS[name]
        w = trained_weights(dims, acts, seed=5)
        net = NativeMLP(dims, acts); net.set_weights(w)
        X0d = torch.from_numpy(np.random.RandomState(7).uniform(size=(1024, dims[0]))).cuda()
        a, _ = run(net, X0d, transform, 0)
        b, _ = run(net, X0d, transform, 1)
        print(name, "equal:", {k: bool(np.array_equal(a[k], b[k])) for k in ("x", "fun", "nit", "nfev", "status", "task")}, flush=True)
