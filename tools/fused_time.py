"""Timing of bore_lbfgsb_minimize on a bench workload: python tools/fused_time.py cfg3 65536 [mode] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from bore_b200 import _lib
from bore_b200.engine import NativeMLP
lib = _lib.require_cuda()
cfg = sys.argv[1]; S = int(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
wl = bench.WORKLOADS[cfg]
X, z, perms = bench.make_problem(wl, 0)
net = NativeMLP(wl["dims"], wl["acts"])
net.set_weights(bench.glorot_init(wl["dims"], 0))
net.fit(X, z, wl["epochs"], wl["batch"], perms)
X0d = torch.from_numpy(np.random.RandomState(1).uniform(size=(S, wl["dims"][0]))).cuda()
_lib.check(lib.bore_lbfgsb_set_mode({0: 2, 1: 1, 2: 0}[mode]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(reps):
    torch.cuda.synchronize(); e0.record()
    r = net.lbfgsb_dev(X0d, 0.0, 1.0, transform=wl["transform"])
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(cfg, S, "mode", mode, "env", {k: v for k, v in os.environ.items() if k.startswith("BORE_")}, "ms", [round(t, 2) for t in ts],
      "evals", r["evals"], "nfev_sum", int(r["nfev"].sum()), "nit", float(r["nit"].float().mean()), "best", float(r["fun"].min()),
      "status", torch.bincount(r["status"], minlength=3).tolist(), "mean_fun", float(r["fun"].mean()))
if len(sys.argv) > 5:
    import numpy as np
    np.save(sys.argv[5], r["fun"].cpu().numpy())
