"""Short single-GPU target for ncu: one cfg-3 BO iteration (fit + 65,536-start argmax)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from bore_b200.engine import NativeMLP

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
S = int(sys.argv[2]) if len(sys.argv) > 2 else wl["starts"]
X, z, perms = bench.make_problem(wl, 0)
net = NativeMLP(wl["dims"], wl["acts"])
net.set_weights(bench.glorot_init(wl["dims"], 0))
net.fit(X, z, wl["epochs"], wl["batch"], perms)
D = wl["dims"][0]
X0 = torch.from_numpy(np.random.RandomState(1).uniform(size=(S, D))).cuda()
r = net.lbfgsb_dev(X0, 0.0, 1.0, transform=wl["transform"])
torch.cuda.synchronize()
print("rounds", r["rounds"], "evals", r["evals"])
