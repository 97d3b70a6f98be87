"""ncu target: the cfg-3 fit alone (one model, Dense64x3, N=2000) for a few epochs, in the given fit mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from bore_b200.engine import NativeMLP

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "cfg3"]
X, z, perms = bench.make_problem(wl, 0)
net = NativeMLP(wl["dims"], wl["acts"])
net.set_fit_mode(mode)
net.set_weights(bench.glorot_init(wl["dims"], 0))
h = net.fit(X, z, epochs, wl["batch"], perms[:epochs])
torch.cuda.synchronize()
print("loss", h[0], h[-1])
