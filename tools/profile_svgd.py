"""ncu target: SVGD batch argmax on the cfg-3 net (64 particles, 50-D)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bore_b200
from bore_b200.layers import Dense
from helpers import NETS
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_ackley50"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dims, acts, tr = NETS[name]
m = bore_b200.BatchMaximizableSequential()
for i, (u, a) in enumerate(zip(dims[1:], acts)):
    m.add(Dense(u, activation=a, input_dim=dims[0] if i == 0 else None))
m.compile(optimizer="adam", loss="binary_crossentropy")
x = m.argmax_batch(n, [(0., 1.)] * dims[0], n_iter=40, random_state=0)
torch.cuda.synchronize()
print("ok", x.shape)
