"""ncu target: the batched fit of cfg 4 alone (M Hartmann-6 problems, Dense32x2, N=500, 125 epochs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import time
import numpy as np
import torch
import bench
from bore_b200.batched import BatchedMaximizableSequential
from bore_b200.layers import Dense

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = bench.WORKLOADS["cfg2"]
rs = np.random.RandomState(0)
N, D = 500, 6
X = rs.uniform(size=(M, N, D))
y = np.stack([bench.hartmann6(x) for x in X])
z = np.stack([row < np.quantile(row, 0.25) for row in y])
perm = np.stack([rs.permutation(N) for _ in range(125)])
b = BatchedMaximizableSequential([Dense(32, activation="relu", input_dim=6), Dense(32, activation="relu"),
                                  Dense(1, activation="sigmoid")], n_problems=M, seed=0)
b.compile()
for r in range(reps):
    torch.cuda.synchronize(); t = time.time()
    loss = b.fit(X, z, batch_size=64, epochs=125, permutations=perm)
    torch.cuda.synchronize(); print("fit ms", (time.time() - t) * 1e3, "loss", loss[:, 0].mean(), loss[:, -1].mean())
