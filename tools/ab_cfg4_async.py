import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from bore_b200 import BatchedMaximizableSequential, Dense
M, N, D, E, B, K, P = 4096, 500, 6, 125, 64, 5, 1024
rs = np.random.RandomState(100)
X = rs.uniform(size=(M, N, D)); y = np.stack([bench.hartmann6(X[p]) for p in range(M)])
z = np.stack([y[p] < np.quantile(y[p], 0.25) for p in range(M)])
perms = np.stack([np.random.RandomState(7).permutation(N) for _ in range(E)])
X_init = rs.uniform(size=(M, P, D))
model = BatchedMaximizableSequential([Dense(32, activation="relu", input_dim=D), Dense(32, activation="relu"), Dense(1, activation="sigmoid")], n_problems=M, seed=0)
model.compile()
params = model._net.params_tensor(); w0d = params.clone()
def step(sync):
    t = [time.perf_counter()]
    params.copy_(w0d); model._net.reset_optimizer()
    h = model.fit(X, z, batch_size=B, epochs=E, permutations=perms)
    if sync: np.asarray(h)
    t.append(time.perf_counter())
    res = model.argmax([(0.0, 1.0)] * D, num_starts=K, num_samples=P, X_init=X_init)
    t.append(time.perf_counter())
    return [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])]
for sync in (1, 0, 1, 0):
    step(sync); torch.cuda.synchronize()
    t0 = time.perf_counter(); parts = [step(sync) for _ in range(3)]; torch.cuda.synchronize()
    print("sync" if sync else "async", round((time.perf_counter() - t0) / 3 * 1e3, 1), parts)
