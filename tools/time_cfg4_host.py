"""Where the host time of a batched fit goes (cfg 4 sizes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from bore_b200.batched import BatchedMaximizableSequential
from bore_b200.layers import Dense
M, N, D = 4096, 500, 6
rs = np.random.RandomState(0)
X = rs.uniform(size=(M, N, D)); y = np.stack([bench.hartmann6(x) for x in X])
z = np.stack([row < np.quantile(row, 0.25) for row in y])
perm = np.stack([rs.permutation(N) for _ in range(125)])
b = BatchedMaximizableSequential([Dense(32, activation="relu", input_dim=6), Dense(32, activation="relu"),
                                  Dense(1, activation="sigmoid")], n_problems=M, seed=0)
b.compile()
net = b._net
def T(f, n=3):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, r
b.fit(X, z, batch_size=64, epochs=125, permutations=perm)
print("fit total ms", T(lambda: b.fit(X, z, batch_size=64, epochs=125, permutations=perm))[0])
print("fit(gamma) from raw y ms", T(lambda: b.fit(X, y, batch_size=64, epochs=125, permutations=perm, gamma=0.25))[0])
t, Xd = T(lambda: net.to_device(X.reshape(M * N, D), np.float32)); print("X to_device ms", t)
t, zd = T(lambda: net.to_device(z.reshape(-1).astype(np.float32), np.float32)); print("z to_device ms", t)
t, pd = T(lambda: net.to_device(np.ascontiguousarray(perm, np.int32), np.int32)); print("perm ms", t)
t, _ = T(lambda: net.fit_dev(Xd, zd, N, 64, 125, pd, model0=0, count=M, shared_data=False, shared_perm=True)); print("fit_dev ms", t)
Xi = rs.uniform(size=(M, 1024, D))
b.argmax([(0., 1.)] * D, num_starts=5, num_samples=1024, X_init=Xi)
print("argmax total ms", T(lambda: b.argmax([(0., 1.)] * D, num_starts=5, num_samples=1024, X_init=Xi))[0])
t, _ = T(lambda: net.to_device(Xi, np.float64)); print("X_init to_device ms", t)
