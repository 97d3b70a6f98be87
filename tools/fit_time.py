"""Times the cfg-3 fit (one model, 992 Adam steps) and a cfg-4 style batched fit in every fit mode."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from bore_b200.engine import NativeMLP


def glorot(dims, seed):
    rs = np.random.RandomState(seed)
    ws = []
    for fi, fo in zip(dims[:-1], dims[1:]):
        lim = np.sqrt(6.0 / (fi + fo))
        ws += [rs.uniform(-lim, lim, size=(fi, fo)).astype(np.float32), np.zeros(fo, np.float32)]
    return ws


def time_fit(dims, acts, M, N, E, mode, reps=3):
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(N, dims[0])).astype(np.float32)
    z = (np.sum((X - 0.4) ** 2, axis=1) < np.quantile(np.sum((X - 0.4) ** 2, axis=1), 0.25)).astype(np.float32)
    perm = np.stack([rs.permutation(N) for _ in range(E)]).astype(np.int32)
    net = NativeMLP(dims, acts, n_models=M)
    net.set_fit_mode(mode)
    w = glorot(dims, 0)
    Xd, zd, pd = net.to_device(X, np.float32), net.to_device(z, np.float32), net.to_device(perm, np.int32)
    out = []
    for r in range(reps):
        for i in range(M):
            net.set_weights(w, model=i)
        net.reset_optimizer()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = net.fit_dev(Xd, zd, N, 64, E, pd, model0=0, count=M, shared_data=True, shared_perm=True)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    l = loss.cpu().numpy()
    return min(out), float(l[0, 0]), float(l[0, -1])


if __name__ == "__main__":
    cfg3 = ([50, 64, 64, 64, 1], ["relu", "relu", "relu", "sigmoid"])
    cfg2 = ([6, 32, 32, 1], ["relu", "relu", "sigmoid"])
    cfg5 = ([8, 32, 32, 32, 1], ["elu", "elu", "elu", "linear"])
    cfg1 = ([2, 16, 16, 1], ["relu", "relu", "sigmoid"])
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "single"):
        for mode in (2, 3, 4):
            print("cfg3 1 model, 992 steps, mode", mode, time_fit(*cfg3, 1, 2000, 31, mode), flush=True)
            print("cfg5 1 model, 1000 steps, mode", mode, time_fit(*cfg5, 1, 500, 125, mode), flush=True)
    if which == "trace":
        print("cfg3 mode 4", time_fit(*cfg3, 1, 2000, 31, 4, reps=1), flush=True)
    if which == "unit":
        for mode in (2, 4):
            print("cfg3 1 model, 992 steps, mode", mode, time_fit(*cfg3, 1, 2000, 31, mode), flush=True)
            print("cfg5 1 model, 1000 steps, mode", mode, time_fit(*cfg5, 1, 500, 125, mode), flush=True)
            print("cfg2 1 model, 1000 steps, mode", mode, time_fit(*cfg2, 1, 500, 125, mode), flush=True)
            print("cfg5 net 18 models, 1000 steps, mode", mode, time_fit(*cfg5, 18, 500, 125, mode), flush=True)
            print("cfg5 net, 40 observations (small batch), 1000 steps, mode", mode, time_fit(*cfg5, 1, 40, 1000, mode), flush=True)
    if which in ("all", "many"):
        for M in (148, 512, 1024, 4096):
            print("cfg4 net", M, "models, 1000 steps, mode 1", time_fit(*cfg2, M, 500, 125, 1, reps=2), flush=True)
        print("cfg5 net 512 models, mode 1", time_fit(*cfg5, 512, 500, 125, 1, reps=2), flush=True)
        print("cfg1 net 1024 models, 400 steps, mode 1", time_fit(*cfg1, 1024, 110, 200, 1, reps=2), flush=True)
