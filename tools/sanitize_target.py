"""Tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
python tools/sanitize_target.py [k1|k1c|k1u|k1t|k2|k3|k3f|k3multi|svgd|data|lstm|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from bore_b200 import _lib
from bore_b200.engine import NativeMLP
lib = _lib.require_cuda()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
rs = np.random.RandomState(0)


def glorot(dims, seed):
    r = np.random.RandomState(seed)
    ws = []
    for fi, fo in zip(dims[:-1], dims[1:]):
        lim = np.sqrt(6.0 / (fi + fo))
        ws += [r.uniform(-lim, lim, size=(fi, fo)).astype(np.float32), np.zeros(fo, np.float32)]
    return ws


def data(N, D, E):
    X = rs.uniform(size=(N, D)); y = np.sum((X - 0.4) ** 2, axis=1)
    z = y < np.quantile(y, 0.25)
    return X, z, np.stack([rs.permutation(N) for _ in range(E)]).astype(np.int32)


dims, acts = [6, 32, 32, 1], ["relu", "relu", "sigmoid"]
if what in ("k1", "all"):      # one CTA per model
    net = NativeMLP(dims, acts, n_models=3)
    for m in range(3): net.set_weights(glorot(dims, m), model=m)
    net.set_fit_mode(1)
    X, z, perms = data(150, 6, 2)
    net.fit_dev(net.to_device(X, np.float32), net.to_device(z.astype(np.float32), np.float32), 150, 64, 2,
                net.to_device(perms, np.int32), model0=0, count=3)
    torch.cuda.synchronize(); print("k1 ok")
if what in ("k1c", "all"):     # one 8-CTA cluster per model
    d3, a3 = [50, 64, 64, 64, 1], ["relu", "relu", "relu", "sigmoid"]
    net = NativeMLP(d3, a3); net.set_weights(glorot(d3, 0)); net.set_fit_mode(2)
    X, z, perms = data(200, 50, 2)
    print("k1c loss", net.fit(X, z, 2, 64, perms))
if what in ("k1u", "all"):     # unit-split cluster kernel (fit mode 4): compile-time shape, run-time shape, l2, 2 models
    d3, a3 = [50, 64, 64, 64, 1], ["relu", "relu", "relu", "sigmoid"]
    net = NativeMLP(d3, a3); net.set_weights(glorot(d3, 0)); net.set_fit_mode(4)
    X, z, perms = data(200, 50, 2)
    print("k1u loss", net.fit(X, z, 2, 64, perms))
    d5, a5 = [7, 24, 40, 1], ["tanh", "elu", "linear"]  # run-time-shape build, ragged widths, l2 (extra cluster barrier)
    net = NativeMLP(d5, a5, n_models=2)
    for m in range(2): net.set_weights(glorot(d5, m), model=m)
    net.set_fit_mode(4)
    X, z, perms = data(100, 7, 2)
    print("k1u generic loss", net.fit(X, z, 2, 32, perms, l2=1e-3))
if what in ("k1t", "all"):     # tensor-pipe kernel (fit mode 3): 16-warp CTA and the 4-warp form
    d3, a3 = [50, 64, 64, 64, 1], ["relu", "relu", "relu", "sigmoid"]
    net = NativeMLP(d3, a3); net.set_weights(glorot(d3, 0)); net.set_fit_mode(3)
    X, z, perms = data(150, 50, 2)
    print("k1t loss", net.fit(X, z, 2, 64, perms))
    os.environ["BORE_FIT_MMA_WARPS"] = "4"
    net = NativeMLP(dims, acts, n_models=2)
    for m in range(2): net.set_weights(glorot(dims, m), model=m)
    net.set_fit_mode(3)
    X, z, perms = data(150, 6, 2)
    net.fit_dev(net.to_device(X, np.float32), net.to_device(z.astype(np.float32), np.float32), 150, 64, 2,
                net.to_device(perms, np.int32), model0=0, count=2)
    torch.cuda.synchronize(); print("k1t ok")
if what in ("k2", "all"):
    net = NativeMLP(dims, acts); net.set_weights(glorot(dims, 0))
    f, g = net.value_and_grad(rs.uniform(size=(100, 6)), "identity", True)
    p = net.predict(rs.uniform(size=(77, 6))); print("k2 ok", f.sum(), p.sum())
if what in ("k3", "k3f", "all"):
    for mode, tag in ((1, "k3"), (2, "k3f")):
        if what not in (tag, "all"): continue
        _lib.check(lib.bore_lbfgsb_set_mode(mode))
        net = NativeMLP(dims, acts); net.set_weights(glorot(dims, 1))
        r = net.lbfgsb(rs.uniform(size=(40, 6)), 0.0, 1.0)
        print(tag, "ok rounds", r["rounds"], "nit", r["nit"].mean())
    _lib.check(lib.bore_lbfgsb_set_mode(0))
if what in ("k3multi", "all"):
    net = NativeMLP(dims, acts, n_models=4)
    for m in range(4): net.set_weights(glorot(dims, 10 + m), model=m)
    for mode in (2, 1):
        _lib.check(lib.bore_lbfgsb_set_mode(mode)); net._work = None
        r = net.lbfgsb_multi_dev(torch.from_numpy(rs.uniform(size=(4, 5, 6))).cuda(), 0.0, 1.0)
        torch.cuda.synchronize()
    _lib.check(lib.bore_lbfgsb_set_mode(0)); print("k3multi ok")
if what in ("svgd", "all"):
    net = NativeMLP(dims, acts); net.set_weights(glorot(dims, 2))
    x = net.svgd_maximize(rs.uniform(size=(8, 6)), "identity", np.zeros(6), np.ones(6), 5, float("nan"), 1e-3, .9, 1e-6, 1.0, 0.0, 1.0)
    print("svgd ok", x.shape)
if what in ("data", "all"):
    net = NativeMLP(dims, acts)
    z = net.quantile_labels_dev(torch.from_numpy(rs.normal(size=(3, 101))).cuda(), 1 / 3)
    k = net.keep_unique_dev(torch.from_numpy(rs.uniform(size=(2, 5, 6))).cuda(), torch.from_numpy(rs.uniform(size=(2, 9, 6))).cuda())
    torch.cuda.synchronize(); print("data ok")
if what in ("lstm", "all"):    # K7: masked forward, one-to-one value + gradient, fit, evaluate, argmax loop
    from bore_b200 import ops
    from bore_b200.layers import BinaryCrossentropy
    from bore_b200.models import StackedRecurrentFactory
    fac = StackedRecurrentFactory(5, 1, num_layers=2, num_units=16, layer_kws=dict(activation="elu"), seed=0)
    Xs = rs.uniform(size=(40, 3, 5)); Ys = (rs.uniform(size=(40, 3, 1)) < 0.4).astype(np.float64)
    Xs[::3, 2] = -1.0; Ys[::3, 2] = -1.0; Xs[1::5, 0] = -1.0; Ys[1::5, 0] = -1.0
    net = fac.build_many_to_many(mask_value=-1.0)
    net.compile(optimizer="adam", loss=BinaryCrossentropy(from_logits=True), metrics=["accuracy"])
    h = net.fit(Xs, Ys, epochs=2, batch_size=16, verbose=0).history["loss"]
    ev = net.evaluate(Xs, Ys); p = net.predict(Xs)
    one = fac.build_one_to_one(3, transform=ops.sigmoid)
    r = one.argmax([(0.0, 1.0)] * 5, num_starts=6, num_samples=32, print_fn=None, random_state=np.random.RandomState(0))
    print("lstm ok", h, ev, p.shape, None if r is None else r.fun)
