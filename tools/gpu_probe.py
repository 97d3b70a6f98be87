"""Quick on-GPU timing probe (not the bench): FFMA peak, K2 throughput, K3 end-to-end."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from bore_b200.engine import NativeMLP, ffma_peak_tflops
from helpers import NETS, trained_weights

print("ffma peak TFLOP/s:", ffma_peak_tflops(0))
for name, S in [("cfg3_ackley50", 65536), ("cfg2_hartmann6", 65536), ("cfg5_plugin8", 65536)]:
    dims, acts, transform = NETS[name]
    w = trained_weights(dims, acts, seed=0, N=500, epochs=40)
    net = NativeMLP(dims, acts)
    net.set_weights(w)
    W = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    X = torch.rand(S, dims[0], device="cuda", dtype=torch.float32)
    f = torch.empty(S, device="cuda"); g = torch.empty(S, dims[0], device="cuda")
    for _ in range(3):
        net.value_and_grad_dev(X, transform, True, f, g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net.value_and_grad_dev(X, transform, True, f, g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: K2 S={S} {ms*1e3:.1f} us  {S/ms/1e3:.2f} Mevals/s  {4*W*S/ms/1e9:.2f} TFLOP/s")
    X0 = torch.rand(S, dims[0], device="cuda", dtype=torch.float64)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = net.lbfgsb_dev(X0, 0.0, 1.0, transform=transform)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    st = r["status"].cpu().numpy()
    print(f"{name}: K3 S={S} {dt*1e3:.1f} ms rounds={r['rounds']} evals={r['evals']} "
          f"evals/s={r['evals']/dt/1e6:.2f}M  mlp TFLOP/s={4*W*r['evals']/dt/1e12:.2f} "
          f"nit mean={r['nit'].float().mean().item():.1f} max={r['nit'].max().item()} "
          f"nfev mean={r['nfev'].float().mean().item():.1f} status={np.bincount(st, minlength=3)}")
