"""Quick device-side timing of edhmc_run on synthetic data (development aid; bench.py is the contract)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from edward_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=581012)
ap.add_argument("--D", type=int, default=54)
ap.add_argument("--T", type=int, default=100)
ap.add_argument("--L", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--plan", type=int, default=0)
ap.add_argument("--eps", type=float, default=None)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
X = torch.randn(a.N, a.D, device=dev, generator=g)
wt = torch.randn(a.D, device=dev, generator=g) / a.D ** 0.5
y = (torch.rand(a.N, device=dev, generator=g) < torch.sigmoid(X @ wt)).to(torch.int32)
s = engine.GLMSampler(engine.GLMSpec(a.D), X, y, plan=a.plan)
s.seed(1)
eps = a.eps if a.eps is not None else 0.5 / a.N
params = torch.zeros(a.T, a.D, device=dev)
s.run(params, 0, a.T, eps, a.L)
torch.cuda.synchronize()
times = []
for _ in range(a.reps):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); s.run(params, 0, a.T, eps, a.L); e1.record(); torch.cuda.synchronize()
  times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
steps = a.T * a.L
bytes_step = 4.0 * a.N * a.D + 4.0 * a.N
print("plan", s.plan_info())
print("N=%d D=%d T=%d L=%d: %.3f ms/run  %.1f leapfrog steps/s  %.2f us/step  %.1f GB/s algorithmic (%.1f%% of 6550)  n_accept=%d" % (
    a.N, a.D, a.T, a.L, ms, steps / ms * 1e3, ms * 1e3 / steps, bytes_step * steps / ms / 1e6,
    100 * bytes_step * steps / ms / 1e6 / 6550.4, s.read_state()[0]))
print("times", times)
