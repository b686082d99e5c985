"""Times the many-chain pass (development aid)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edward_b200 import engine
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=581012); ap.add_argument("--D", type=int, default=54)
ap.add_argument("--C", type=int, default=256); ap.add_argument("--T", type=int, default=4); ap.add_argument("--L", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
X = torch.randn(a.N, a.D, device=dev, generator=g)
wt = torch.randn(a.D, device=dev, generator=g) / a.D ** 0.5
y = (torch.rand(a.N, device=dev, generator=g) < torch.sigmoid(X @ wt)).to(torch.int32)
s = engine.GLMSampler(engine.GLMSpec(a.D), X, y, n_chains=a.C)
s.seed(1)
params = torch.zeros(a.T, a.C, a.D, device=dev)
s.run_chains(params, 0, a.T, 0.5 / a.N, a.L)
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); s.run_chains(params, 0, a.T, 0.5 / a.N, a.L); e1.record(); torch.cuda.synchronize()
  ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts)); steps = a.T * a.L
flops = 4.0 * a.N * a.D * a.C
print("impl=%s N=%d D=%d C=%d: %.3f ms/run, %.1f us/leapfrog step (all chains), %.0f chain-steps/s, %.1f algorithmic TFLOP/s (x3 executed), n_accept mean %.2f" % (
  os.environ.get("EDHMC_MC_IMPL", "tc"), a.N, a.D, a.C, ms, ms * 1e3 / steps, a.C * steps / ms * 1e3, flops * steps / ms / 1e9, s.read_chain_state()[0].mean()))
