"""Small-shape pass over every kernel family, meant to run under compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from edward_b200 import _C, engine

dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)


def data(N, D):
  X = rng.randn(N, D).astype(np.float32)
  y = (rng.rand(N) < 0.5).astype(np.int32)
  return X, y


# (D <= 64 takes ring mode 2: re-laid tiles, shared-memory- and tensor-memory-resident tiles on the persistent plan)
for (N, D, bias) in [(1000, 54, False), (777, 5, True), (300, 130, False), (4096, 64, False), (513, 1000, False), (20011, 16, True),
                     (9000, 32, False), (70001, 54, False)]:
  X, y = data(N, D)
  for plan in (_C.PLAN_PERSISTENT, _C.PLAN_STEPWISE):
    s = engine.GLMSampler(engine.GLMSpec(D, has_bias=bias), X, y, device=dev, plan=plan)
    P = D + (1 if bias else 0)
    s.logp_grad(np.zeros(P, np.float32))
    params = torch.zeros(4, P, device=dev)
    s.seed(3)
    s.run(params, 0, 4, 0.01 / N, 3)
    assert torch.isfinite(params).all()
    s.close()
  print("single chain ok", N, D, bias, flush=True)

X, y = data(2000, 54)
s = engine.GLMSampler(engine.GLMSpec(54), X, y, device=dev)
params = torch.zeros(6, 54, device=dev)
for kind in ("sgld", "sghmc"):
  s.seed(5)
  s.sgmcmc_run(kind, params, 0, 6, 1e-4, batch_rows=256, velocity=torch.zeros(54, device=dev) if kind == "sghmc" else None)
print("sgmcmc ok", flush=True)
s.close()

for impl in ("simple", "tc"):
  os.environ["EDHMC_MC_IMPL"] = impl
  for (N, D) in [(1000, 54), (300, 64), (257, 8)]:
    X, y = data(N, D)
    s = engine.GLMSampler(engine.GLMSpec(D), X, y, device=dev, n_chains=128)
    s.seed(7)
    params = torch.zeros(3, 128, D, device=dev)
    s.run_chains(params, 0, 3, 0.01 / N, 2)
    assert torch.isfinite(params).all()
    s.close()
  print("chains ok", impl, flush=True)
os.environ["EDHMC_MC_IMPL"] = "tc"
for (N, D) in [(700, 200), (300, 1000)]:
  X, y = data(N, D)
  s = engine.GLMSampler(engine.GLMSpec(D), X, y, device=dev, n_chains=128)
  s.seed(7)
  params = torch.zeros(2, 128, D, device=dev)
  s.run_chains(params, 0, 2, 0.01 / N, 2)
  assert torch.isfinite(params).all()
  s.close()
print("wide chains ok", flush=True)
print("ALL OK")
