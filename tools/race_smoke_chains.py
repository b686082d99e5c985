"""compute-sanitizer target: one small launch of the tcgen05 many-chain pass (for comparing tool reports with race_smoke.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edward_b200 import engine
rng = np.random.RandomState(0)
X = rng.randn(1000, 54).astype(np.float32); y = (rng.rand(1000) < 0.5).astype(np.int32)
s = engine.GLMSampler(engine.GLMSpec(54), X, y, n_chains=128)
lp, g = s.logp_grad_chains(np.zeros((128, 54), np.float32))
assert torch.isfinite(lp).all(); print("chains ok", flush=True); s.close()
