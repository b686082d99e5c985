import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "./oracle")
import numpy as np, torch, hmc_oracle as o
from edward_b200 import engine
rng = np.random.default_rng(0)
N, D, C = 64, 54, 128
X = rng.standard_normal((N, D)).astype(np.float32)
y = (rng.random(N) < 0.5).astype(np.int32)
s = engine.GLMSampler(engine.GLMSpec(D), X, y, n_chains=C)
theta = np.zeros((C, D), np.float32)
lp, g = s.logp_grad_chains(theta)
g = g.cpu().numpy()
g64 = o.grad_log_joint(X, y, theta[0], o.GLMSpec(D))
print("variant", os.environ.get("EDHMC_MC_MN", "0"))
print("got ", np.round(g[0][:10], 3)); print("want", np.round(g64[:10], 3))
# which (row,feature) combination does got match? try X^T r with permutations
r = (y - 0.5)
print("sum r*X[:, :10]   ", np.round(r @ X[:, :10], 3))
print("got nonzero count", np.count_nonzero(g[0]), "max", np.abs(g[0]).max())
