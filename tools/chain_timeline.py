"""Prints the per-role clock64 timeline of CTA (0,0) of k_mc_pass_tc2 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edward_b200 import engine, _C
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
N, D, Cn = 581012, 54, 256
X = torch.randn(N, D, device=dev, generator=g)
y = (torch.rand(N, device=dev, generator=g) < 0.5).to(torch.int32)
s = engine.GLMSampler(engine.GLMSpec(D), X, y, n_chains=Cn)
dbg = torch.zeros(64, 16, dtype=torch.int64, device=dev)
theta = torch.zeros(Cn, D, device=dev)
s.logp_grad_chains(theta)
_C.check(s.lib.edhmc_set_chain_debug(s._h, dbg.data_ptr()))
params = torch.zeros(2, Cn, D, device=dev)
s.run_chains(params, 0, 1, 1e-6, 2)   # last pass is a gradient+logp pass; overwritten by each pass
torch.cuda.synchronize()
t = dbg.cpu().numpy().astype(np.int64)
t0 = t[t > 0].min()
names = ["ctl:b_ready", "ctl:mma1_iss", "m2:r_ready", "m2:issued", "bld:start", "bld:raw_ok", "bld:built", "bld:arrived",
         "epi:s_ready", "epi:ld_done", "epi:computed", "epi:st_done", "epi:arrived"]
print("cycles since first event; tile rows")
print("tile " + " ".join("%12s" % n for n in names))
for i in range(2, 14):
  print("%4d " % i + " ".join("%12d" % (t[i, k] - t0 if t[i, k] else -1) for k in range(13)))
