"""Per-role timeline of CTA (0,0) of the many-chain tensor-core pass (edhmc_set_chain_debug; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edward_b200 import engine, _C
N, D, C = 581012, 54, 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
X = torch.randn(N, D, device=dev, generator=g)
y = (torch.rand(N, device=dev, generator=g) < 0.5).to(torch.int32)
s = engine.GLMSampler(engine.GLMSpec(D), X, y, n_chains=C)
s.seed(1)
params = torch.zeros(2, C, D, device=dev)
s.run_chains(params, 0, 2, 0.5 / N, 4)
buf = torch.zeros(64, 16, dtype=torch.int64, device=dev)
_C.check(s.lib.edhmc_set_chain_debug(s._h, buf.data_ptr()))
s.run_chains(params, 0, 1, 0.5 / N, 2)   # the last pass of the call leaves its stamps (a gradient-only + an lp pass)
torch.cuda.synchronize()
_C.check(s.lib.edhmc_set_chain_debug(s._h, None))
t = buf.cpu().numpy()
names = ["tma", "m1rdy", "m1iss", "m2rdy", "m2iss", "landed", "transp", "Srdy", "Sld", "Rst"]
b0 = t[16, 0]
for i in range(16, 30):
  print(i, " ".join("%s=%d" % (n, t[i, k] - b0) for k, n in enumerate(names)))
d = np.diff(t[8:60, 9])
print("period per tile (R stored): median %.0f  min %d max %d" % (np.median(d), d.min(), d.max()))
print("epilogue S ready -> R stored: median %.0f" % np.median(t[8:60, 9] - t[8:60, 7]))
print("epilogue waits for S (transposed -> S ready): median %.0f" % np.median(t[8:60, 7] - t[8:60, 6]))
print("epilogue landed wait+transposes (prev R stored -> transposed): median %.0f" % np.median(t[9:60, 6] - t[8:59, 9]))
print("MMA1 issue (ready -> issued): median %.0f; MMA2: %.0f" % (np.median(t[8:60, 2] - t[8:60, 1]), np.median(t[8:60, 4] - t[8:60, 3])))
print("S ready after MMA1 issued: median %.0f" % np.median(t[8:60, 7] - t[8:60, 2]))
print("MMA2 ready after R stored: median %.0f" % np.median(t[8:60, 3] - t[8:60, 9]))
print("MMA1(i) ready after MMA2(i-2) issued: median %.0f" % np.median(t[10:60, 1] - t[8:58, 4]))
