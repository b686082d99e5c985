"""compute-sanitizer target (racecheck): ring mode 2 on the persistent plan, one shape whose resident tiles live in tensor
memory and one whose tiles live in shared memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from edward_b200 import _C, engine
rng = np.random.RandomState(0)
for (N, D, bias) in [(3000, 54, False), (5000, 16, True)]:
  X = rng.randn(N, D).astype(np.float32); y = (rng.rand(N) < 0.5).astype(np.int32)
  s = engine.GLMSampler(engine.GLMSpec(D, has_bias=bias), X, y, plan=_C.PLAN_PERSISTENT)
  P = D + int(bias)
  params = torch.zeros(3, P, device="cuda"); s.seed(3); s.run(params, 0, 3, 0.01 / N, 2)
  assert torch.isfinite(params).all(); print("ok", N, D, s.plan_info(), flush=True); s.close()
