"""Per-pass timeline of the persistent plan on cfg 2 under the current EDHMC_LEADER setting (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from edward_b200 import engine
N, D, T, L = int(os.environ.get("N", 581012)), 54, 100, 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
X = torch.randn(N, D, device=dev, generator=g)
y = (torch.rand(N, device=dev, generator=g) < 0.5).to(torch.int32)
s = engine.GLMSampler(engine.GLMSpec(D), X, y)
s.seed(1)
params = torch.zeros(T, D, device=dev)
s.run(params, 0, T, 0.5 / N, L)
tl = bench.timeline_of(torch, s, lambda: s.run(params, 0, T, 0.5 / N, L))
print(json.dumps(tl))
if os.environ.get("RAW"):
  t = s.set_timeline(48); s.run(params, 0, T, 0.5 / N, L); torch.cuda.synchronize()
  a = t.cpu().numpy()
  p = 20
  base = a[p, 0, 0]
  names = {0: "start", 16: "loop start", 17: "loop done", 18: "w0 reduced", 19: "all reduced", 1: "cta sums", 2: "published", 20: "grp0 summed",
           21: "groups seen", 3: "totals", 4: "tl4", 22: "gradient", 23: "integrated", 24: "sent", 5: "end", 25: "theta seen"}
  for pp in (21, 22, 23, 30):
    for c in (0, 1, 12, a.shape[1] - 1):
      b = a[pp, c, 0]
      ev = sorted((int(a[pp, c, k] - b), n) for k, n in names.items() if a[pp, c, k] != 0)
      ev.append((int(a[pp + 1, c, 0] - b), "next start"))
      if a[pp + 1, c, 25]:
        ev.append((int(a[pp + 1, c, 25] - b), "next theta seen"))
      print("pass", pp, "cta", c, " ".join("%s=%d" % (n, v) for v, n in sorted(ev)))
  print("leader pass", p, [int(v - base) for v in a[p, 0, :6]], "next start", int(a[p + 1, 0, 0] - base))
  for c in (1, a.shape[1] - 1):
    print("cta", c, [int(v - base) for v in a[p, c, :3]], "next start", int(a[p + 1, c, 0] - base))
  ends = a[p, :, 1] - base
  print("pass-end spread: min %d med %d max %d" % (ends.min(), np.median(ends), ends.max()))
