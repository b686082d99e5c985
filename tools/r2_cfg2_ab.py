"""Same-box A/B of the cfg 2 data pass: ring mode 0 (one TMA ring per warp) vs ring mode 1 (one ring per CTA) at several
lane / warp geometries, each checked against the ring-0 results, plus the per-pass timeline of the persistent kernel.
Development aid (bench.py is the contract). Usage: python tools/r2_cfg2_ab.py [--N ..] [--D ..] [--variants a,b,..]"""
import argparse, json, os, sys
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
ap.add_argument("--variants", default="r0,r1g1,r1g2,r1g4,r1g1n12")
ap.add_argument("--timeline", default="r0,r1g1")
ap.add_argument("--out", default=None)
ap.add_argument("--one", default=None, help="(internal) run this single variant in this process")
ap.add_argument("--per-variant-timeout", type=int, default=90)
a = ap.parse_args()

if a.one is None:
  # one subprocess per variant with its own timeout: a variant that deadlocks costs 90 s, not the whole GPU call
  import subprocess
  for name in a.variants.split(","):
    cmd = ["timeout", str(a.per_variant_timeout), sys.executable, os.path.abspath(__file__), "--one", name, "--N", str(a.N),
           "--D", str(a.D), "--T", str(a.T), "--L", str(a.L), "--reps", str(a.reps), "--timeline", a.timeline]
    r = subprocess.run(cmd, capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    if r.returncode != 0:
      print(name, "FAILED rc=%d %s" % (r.returncode, r.stderr[-300:].replace("\n", " | ")))
    sys.stdout.flush()
  sys.exit(0)
a.variants = a.one

VARIANTS = {
    "r0": {"EDHMC_RING": "0"},
    "r1": {"EDHMC_RING": "1"},
    "r1g1": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1"},
    "r1g1n12": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1", "EDHMC_FORCE_NW": "12"},
    "r1g2": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "2"},
    "r1g2n8": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "2", "EDHMC_FORCE_NW": "8"},
    "r1g4": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "4"},
    "r1g2j2": {"EDHMC_RING": "1", "EDHMC_FORCE_G": "2", "EDHMC_FORCE_J": "2"},
}
for fr in (50, 60, 70, 75, 80, 85, 90):
  VARIANTS["r1g1f%d" % fr] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1", "EDHMC_L2_HINT": "2", "EDHMC_L2_FRAC": "%.2f" % (fr / 100.0)}
  VARIANTS["r0f%d" % fr] = {"EDHMC_RING": "0", "EDHMC_L2_HINT": "2", "EDHMC_L2_FRAC": "%.2f" % (fr / 100.0)}
for w in (1, 2, 4, 8):
  VARIANTS["r1g1w%d" % w] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1", "EDHMC_FORCE_WPG": str(w)}
  VARIANTS["r1g2w%d" % w] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "2", "EDHMC_FORCE_WPG": str(w)}
  VARIANTS["r1g2n8w%d" % w] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "2", "EDHMC_FORCE_NW": "8", "EDHMC_FORCE_WPG": str(w)}
VARIANTS["r1g1nz"] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1", "EDHMC_ZIGZAG": "0"}
VARIANTS["r1g1f80nz"] = {"EDHMC_RING": "1", "EDHMC_FORCE_G": "1", "EDHMC_ZIGZAG": "0", "EDHMC_L2_HINT": "2", "EDHMC_L2_FRAC": "0.80"}
KEYS = ["EDHMC_RING", "EDHMC_FORCE_G", "EDHMC_FORCE_NW", "EDHMC_FORCE_J", "EDHMC_FORCE_WPG", "EDHMC_L2_HINT", "EDHMC_L2_FRAC", "EDHMC_ZIGZAG"]

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
X = torch.randn(a.N, a.D, device=dev, generator=g)
wt = torch.randn(a.D, device=dev, generator=g) / a.D ** 0.5
y = (torch.rand(a.N, device=dev, generator=g) < torch.sigmoid(X @ wt)).to(torch.int32)
eps = 0.5 / a.N
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
theta = (0.1 * torch.randn(a.D, device=dev, generator=g)).contiguous()

ref = None
out = {}
for name in a.variants.split(","):
  for k in KEYS:
    os.environ.pop(k, None)
  os.environ.update(VARIANTS[name])
  try:
    s = engine.GLMSampler(engine.GLMSpec(a.D), X, y)
  except Exception as e:  # noqa: BLE001
    print(name, "FAILED to create:", e)
    continue
  s.seed(1)
  info = s.plan_info()
  lp, gr = s.logp_grad(theta)
  params = torch.zeros(a.T, a.D, device=dev)
  s.run(params, 0, a.T, eps, a.L)
  torch.cuda.synchronize()
  times = []
  for _ in range(a.reps):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.run(params, 0, a.T, eps, a.L); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
  ms = float(np.median(times))
  nacc = s.read_state()[0]
  res = {"plan": info, "ms": ms, "us_per_step": ms * 1e3 / (a.T * a.L), "n_accept": nacc, "times": times,
         "gbs": (4.0 * a.N * a.D + 4.0 * a.N) * a.T * a.L / ms / 1e6}
  cur = (float(lp.cpu()[0]), gr.cpu().numpy().astype(np.float64), params.cpu().numpy().astype(np.float64))
  res["logp"] = cur[0]
  res["grad_l2"] = float(np.sqrt((cur[1] ** 2).sum()))
  res["params_l2"] = float(np.sqrt((cur[2] ** 2).sum()))
  if ref is None:
    ref = cur
  else:
    res["logp_rel"] = abs(cur[0] - ref[0]) / abs(ref[0])
    res["grad_rel"] = float(np.abs(cur[1] - ref[1]).max() / np.abs(ref[1]).max())
    res["params_rel"] = float(np.abs(cur[2] - ref[2]).max() / max(np.abs(ref[2]).max(), 1e-30))
  if name in a.timeline.split(","):
    npass = 40
    tl = s.set_timeline(npass)
    s.run(params, 0, a.T, eps, a.L)
    torch.cuda.synchronize()
    t = tl.cpu().numpy().astype(np.int64)[8:npass]  # [pass, cta, 8], steady state
    s.set_timeline(0)
    d = {}
    d["pass_compute"] = np.median(t[:, :, 1] - t[:, :, 0])
    d["publish"] = np.median(t[:, :, 2] - t[:, :, 1])
    d["barrier_wait_med"] = np.median(t[:, :, 3] - t[:, :, 2])
    d["barrier_wait_min"] = np.median((t[:, :, 3] - t[:, :, 2]).min(axis=1))
    d["barrier_wait_max"] = np.median((t[:, :, 3] - t[:, :, 2]).max(axis=1))
    d["read_partials"] = np.median(t[:, :, 4] - t[:, :, 3])
    d["integrator"] = np.median(t[:, :, 5] - t[:, :, 4])
    d["step_cycles"] = np.median(t[1:, :, 0] - t[:-1, :, 0])
    d["serial_after_compute"] = np.median(t[:, :, 5] - t[:, :, 1])
    gt_start = t[:, :, 6]
    d["pass_start_spread_ns"] = float(np.median(gt_start.max(axis=1) - gt_start.min(axis=1)))
    d["step_ns_globaltimer"] = float(np.median(gt_start[1:, 0] - gt_start[:-1, 0]))
    d["tile_wait_per_warp"] = np.median(t[:, :, 8:8 + min(8, info["warps_per_cta"])])
    # arrival spread at the barrier: when (globaltimer-aligned) does each CTA finish its pass relative to the barrier release
    res["timeline_cycles"] = {k: float(v) for k, v in d.items()}
  out[name] = res
  print(name, json.dumps({k: v for k, v in res.items() if k != "times"}), flush=True)
  s.close()
if a.out:
  with open(a.out, "w") as f:
    json.dump(out, f, indent=1)
