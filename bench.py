#!/usr/bin/env python
"""bench.py — HMC leapfrog steps/s for Bayesian logistic regression (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4]

A bench "step" is one `inference.run(step_size, n_steps)`-sized unit of the hot path: T transitions x L
leapfrog steps of one chain over the whole (synthetic, resident) design matrix.

  --gpus 1 (default)  workload cfg2: covertype shape N=581,012 D=54, T=100, L=10, eps=0.5/N
                      (docs/tex/iclr2017.tex:241-265), one chain, persistent plan.
  --gpus N>1          workload cfg4: N=10,000,000 D=1,000, rows sharded over the N GPUs in blocks of 65,536
                      rows, NCCL all-reduce of [grad, logp] per leapfrog step; strong scaling (fixed
                      total rows).  Launched by torchrun, one rank per GPU.
  --impl reference    the reference's CPU schedule (oracle/hmc_ref.c, OpenMP on the host cores; TensorFlow
                      itself cannot be installed here) on the same workload, bounded sample per step.

`value` times edhmc_run with CUDA events on the launching stream, inputs resident in HBM, L2 flushed
between steps. `e2e` times the public call chain ed.HMC(...).run() from pinned host arrays, including
the H2D copy of X and y and the D2H read of the samples. One JSON line on stdout (rank 0).

Beside the contract keys the line carries: `roofline` (HBM copy peak from MEASURED_PEAKS.json; `l2_peak` / `hbm_read_peak`
= read throughput of an L2-resident 100 MB buffer and of a 2 GiB buffer measured in this run with edhmc_probe_read),
`timeline` (per-pass clock64 stamps of the persistent kernel, medians over CTAs and passes), `parity_check` (the run's
own results against a float64 torch evaluation on the same device data; at N>1 also that every rank holds the same
chain), `chains` (cfg 3 on the N=1 line, cfg 5 on the N=8 line: the vectorised-chain tensor-core paths), `cfg1` (the
reference example as shipped, transitions/s), and the same-workload scaling references (`scale_series_n1` at N=1,
`n1_same_workload` + `efficiency_same_workload` at N>1, timed over the same number of steps).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg2": dict(N=581012, D=54, T=100, L=10, name="cfg2: covertype-shaped N=581012 D=54, 1 chain, T=100, n_steps=10, step_size=0.5/N"),
    "cfg4": dict(N=10_000_000, D=1000, T=100, L=10, name="cfg4: N=10000000 D=1000, 1 chain, T=100, n_steps=10, step_size=0.5/N, rows sharded"),
    "cfg5": dict(N=1_250_000, D=1000, T=1, L=4, C=1024, name="cfg5: N=1250000 rows PER GPU (10M at 8 GPUs) D=1000, 1024 vectorised chains (two tcgen05 3xTF32 GEMMs per step), rows sharded, ncclAllReduce of [grad, logp] per leapfrog step"),
    "cfg3": dict(N=581012, D=54, T=4, L=10, C=256, name="cfg3: covertype-shaped N=581012 D=54, 256 vectorised chains (tcgen05 3xTF32), T=4, n_steps=10"),
}
GEN_BLOCK = 65536
BASE_SEED = 42


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default=None, choices=[None, "cfg2", "cfg3", "cfg4", "cfg5"])
  ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                  help="N>1: in-kernel all-reduce over peer memory (default) or per-pass launch + ncclAllReduce")
  ap.add_argument("--no-scale-ref", action="store_true", help="skip the 1-GPU cfg4 point added to the N=1 line")
  ap.add_argument("--rows", type=int, default=None, help="override the row count (debugging)")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-extras", action="store_true", help="skip the chains / cfg1 / probe / timeline / parity blocks")
  ap.add_argument("--scale-steps", type=int, default=None, help="steps of the same-workload scaling reference (default min(steps, 5))")
  return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data: X ~ N(0,1), w_true ~ N(0,1/D), y ~ Bernoulli(sigmoid(X w_true)), generated
# in 65,536-row blocks seeded by block index so that any sharding on block boundaries gives the same data
# ------------------------------------------------------------------------------------------------
def gen_device_rows(torch, dev, row_start, n_rows, D):
  assert row_start % GEN_BLOCK == 0
  gw = torch.Generator(device=dev).manual_seed(BASE_SEED + (1 << 40))
  w_true = torch.randn(D, device=dev, generator=gw) / D ** 0.5
  X = torch.empty(n_rows, D, device=dev, dtype=torch.float32)
  y = torch.empty(n_rows, device=dev, dtype=torch.int32)
  done = 0
  b = row_start // GEN_BLOCK
  while done < n_rows:
    n = min(GEN_BLOCK, n_rows - done)
    g = torch.Generator(device=dev).manual_seed(BASE_SEED + b)
    xb = torch.randn(n, D, device=dev, generator=g)
    X[done:done + n] = xb
    y[done:done + n] = (torch.rand(n, device=dev, generator=g) < torch.sigmoid(xb @ w_true)).to(torch.int32)
    done += n
    b += 1
  return X, y


def shard_bounds(N, world, rank):
  from edward_b200.sharding import shard_bounds as sb
  return sb(N, world, rank, GEN_BLOCK)


class ClockSampler(object):
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

  def __init__(self, index):
    self.rows = []
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits", "-lms", "100"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.th = threading.Thread(target=self._read, daemon=True)
      self.th.start()
    except Exception:
      self.proc = None
    self.mark = 0

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.time(), line.strip()))

  def start(self):
    deadline = time.time() + 8.0
    while self.proc and not self.rows and time.time() < deadline:  # nvidia-smi takes a moment to emit
      time.sleep(0.05)
    self.t0 = time.time()

  def stop(self):
    self.t1 = time.time()
    if self.proc:
      self.proc.terminate()
    sm, smmax, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ts, line in self.rows:
      if not (self.t0 - 0.05 <= ts <= self.t1 + 0.15):
        continue
      f = [x.strip() for x in line.split(",")]
      try:
        sm.append(float(f[0]))
        smmax = max(smmax, float(f[1]))
        for nm, v in zip(names, f[3:7]):
          if v.lower().startswith("active"):
            reasons.add(nm)
      except Exception:
        pass
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smmax or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's schedule in C/OpenMP (oracle/hmc_ref.c)
# ------------------------------------------------------------------------------------------------
_CPU_ROWS = {}


def cpu_rows(wl, rows_cap):
  key = (wl["N"], wl["D"], rows_cap)
  if key not in _CPU_ROWS:
    _CPU_ROWS[key] = _cpu_rows(wl, rows_cap)
  return _CPU_ROWS[key]


def _cpu_rows(wl, rows_cap):
  """Host copy of the first rows of the workload (numpy Philox stream; values differ from the device
  generator's, the shapes and distributions are the same — only the time matters here)."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import hmc_oracle as o
  n = min(wl["N"], rows_cap)
  X, y, _ = o.synth_data(n, wl["D"])
  return o, X, y


def time_reference(wl, n_transitions, rows_cap, repeats=1):
  import numpy as np
  o, X, y = cpu_rows(wl, rows_cap)
  import ref_c
  spec = o.GLMSpec(wl["D"])
  r0, u = o.synth_draws(n_transitions, wl["D"])
  params = np.zeros((n_transitions, wl["D"]), np.float32)
  best = None
  for _ in range(repeats):
    t = time.perf_counter()
    ref_c.run(X, y, params, r0, u, 0.5 / wl["N"], wl["L"], spec, trace=False)
    dt = time.perf_counter() - t
    best = dt if best is None else min(best, dt)
  scale = wl["N"] / float(X.shape[0])  # linear extrapolation when a row subsample is timed
  return best * scale, X.shape[0], ref_c.num_threads()


def run_reference(args, wl, wl_key):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is one process that may use the whole host
  if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
  rows_cap = wl["N"] if wl_key == "cfg2" else wl["N"] // 64
  n_tr = 2
  for _ in range(args.warmup):
    time_reference(wl, 1, rows_cap)
  times = []
  for _ in range(args.steps):
    dt, nrows, threads = time_reference(wl, n_tr, rows_cap)
    times.append(dt)
  total = sum(times)
  steps_per_s = args.steps * n_tr * wl["L"] / total
  sample = "%d transitions x (L+1 gradient + 2 forward evaluations) per step on %d of %d rows%s" % (
      n_tr, nrows, wl["N"], "" if nrows == wl["N"] else " (time scaled linearly to all rows)")
  line = {
      "impl": "reference", "metric": "hmc_leapfrog_steps_per_s", "value": steps_per_s, "unit": "leapfrog steps/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
      "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": wl["name"], "note": "CPU restatement of the reference's TF schedule (not TensorFlow)"},
      "cpu_baseline": {"value": steps_per_s, "unit": "leapfrog steps/s", "cores": threads, "kind": "port", "sample": sample},
      "e2e": {"value": steps_per_s, "unit": "leapfrog steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line))


def tf32_peak(torch=None, dev=None):
  """Dense TF32 GEMM throughput of this GPU (the tensor roofline denominator; 3xTF32 executes 3 MMA flops per
  algorithmic flop): measured in this run with the protocol of MEASURED_PEAKS.json (torch.matmul 8192^3, best of 10)
  when a device is given, else the figure recorded in profiles/."""
  if torch is not None:
    try:
      old = torch.backends.cuda.matmul.allow_tf32
      torch.backends.cuda.matmul.allow_tf32 = True
      n = 8192
      a = torch.randn(n, n, device=dev)
      b = torch.randn(n, n, device=dev)
      for _ in range(3):
        a @ b
      torch.cuda.synchronize(dev)
      best = None
      for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
      torch.backends.cuda.matmul.allow_tf32 = old
      del a, b
      return 2.0 * n ** 3 / best / 1e9, "measured in this run (torch.matmul TF32 8192^3, best of 10)"
    except Exception:  # noqa: BLE001
      pass
  for name in ("r02_tf32_peak.json", "r01_tf32_peak.json"):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
      with open(p) as f:
        return float(json.load(f)["tf32_tflops"]), "measured on this pool (profiles/%s, tools/tf32_peak.py)" % name
  return 1100.0, "nominal dense TF32 (no measurement available)"


def read_peaks(torch, dev):
  """Read throughput of an L2-resident buffer (100 MB, re-read in place) and of an HBM-sized one (2 GiB) with
  edhmc_probe_read, best of 5 launches timed with CUDA events; LDG.128 and TMA bulk-copy paths, the better one counts."""
  from edward_b200 import _C
  lib = _C.lib()
  out = {}
  st = torch.cuda.current_stream(dev).cuda_stream
  sink = torch.zeros(4, dtype=torch.float32, device=dev)
  for key, nbytes, iters in (("l2", 100 * 1000 * 1000 // 32768 * 32768, 100), ("hbm_read", 2 * 1024 ** 3, 4)):
    buf = torch.ones(nbytes // 4, dtype=torch.float32, device=dev)
    best = 0.0
    for mode in (0, 1):
      _C.check(lib.edhmc_probe_read(buf.data_ptr(), nbytes, 2, mode, sink.data_ptr(), st))
      torch.cuda.synchronize(dev)
      for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.check(lib.edhmc_probe_read(buf.data_ptr(), nbytes, iters, mode, sink.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize(dev)
        best = max(best, nbytes * iters / e0.elapsed_time(e1) / 1e6)
    out[key] = best
    del buf
  return out


def timeline_of(torch, s, run_once, n_passes=48):
  """Medians (over CTAs and steady-state passes) of the persistent kernel's per-pass stamps, in SM cycles."""
  import numpy as np
  tl = s.set_timeline(n_passes)
  run_once()
  torch.cuda.synchronize()
  t = tl.cpu().numpy().astype(np.int64)[8:n_passes]
  s.set_timeline(0)
  if t.shape[0] < 4 or not np.any(t[:, :, 1]):
    return None
  wait = t[:, :, 8:16]
  step = float(np.median(t[1:, :, 0] - t[:-1, :, 0]))
  if t.shape[1] > 1 and not np.any(t[:, 1:, 3]):
    # leader protocol (chain.cuh): CTA 0 gathers the partials, integrates and sends the next position; the other CTAs
    # publish their sums and poll for it. Slots 3..5 are stamped by the leader only.
    lead, work = t[:, 0, :], t[:, 1:, :]
    d = {
        "unit": "SM cycles (clock64), median over passes 8..%d of one launch (workers: also over CTAs)" % (n_passes - 1),
        "protocol": "leader (flag-in-data partials -> CTA 0 -> flag-in-data position)",
        "pass_tiles_and_cta_reduce": float(np.median(work[:, :, 1] - work[:, :, 0])),
        "publish_partials": float(np.median(work[:, :, 2] - work[:, :, 1])),
        "worker_wait_for_next_position_median": float(np.median(work[1:, :, 0] - work[:-1, :, 2])),
        "worker_wait_fastest_cta": float(np.median((work[1:, :, 0] - work[:-1, :, 2]).min(axis=1))),
        "worker_wait_slowest_cta": float(np.median((work[1:, :, 0] - work[:-1, :, 2]).max(axis=1))),
        "leader_gather_partials": float(np.median(lead[:, 4] - lead[:, 2])),
        "leader_integrator_and_send": float(np.median(lead[:, 5] - lead[:, 4])),
        "serial_section": float(np.median(work[1:, :, 0] - work[:-1, :, 1])),
        "step": step,
        "step_ns_globaltimer": float(np.median(t[1:, 0, 6] - t[:-1, 0, 6])),
        "tile_wait_per_warp": float(np.median(wait[wait > 0])) if np.any(wait > 0) else 0.0,
    }
  else:
    d = {
        "unit": "SM cycles (clock64), median over CTAs and passes 8..%d of one launch" % (n_passes - 1),
        "protocol": ("flag-in-data partials -> group sums -> every CTA sums the groups and integrates (grid_barrier_* = publish -> totals)"
                     if t.shape[1] > 1 and np.any(t[:, 1:, 21]) else "grid barrier, every CTA reads every CTA's partials"),
        "pass_tiles_and_cta_reduce": float(np.median(t[:, :, 1] - t[:, :, 0])),
        "publish_partials": float(np.median(t[:, :, 2] - t[:, :, 1])),
        "grid_barrier_wait_median": float(np.median(t[:, :, 3] - t[:, :, 2])),
        "grid_barrier_wait_fastest_cta": float(np.median((t[:, :, 3] - t[:, :, 2]).min(axis=1))),
        "grid_barrier_wait_slowest_cta": float(np.median((t[:, :, 3] - t[:, :, 2]).max(axis=1))),
        "totals_ready_after_barrier": float(np.median(t[:, :, 4] - t[:, :, 3])),
        "integrator": float(np.median(t[:, :, 5] - t[:, :, 4])),
        "serial_section": float(np.median(t[:, :, 5] - t[:, :, 1])),
        "step": step,
        "step_ns_globaltimer": float(np.median(t[1:, 0, 6] - t[:-1, 0, 6])),
        "tile_wait_per_warp": float(np.median(wait[wait > 0])) if np.any(wait > 0) else 0.0,
    }
  d["serial_fraction_of_step"] = d["serial_section"] / d["step"] if d["step"] > 0 else None
  return d


def float64_logp_grad(torch, X, y, theta, world, dist):
  """log joint (Normal(0,1) priors) and gradient of the Bernoulli-logit model in float64 on the device data of this
  rank, summed over ranks: an evaluation independent of libedhmc (plain torch ops, 65,536-row chunks)."""
  th = theta.double()
  D = X.shape[1]
  g = torch.zeros(D, dtype=torch.float64, device=X.device)
  ll = torch.zeros((), dtype=torch.float64, device=X.device)
  for lo in range(0, X.shape[0], 65536):
    xb = X[lo:lo + 65536].double()
    yb = y[lo:lo + 65536].double()
    eta = xb @ th
    ll += (yb * eta - torch.nn.functional.softplus(eta)).sum()
    g += xb.t() @ (yb - torch.sigmoid(eta))
  pack = torch.cat([g, ll.reshape(1)])
  if world > 1:
    dist.all_reduce(pack)
  g, ll = pack[:D], pack[D]
  lp = ll + (-0.5 * th * th).sum() - D * 0.9189385332046727
  return float(lp), g - th


def parity_check(torch, s, X, y, params, world, rank, dist):
  """After the timed region: (1) log joint and gradient of the handle at theta = 0 and at a random theta against the
  float64 torch evaluation (1e-5 relative); (2) every rank holds bit-identical samples (N>1)."""
  D = X.shape[1]
  gen = torch.Generator(device=X.device).manual_seed(99)
  out = {"tolerance": 1e-5, "checks": []}
  ok = True
  for name, th in (("zero", torch.zeros(D, device=X.device)),
                   ("random", torch.randn(D, device=X.device, generator=gen) / D ** 0.5)):
    lp, g = s.logp_grad(th)
    lp64, g64 = float64_logp_grad(torch, X, y, th, world, dist)
    e_lp = abs(float(lp[0]) - lp64) / abs(lp64)
    e_g = float((g.double() - g64).abs().max() / g64.abs().max())
    out["checks"].append({"theta": name, "logp_rel_err": e_lp, "grad_rel_err": e_g})
    ok = ok and e_lp <= 1e-5 and e_g <= 1e-5
  if world > 1:
    h = params.view(torch.int32).to(torch.int64)
    sig = torch.stack([h.sum(), (h * torch.arange(1, h.numel() + 1, device=h.device).reshape(h.shape)).sum()])
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    same = all(bool(torch.equal(sigs[0], q)) for q in sigs)
    out["ranks_hold_identical_samples"] = same
    ok = ok and same
  out["samples_finite"] = bool(torch.isfinite(params).all())
  out["ok"] = bool(ok and out["samples_finite"])
  return out


def time_single_chain(torch, s, params, T, eps, L, steps, warmup, flush, barrier):
  """W warm-up + K timed steps of edhmc_run (CUDA events on the launching stream, L2 flushed between steps)."""
  def one_step():
    if flush is not None:
      flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(params, 0, T, eps, L)
    e1.record()
    return e0, e1
  for _ in range(warmup):
    one_step()
  barrier()
  evs = [one_step() for _ in range(steps)]
  barrier()
  return sum(a.elapsed_time(b) for a, b in evs)


def bench_chains(torch, dist, args, wl, dev, world, rank, steps, with_e2e=True, cpu_single=None):
  """Vectorised chains (cfg 3 on one GPU, cfg 5 row-sharded): value = chain leapfrog steps/s, tensor roofline."""
  from edward_b200 import engine
  n_loc, D, T, L, C = wl["N"], wl["D"], wl["T"], wl["L"], wl["C"]
  N = n_loc * world if world > 1 else n_loc  # rows are per GPU for the sharded workload (weak scaling)
  r_lo, r_hi = shard_bounds(N, world, rank)
  X, y = gen_device_rows(torch, dev, r_lo, r_hi - r_lo, D)
  s = engine.GLMSampler(engine.GLMSpec(D), X, y, device=dev, n_chains=C, n_rows_global=N)
  if world > 1:
    s.init_comm(world, rank, peers=False)
  s.seed(1234)
  params = torch.zeros(T, C, D, device=dev)
  flush = torch.empty(512 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
  for _ in range(3):
    s.run_chains(params, 0, T, 0.5 / N, L)
  torch.cuda.synchronize(dev)
  if world > 1:
    dist.barrier()
    torch.cuda.synchronize(dev)
  sampler = ClockSampler(dev.index) if rank == 0 else None
  if sampler:
    sampler.start()
  evs = []
  for _ in range(steps):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run_chains(params, 0, T, 0.5 / N, L)
    e1.record()
    evs.append((e0, e1))
  torch.cuda.synchronize(dev)
  if world > 1:
    dist.barrier()
    torch.cuda.synchronize(dev)
  clocks = sampler.stop() if sampler else None
  ms = sum(a.elapsed_time(b) for a, b in evs)
  per_rank_ms, allreduce_us = None, None
  if world > 1:
    mine = torch.tensor([ms], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    per_rank_ms = [float(g.item()) / steps for g in gathered]
    ms = max(float(g.item()) for g in gathered)
    buf = torch.zeros(C * (D + 1), dtype=torch.float64, device=dev)
    for _ in range(3):
      dist.all_reduce(buf)
    torch.cuda.synchronize(dev)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(10):
      dist.all_reduce(buf)
    a1.record()
    torch.cuda.synchronize(dev)
    allreduce_us = a0.elapsed_time(a1) * 100.0
  info = s.plan_info()
  # parity spot check: chains 0, C/2, C-1 of one gradient evaluation against the float64 torch evaluation
  gen = torch.Generator(device=dev).manual_seed(5)
  th = torch.randn(C, D, device=dev, generator=gen) / D ** 0.5
  lpc, gc = s.logp_grad_chains(th)
  worst_lp = worst_g = 0.0
  for c in (0, C // 2, C - 1):
    lp64, g64 = float64_logp_grad(torch, X, y, th[c], world, dist)
    worst_lp = max(worst_lp, abs(float(lpc[c]) - lp64) / abs(lp64))
    worst_g = max(worst_g, float((gc[c].double() - g64).abs().max() / g64.abs().max()))
  parity = {"chains_checked": [0, C // 2, C - 1], "logp_rel_err": worst_lp, "grad_rel_err": worst_g, "tolerance": 1e-5,
            "ok": bool(worst_lp <= 1e-5 and worst_g <= 1e-5)}
  # end to end through the public chains call: pinned host arrays -> GLMSampler(n_chains=C).run_chains -> samples on the host
  e2e = None
  if with_e2e and world == 1:
    Xh = torch.empty(X.shape, dtype=X.dtype, pin_memory=True)
    Xh.copy_(X)
    yh = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
    yh.copy_(y)
    times = []
    for i in range(4):
      flush.fill_(1.0)
      torch.cuda.synchronize(dev)
      t0 = time.perf_counter()
      s2 = engine.GLMSampler(engine.GLMSpec(D), Xh, yh, device=dev, n_chains=C)
      s2.seed(1234)
      p2 = torch.zeros(T, C, D, device=dev)
      s2.run_chains(p2, 0, T, 0.5 / N, L)
      host = p2.cpu()
      dt = time.perf_counter() - t0
      if i > 0:
        times.append(dt)
      assert host.shape == (T, C, D)
      s2.close()
    e2e = {"value": 3 * C * T * L / sum(times), "unit": "chain leapfrog steps/s", "h2d_bytes_per_step": Xh.numel() * 4 + yh.numel() * 4,
           "d2h_bytes_per_step": T * C * D * 4, "steps": 3,
           "call": "engine.GLMSampler(spec, pinned host X, y, n_chains=C).run_chains(params, 0, T, step_size, n_steps) + params.cpu() "
                   "(includes the one-off re-lay of X into the tensor-core operand layout)"}
    del Xh, yh
  s.close()
  del X, y, flush
  torch.cuda.empty_cache()
  if rank != 0:
    return None
  nsteps = steps * T * L
  alg = 4.0 * (float(N) / world) * D * C * nsteps / (ms * 1e-3) / 1e12  # per GPU (mean shard)
  peak, src = tf32_peak(torch, dev)
  wide = D > 64
  cpu = None
  if cpu_single is not None:
    cpu = {"value": cpu_single["value"], "unit": "chain leapfrog steps/s", "cores": cpu_single["cores"], "kind": "port",
           "sample": "C independent chains cost C single-chain runs on the CPU port, so its chain-steps/s equal the "
                     "single-chain figure measured in this run (" + cpu_single["sample"][:120] + " ...)"}
  return {
      "workload": wl["name"], "metric": "hmc_chain_leapfrog_steps_per_s", "value": C * nsteps / (ms * 1e-3), "unit": "chain leapfrog steps/s",
      "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "scaling": "weak", "dtype": "tf32x3 (fp32 accumulate)",
      "config": {"rows": N, "rows_per_gpu": n_loc, "features": D, "chains": C, "transitions_per_step": T, "leapfrog_per_transition": L,
                 "l2": "L2 flushed between steps", "rng": "device Philox",
                 "parallelism": "rows sharded over %d GPUs, ncclAllReduce of [grad, logp] (%d float64) per leapfrog step" % (world, C * (D + 1)) if world > 1 else "1 GPU"},
      "leapfrog_steps_of_all_chains_per_s": nsteps / (ms * 1e-3), "rows_steps_per_s": float(N) * C * nsteps / (ms * 1e-3),
      "per_rank_ms_per_step": per_rank_ms, "allreduce_us_same_payload": allreduce_us,
      "roofline": {"bound": "tensor", "achieved": 3.0 * alg, "peak": peak, "unit": "TFLOP/s", "frac": 3.0 * alg / peak,
                   "traffic": None, "peak_source": src, "algorithmic_tflops": alg,
                   "kernel": ("edhmc::k_mcw_gemm<1> + k_mcw_gemm<2>" if wide else "edhmc::k_mc_pass_tc3") + " (3xTF32: 3 executed MMA flops per algorithmic flop); per GPU"},
      "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": info["launches_last_run"] * steps, "clocks": clocks, "parity_check": parity,
  }


def run_chains(args, wl):
  """--workload cfg3 / cfg5 as the headline of the line (development runs; the driver's lines carry them as `chains`)."""
  import torch
  import torch.distributed as dist
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
  torch.cuda.set_device(dev)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
  blk = bench_chains(torch, dist, args, wl, dev, world, rank, args.steps)
  if rank == 0:
    line = dict(blk)
    line["metric"] = "hmc_leapfrog_steps_per_s"
    line.update({"warmup": 3, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
    line["config"] = dict(blk["config"], workload=blk["workload"])
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def bench_cfg1():
  """BASELINE config 1: the reference's example script as shipped (tools/run_reference_example.py) — 5,000 update()
  calls with a progress line and a posterior-predictive evaluation every 10 iterations; launch-latency bound."""
  try:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_reference_example as rre
    script = rre.find_script()
    import contextlib
    import io
    if script is None:
      return {"unavailable": "the reference example is not staged under baseline/_ref (run __graft_entry__.build() where /root/reference exists)"}
    with contextlib.redirect_stdout(io.StringIO()):
      rre.run(["--T", "300"], script)  # warm-up: first launch, allocator
      r = rre.run([], script)
    return {"workload": "cfg1: examples/bayesian_logistic_regression.py as shipped (N=40, D=1 + bias, T=5000, step_size=0.6, n_steps=2), "
                        "update() loop with print_progress and a 50-draw posterior-predictive evaluation every 10 iterations",
            "script": "reference file executed unchanged through module aliases (tools/run_reference_example.py)",
            "transitions_per_s": r["transitions_per_s"], "seconds": r["seconds"], "transitions": r["t"], "n_accept": r["n_accept"]}
  except Exception as e:  # noqa: BLE001
    return {"error": str(e)[:200]}


# ------------------------------------------------------------------------------------------------
def main():
  args = parse()
  wl_key = args.workload or ("cfg2" if args.gpus == 1 else "cfg4")
  wl = dict(WORKLOADS[wl_key])
  if args.rows:
    wl["N"] = args.rows
  if args.impl == "reference":
    run_reference(args, wl, wl_key)
    return
  if wl_key in ("cfg3", "cfg5"):
    run_chains(args, wl)
    return

  import numpy as np
  import torch
  import torch.distributed as dist
  from edward_b200 import engine

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if args.gpus > 1 and world != args.gpus:
    raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
  dev = torch.device("cuda", local_rank)
  torch.cuda.set_device(dev)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
  extras = not args.no_extras and not args.rows
  warmup = max(args.warmup, 3)

  N, D, T, L = wl["N"], wl["D"], wl["T"], wl["L"]
  eps = 0.5 / N
  r_lo, r_hi = shard_bounds(N, world, rank)
  X, y = gen_device_rows(torch, dev, r_lo, r_hi - r_lo, D)
  s = engine.GLMSampler(engine.GLMSpec(D), X, y, device=dev, n_rows_global=N,
                        plan=engine._C.PLAN_STEPWISE if (world > 1 and args.collective == "nccl") else engine._C.PLAN_AUTO)
  if world > 1:
    s.init_comm(world, rank, peers=args.collective == "peer")
  s.seed(1234)
  params = torch.zeros(T, D, device=dev)
  flush = torch.empty(512 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  # ---- the timed region: W warm-up steps, then exactly K steps between barriers; device time, max over ranks ----
  for _ in range(warmup):
    flush.fill_(1.0)
    s.run(params, 0, T, eps, L)
  barrier()
  sampler = ClockSampler(local_rank) if rank == 0 else None
  if sampler:
    sampler.start()
  barrier()
  wall0 = time.perf_counter()
  dev_ms = time_single_chain(torch, s, params, T, eps, L, args.steps, 0, flush, barrier)
  wall = time.perf_counter() - wall0
  clocks = sampler.stop() if sampler else None
  t_ms = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
  dev_ms = float(t_ms.item())
  info = s.plan_info()
  n_accept, _ = s.read_state()
  steps_total = args.steps * T * L
  value = steps_total / (dev_ms * 1e-3)
  launches = info["launches_last_run"] * args.steps

  # ---- parity self-check on the data and handle that were just timed ----
  parity = parity_check(torch, s, X, y, params, world, rank, dist) if extras or world > 1 else None

  # ---- per-pass timeline of the persistent kernel (one extra launch, outside the timed region) ----
  timeline = None
  if (extras or world > 1) and info["plan_in_use"] == 1:
    try:
      tl = timeline_of(torch, s, lambda: s.run(params, 0, T, eps, L))
      if world > 1:
        box = [None] * world
        dist.all_gather_object(box, tl)
        timeline = {"per_rank": box}
      else:
        timeline = tl
    except Exception as e:  # noqa: BLE001
      timeline = {"error": str(e)[:200]}

  # ---- strong-scaling reference: the same workload on rank 0 alone, same number of steps (only when sharded) ----
  n1_same = None
  scale_steps = args.scale_steps or min(args.steps, 5)
  if world > 1:
    if rank == 0:
      try:
        X1, y1 = gen_device_rows(torch, dev, 0, N, D)
        s1 = engine.GLMSampler(engine.GLMSpec(D), X1, y1, device=dev)
        s1.seed(1234)
        p1 = torch.zeros(T, D, device=dev)
        ms1 = time_single_chain(torch, s1, p1, T, eps, L, scale_steps, 1, flush, lambda: torch.cuda.synchronize(dev))
        n1_same = {"value": scale_steps * T * L / (ms1 * 1e-3), "unit": "leapfrog steps/s", "steps": scale_steps, "warmup": 1,
                   "ms_per_step": ms1 / scale_steps, "plan": "persistent, 1 GPU (rank 0 alone), same job, same data"}
        s1.close()
        del X1, y1, p1
      except Exception as e:  # noqa: BLE001
        n1_same = {"error": str(e)[:200]}
    dist.barrier()

  # ---- end to end through the public API: host arrays -> ed.HMC(...).run() -> samples on the host ----
  e2e = None
  if not args.no_e2e:
    import edward_b200 as ed
    from edward_b200 import graph as g
    from edward_b200 import tfshim as tf
    from edward_b200.models import Bernoulli, Empirical, Normal
    n_loc = r_hi - r_lo
    Xh = torch.empty(X.shape, dtype=X.dtype, pin_memory=True)
    Xh.copy_(X)
    yh = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
    yh.copy_(y)
    e2e_steps = max(3, min(args.steps, 20)) if world == 1 else max(2, min(args.steps, 4))
    times = []
    h2d = Xh.numel() * 4 + yh.numel() * 4
    d2h = T * D * 4 + 16
    for i in range(e2e_steps + 1):
      g.reset_default_graph()
      flush.fill_(1.0)
      torch.cuda.synchronize(dev)
      if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
      t0 = time.perf_counter()
      xs = tf.placeholder(tf.float32, [n_loc, D])  # under torch.distributed each rank passes its row shard
      beta = Normal(loc=tf.zeros(D), scale=tf.ones(D))
      ys = Bernoulli(logits=ed.dot(xs, beta))
      qbeta = Empirical(params=tf.Variable(tf.zeros([T, D])))
      inference = ed.HMC({beta: qbeta}, data={xs: Xh, ys: yh})
      inference.run(step_size=eps, n_steps=L, n_print=0, device=dev)
      samples = qbeta.params.eval()  # D2H read of the result
      n_acc = int(inference.n_accept.eval())
      dt = time.perf_counter() - t0
      if world > 1:  # whole job: the slowest rank
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
      if i > 0:
        times.append(dt)
      assert samples.shape == (T, D) and 0 <= n_acc <= T
      inference._sampler.close()
      del inference, qbeta
    e2e = {"value": e2e_steps * T * L / sum(times), "unit": "leapfrog steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps, "seconds_per_step": sum(times) / e2e_steps,
           "call": "ed.HMC({beta: qbeta}, data={X: pinned host array%s, y: ...}).run(step_size, n_steps) + qbeta.params.eval()"
                   % (" (this rank's row shard; bytes are per rank)" if world > 1 else "")}
    del Xh, yh

  peaks = None
  if extras and world == 1:
    try:
      peaks = read_peaks(torch, dev)
    except Exception as e:  # noqa: BLE001
      peaks = {"error": str(e)[:200]}
  s.close()
  del X, y, flush, params
  torch.cuda.empty_cache()

  # ---- CPU baseline (rank 0, N=1): the reference's schedule in C/OpenMP on a bounded sample ----
  cpu = None
  if world == 1 and rank == 0 and not args.no_cpu_baseline:
    rows_cap = N if wl_key == "cfg2" else N // 64
    probe, nrows, threads = time_reference(wl, 1, rows_cap)
    n_tr = int(max(2, min(60, 12.0 / max(probe * nrows / N, 1e-3))))
    dt, nrows, threads = time_reference(wl, n_tr, rows_cap)
    cpu = {"value": n_tr * L / dt, "unit": "leapfrog steps/s", "cores": threads, "kind": "port",
           "sample": "%d transitions x (L+1 gradient + 2 forward evaluations, CheckNumerics on) on %d of %d rows%s; "
                     "C/OpenMP restatement of the reference's TensorFlow schedule (oracle/hmc_ref.c), host has %d logical CPUs"
                     % (n_tr, nrows, N, "" if nrows == N else ", time scaled linearly", os.cpu_count())}

  # ---- vectorised chains: cfg 3 on the N=1 line, cfg 5 on the N=8 line ----
  chains = None
  if extras and wl_key in ("cfg2", "cfg4"):
    try:
      if world == 1 and wl_key == "cfg2":
        chains = bench_chains(torch, dist, args, dict(WORKLOADS["cfg3"]), dev, 1, 0, max(3, min(args.steps, 10)), cpu_single=cpu)
      elif world == 8:
        chains = bench_chains(torch, dist, args, dict(WORKLOADS["cfg5"]), dev, world, rank, max(2, min(args.steps, 5)))
    except Exception as e:  # noqa: BLE001
      chains = {"error": str(e)[:300]}

  # ---- the N=1 point of the row-sharded scaling series (cfg 4 on this GPU alone) ----
  scale_n1 = None
  if world == 1 and wl_key == "cfg2" and not args.no_scale_ref and not args.rows:
    try:
      w4 = WORKLOADS["cfg4"]
      X4, y4 = gen_device_rows(torch, dev, 0, w4["N"], w4["D"])
      s4 = engine.GLMSampler(engine.GLMSpec(w4["D"]), X4, y4, device=dev)
      s4.seed(1234)
      p4 = torch.zeros(w4["T"], w4["D"], device=dev)
      ms4 = time_single_chain(torch, s4, p4, w4["T"], 0.5 / w4["N"], w4["L"], scale_steps, 1, None, lambda: torch.cuda.synchronize(dev))
      v4 = scale_steps * w4["T"] * w4["L"] / (ms4 * 1e-3)
      scale_n1 = {"workload": w4["name"], "value": v4, "unit": "leapfrog steps/s", "steps": scale_steps, "warmup": 1,
                  "ms_per_step": ms4 / scale_steps,
                  "hbm_frac": (4.0 * w4["N"] * w4["D"] + 4.0 * w4["N"]) * v4 / 1e9 / measured_peaks()[0],
                  "note": "X (40 GB) is far larger than L2: every leapfrog step streams it from HBM"}
      s4.close()
      del X4, y4, p4
      torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
      scale_n1 = {"error": str(e)[:200]}

  cfg1 = bench_cfg1() if (extras and world == 1 and rank == 0 and wl_key == "cfg2") else None

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  peak, peak_src = measured_peaks()
  alg_bytes_step = 4.0 * (r_hi - r_lo) * D + 4.0 * (r_hi - r_lo)  # this rank's shard: X once + y once per leapfrog step
  launch_ms = dev_ms / args.steps if info["plan_in_use"] == 1 else None
  achieved = alg_bytes_step * T * L / (dev_ms / args.steps * 1e-3) / 1e9
  traffic, traffic_src = None, None
  try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      tj = json.load(f)
    if wl_key in tj and info["plan_in_use"] == 1 and not args.rows:
      traffic = tj[wl_key]["bytes_per_launch"]
      traffic_src = "static capture: %s (ncu --set full of the same launch; not measured in this run)" % tj[wl_key]["source"]
  except Exception:
    pass
  roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
              "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
              "kernel": "edhmc::k_hmc (one persistent launch = T*L passes)" if info["plan_in_use"] == 1
              else "edhmc::k_hmc mode 1 (one pass per launch) + NCCL all-reduce + chain kernels",
              "algorithmic_bytes_per_launch": alg_bytes_step * T * L if info["plan_in_use"] == 1 else alg_bytes_step,
              "launch_ms": launch_ms}
  if peaks and "l2" in peaks:
    x_mb = 4e-6 * (r_hi - r_lo) * (D + 1)
    roofline.update({
        "l2_peak": peaks["l2"], "l2_frac": achieved / peaks["l2"], "hbm_read_peak": peaks["hbm_read"],
        "hbm_read_frac": achieved / peaks["hbm_read"],
        "peaks_measured_here": "edhmc_probe_read, best of LDG.128 / TMA bulk paths, best of 5 launches: 100 MB re-read in place "
                               "(L2-resident) and 2 GiB (HBM)",
        "residency": "the pass re-reads %.1f MB per leapfrog step; the L2 holds 126 MB, so most of it is served from L2 "
                     "(DRAM traffic %s of the algorithmic bytes) and `frac` above 1 is against the HBM COPY peak, not a "
                     "bandwidth bound: the pass is bound by per-row latency (profiles/README round 2)"
                     % (x_mb, ("%.0f %%" % (100.0 * traffic / roofline["algorithmic_bytes_per_launch"])) if traffic else "a fraction")})
    if info.get("ring_mode") == 2:
      n_res, n_tm = info.get("smem_resident_tiles_per_cta", 0), info.get("tmem_resident_tiles_per_cta", 0)
      tiles = ((r_hi - r_lo) + 31) // 32
      per_cta = -(-tiles // max(info["grid_ctas"], 1))
      roofline["on_chip"] = {
          "smem_resident_tiles_per_cta": n_res, "tmem_resident_tiles_per_cta": n_tm, "tiles_per_cta": per_cta,
          "fraction_of_rows_on_chip": min(1.0, (n_res + n_tm) / max(per_cta, 1)),
          "note": "ring mode 2 on the persistent plan: these 32-row tiles are copied once per launch into shared memory / "
                  "tensor memory and read from there in every pass; only the remaining rows cross the SM's L2 port, so the "
                  "algorithmic-bytes rate above can exceed both the HBM and the L2 figure without contradiction"}

  line = {
      "metric": "hmc_leapfrog_steps_per_s", "value": value, "unit": "leapfrog steps/s", "n_gpus": world,
      "steps": args.steps, "warmup": warmup, "ms_per_step": dev_ms / args.steps,
      "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": wl["name"], "rows": N, "features": D, "transitions_per_step": T, "leapfrog_per_transition": L,
                 "chains": 1, "plan": {1: "persistent", 2: "stepwise"}[info["plan_in_use"]],
                 "parallelism": ("rows sharded over %d GPUs, %s" % (world, "one persistent launch per GPU, in-kernel all-reduce of [grad, logp] through peer inboxes over NVLink each leapfrog step"
                                  if info["plan_in_use"] == 1 else "NCCL all-reduce per leapfrog step")) if world > 1 else "1 GPU",
                 "l2": ("L2 flushed between steps (512 MiB write); within a step X (%.1f MB) is re-streamed every leapfrog step" % (4e-6 * (r_hi - r_lo) * D))
                       + (" — except the rows a persistent launch keeps in shared / tensor memory (roofline.on_chip)" if info.get("ring_mode") == 2 else ""),
                 "rng": "device Philox", "grid_ctas": info["grid_ctas"], "ring_stages": info["ring_stages"], "tile_rows": info["tile_rows"],
                 "ring_mode": info.get("ring_mode"), "warps_per_cta": info["warps_per_cta"]},
      "rows_steps_per_s": value * N,
      "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
      "clocks": clocks, "n_accept_total": n_accept, "wall_s": wall,
  }
  if parity is not None:
    line["parity_check"] = parity
  if timeline is not None:
    line["timeline"] = timeline
  if n1_same is not None:
    line["n1_same_workload"] = n1_same
    if "value" in n1_same:
      line["efficiency_same_workload"] = value / (world * n1_same["value"])
      line["speedup_same_workload"] = value / n1_same["value"]
  if scale_n1 is not None:
    line["scale_series_n1"] = scale_n1
  if chains is not None:
    line["chains"] = chains
  if cfg1 is not None:
    line["cfg1"] = cfg1
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
