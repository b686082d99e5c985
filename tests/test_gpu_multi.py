"""Row-sharded HMC over 2 GPUs against the single-GPU result and the oracle, on both multi-GPU plans: the
persistent kernel with its in-kernel all-reduce over peer memory, and the per-pass launch + ncclAllReduce plan.
Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, N, D, T, L, eps, out, plan_name="stepwise"):
  import torch
  import torch.distributed as dist
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    import hmc_oracle as o
    from edward_b200 import _C, engine
    from edward_b200.sharding import shard_bounds
    X, y, _ = o.synth_data(N, D)
    lo, hi = shard_bounds(N, world, rank, block=1024 if N >= 4096 else 64)
    plan = {"stepwise": _C.PLAN_STEPWISE, "auto": _C.PLAN_AUTO}[plan_name]
    s = engine.GLMSampler(engine.GLMSpec(D), X[lo:hi], y[lo:hi], device="cuda:%d" % rank, plan=plan, n_rows_global=N)
    s.init_comm(world, rank)
    r0, u = o.synth_draws(T, D)
    params = torch.zeros(T, D, device="cuda:%d" % rank)
    sc, pos = s.set_trace(T)
    s.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
    n_acc, logp = s.read_state()
    th = (0.1 * np.arange(D) / D).astype(np.float32)
    lp, g = s.logp_grad(th)
    # a second, chunked call continues the chain: sequence numbers of the peer exchange carry across launches
    params2 = torch.zeros(T, D, device="cuda:%d" % rank)
    s.reset()
    s.run(params2, 0, 3, eps, L, r0=torch.tensor(r0[:3]), u=torch.tensor(u[:3]))
    s.run(params2, 3, T - 3, eps, L, r0=torch.tensor(r0[3:]), u=torch.tensor(u[3:]))
    s.read_state()
    info = s.plan_info()
    out[rank] = (params.cpu().numpy(), n_acc, logp, float(lp.cpu()[0]), g.cpu().numpy(), sc.cpu().numpy())
    out["extra%d" % rank] = (params2.cpu().numpy(), bool(getattr(s, "peer_exchange", False)), info["plan_in_use"],
                             info["launches_last_run"])
    s.close()
  finally:
    dist.destroy_process_group()


def test_two_gpu_row_shards_match_single_gpu_and_oracle():
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import torch.multiprocessing as mp
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import hmc_oracle as o
  N, D, T, L, eps = 20000, 54, 8, 5, 0.003
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(2, _free_port(), N, D, T, L, eps, out), nprocs=2, join=True)
  p0, n0, lp0, l0, g0, sc0 = out[0]
  p1, n1, lp1, l1, g1, sc1 = out[1]
  # every rank integrates the chain redundantly on identical all-reduced sums → bitwise identical
  assert np.array_equal(p0, p1) and n0 == n1 and lp0 == lp1 and l0 == l1 and np.array_equal(g0, g1)
  X, y, _ = o.synth_data(N, D)
  spec = o.GLMSpec(D)
  r0, u = o.synth_draws(T, D)
  p64 = np.zeros((T, D))
  infos, nacc = o.run(X, y, p64, r0, u, eps, L, spec)
  assert np.max(np.abs(p0 - p64)) <= 1e-4 * np.max(np.abs(p64))
  assert n0 == nacc
  th = (0.1 * np.arange(D) / D).astype(np.float32)
  assert abs(l0 - o.log_joint(X, y, th, spec)) <= 1e-5 * abs(l0)
  g64 = o.grad_log_joint(X, y, th, spec)
  assert np.max(np.abs(g0 - g64)) <= 1e-5 * np.max(np.abs(g64))


@pytest.mark.parametrize("N,D,T,L,eps", [(30000, 200, 8, 5, 0.002), (30000, 300, 6, 4, 0.002), (150, 300, 6, 4, 0.05)])
def test_two_gpu_peer_exchange_matches_nccl_plan_bitwise(N, D, T, L, eps):
  """The persistent kernel's in-kernel all-reduce (stores into peer inboxes over NVLink) against the ncclAllReduce
  plan: with two ranks both sum a+b, so the chains must agree bit for bit; one launch per run() call. D=200 uses the
  flag protocol, D=300 the two-level slice protocol of wide models; 150 rows leave each rank with a one-CTA grid."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import torch.multiprocessing as mp
  from edward_b200 import _C
  res = {}
  for plan_name in ("stepwise", "auto"):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), N, D, T, L, eps, out, plan_name), nprocs=2, join=True)
    res[plan_name] = (out[0], out[1], out["extra0"], out["extra1"])
  a0, a1, ax0, ax1 = res["auto"]
  s0, s1, sx0, sx1 = res["stepwise"]
  assert ax0[1] and ax1[1], "peer inboxes were not mapped"
  assert ax0[2] == _C.PLAN_PERSISTENT and ax0[3] == 1
  assert sx0[2] == _C.PLAN_STEPWISE
  assert np.array_equal(a0[0], a1[0]) and a0[1] == a1[1] and a0[2] == a1[2]
  assert np.array_equal(a0[0], s0[0]) and a0[1] == s0[1] and a0[2] == s0[2]
  assert np.array_equal(a0[5], s0[5])  # per-transition trace: logp, K, ratio, accept
  assert np.array_equal(ax0[0], a0[0]) and np.array_equal(ax1[0], a0[0])  # chunked launches == one launch


def _chain_worker(rank, world, port, N, D, C, T, L, eps, out):
  import torch
  import torch.distributed as dist
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    import hmc_oracle as o
    from edward_b200 import engine
    from edward_b200.sharding import shard_bounds
    X, y, _ = o.synth_data(N, D)
    lo, hi = shard_bounds(N, world, rank, block=1024)
    s = engine.GLMSampler(engine.GLMSpec(D), X[lo:hi], y[lo:hi], device="cuda:%d" % rank, n_rows_global=N, n_chains=C)
    s.init_comm(world, rank)
    rng = np.random.Generator(np.random.Philox(key=5))
    r0 = rng.standard_normal((T, C, D), dtype=np.float32)
    u = np.clip(rng.random((T, C), dtype=np.float32), 1e-7, 1 - 1e-7).astype(np.float32)
    params = torch.zeros(T, C, D, device="cuda:%d" % rank)
    s.run_chains(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
    n_acc, logp = s.read_chain_state()
    out[rank] = (params.cpu().numpy(), np.asarray(n_acc), np.asarray(logp), r0, u)
    s.close()
  finally:
    dist.destroy_process_group()


def test_two_gpu_row_sharded_vectorised_chains_match_oracle():
  """Config-5 shape in miniature: wide model, many chains, rows sharded over 2 GPUs, [grad, logp] of all chains
  all-reduced once per leapfrog step; every rank holds the same chains, each equal to an independent oracle run."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import torch.multiprocessing as mp
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import hmc_oracle as o
  N, D, C, T, L, eps = 6000, 200, 128, 4, 3, 0.01
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_chain_worker, args=(2, _free_port(), N, D, C, T, L, eps, out), nprocs=2, join=True)
  p0, n0, lp0, r0, u = out[0]
  p1, n1, lp1, _, _ = out[1]
  assert np.array_equal(p0, p1) and np.array_equal(n0, n1) and np.array_equal(lp0, lp1)
  X, y, _ = o.synth_data(N, D)
  spec = o.GLMSpec(D)
  for c in (0, 77, 127):
    p64 = np.zeros((T, D))
    infos, nacc = o.run(X, y, p64, r0[:, c], u[:, c], eps, L, spec)
    if any(info.margin < 1e-3 for info in infos):
      continue
    assert np.max(np.abs(p0[:, c] - p64)) <= 1e-4 * max(np.max(np.abs(p64)), 1e-3), c
    assert n0[c] == nacc


def _frontend_worker(rank, world, port, N, D, T, out):
  import torch
  import torch.distributed as dist
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  if world > 1:
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    import hmc_oracle as o
    import edward_b200 as ed
    from edward_b200 import tfshim as tf
    from edward_b200.models import Bernoulli, Empirical, Normal
    from edward_b200.sharding import shard_bounds
    Xall, yall, _ = o.synth_data(N, D)
    lo, hi = shard_bounds(N, world, rank, block=1024)
    ed.set_seed(7)
    X = tf.placeholder(tf.float32, [hi - lo, D])
    w = Normal(loc=tf.zeros(D), scale=tf.ones(D))
    y = Bernoulli(logits=ed.dot(X, w))
    qw = Empirical(params=tf.Variable(tf.zeros([T, D])))
    inference = ed.HMC({w: qw}, data={X: Xall[lo:hi], y: yall[lo:hi]})
    inference.run(step_size=0.5 / N, n_steps=5, n_print=0, device="cuda:%d" % rank)
    info = inference._sampler.plan_info()
    out[(world, rank)] = (qw.params.eval(), int(inference.n_accept.eval()), info["plan_in_use"], info["launches_last_run"])
  finally:
    if world > 1:
      dist.destroy_process_group()


def test_two_gpu_ed_hmc_row_sharded_through_the_frontend():
  """ed.HMC under torch.distributed: each rank passes its row shard as `data`, the sampler runs as one persistent
  launch per GPU with the in-kernel all-reduce, and every rank ends with the chain a single GPU produces."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import torch.multiprocessing as mp
  from edward_b200 import _C
  N, D, T = 40000, 54, 30
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_frontend_worker, args=(2, _free_port(), N, D, T, out), nprocs=2, join=True)
  mp.spawn(_frontend_worker, args=(1, _free_port(), N, D, T, out), nprocs=1, join=True)
  p0, n0, plan0, launches0 = out[(2, 0)]
  p1, n1, _, _ = out[(2, 1)]
  ps, ns, _, _ = out[(1, 0)]
  assert plan0 == _C.PLAN_PERSISTENT and launches0 == 1
  assert np.array_equal(p0, p1) and n0 == n1
  assert n0 == ns
  assert np.max(np.abs(p0 - ps)) <= 1e-4 * np.max(np.abs(ps))
