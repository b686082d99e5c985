"""World-size-2 tests of the host-side logic of the row-sharded path on CPU (gloo backend): shard bounds,
the unique-id broadcast, the global row count, and that the per-shard [grad, logp] sums all-reduce to the
full-data oracle value (the arithmetic the NCCL all-reduce carries on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    import hmc_oracle as o
    from edward_b200.sharding import allgather_bytes, broadcast_bytes, shard_bounds
    N, D, block = 5000, 7, 512
    lo, hi = shard_bounds(N, world, rank, block)
    X, y, _ = o.synth_data(N, D)
    spec = o.GLMSpec(D)
    theta = (0.1 * np.arange(D)).astype(np.float32)
    # unique-id style broadcast
    payload = bytes(range(128)) if rank == 0 else None
    uid = broadcast_bytes(payload, src=0)
    # global row count, as ed.HMC computes it for row shards
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt)
    # per-shard likelihood sums (no prior) → all-reduce → add the prior once
    Xs, ys = X[lo:hi], y[lo:hi]
    eta = o.linear_predictor(Xs, theta, spec, np.float64)
    r = o.log_lik_grad_eta(eta, ys, spec, np.float64)
    sums = np.concatenate([Xs.astype(np.float64).T @ r, [np.sum(o.log_lik_terms(eta, ys, spec, np.float64))]])
    t = torch.tensor(sums, dtype=torch.float64)
    dist.all_reduce(t)
    tot = t.numpy()
    # the peer-inbox exchange of the persistent plan: every rank's entry gathered (here over gloo), summed in rank
    # order on every rank → identical bits everywhere; handles travel with allgather_bytes in rank order
    table = allgather_bytes(bytes([rank]) * 64)
    inbox = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(inbox, torch.tensor(sums, dtype=torch.float64))
    acc = np.zeros_like(sums)
    for entry in inbox:
      acc = acc + entry.numpy()
    peer_ok = table == b"".join(bytes([q]) * 64 for q in range(world)) and np.allclose(acc, tot, rtol=1e-14)
    grad = tot[:D] + o.normal_log_prob_grad(theta, spec.prior_loc, spec.prior_scale, np.float64)
    logp = tot[D] + np.sum(o.normal_log_prob(theta, spec.prior_loc, spec.prior_scale, np.float64))
    want_g = o.grad_log_joint(X, y, theta, spec)
    want_lp = o.log_joint(X, y, theta, spec)
    ok = (peer_ok and uid == bytes(range(128)) and int(cnt.item()) == N and np.allclose(grad, want_g, rtol=1e-10)
          and abs(logp - want_lp) < 1e-8 * abs(want_lp))
    out[rank] = (ok, lo, hi)
  finally:
    dist.destroy_process_group()


def test_row_shard_host_logic_world2():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  assert all(out[r][0] for r in range(world)), dict(out)
  assert out[0][1] == 0 and out[0][2] == out[1][1] and out[1][2] == 5000


def test_shard_bounds_cover_rows_exactly():
  sys.path.insert(0, ROOT)
  from edward_b200.sharding import shard_bounds
  for n in (1, 65535, 65536, 65537, 581012, 10_000_000):
    for world in (1, 2, 3, 4, 8):
      prev = 0
      for rank in range(world):
        lo, hi = shard_bounds(n, world, rank)
        assert lo == prev and hi >= lo and (lo % 65536 == 0 or lo == n)
        prev = hi
      assert prev == n
  with pytest.raises(ValueError):
    shard_bounds(10, 2, 2)
