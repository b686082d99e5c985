"""GPU parity: libedhmc (through the C ABI, via edward_b200.engine) vs the CPU oracle on the same inputs.

Tolerances (from BASELINE.json north_star): log joint and gradient within 1e-5 relative; positions
after L leapfrog steps within 1e-4 relative; accept decisions identical (near-ties, where
|log u - ratio| is below TIE_EPS, are counted separately and must be rare).
"""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu

REL_LOGP = 1e-5
REL_GRAD = 1e-5
REL_POS = 1e-4
TIE_EPS = 1e-3


def _mk(N, D, has_bias=False, family=o.BERNOULLI_LOGIT, seed=0, prior_scale=1.0, lik_scale=1.0):
  rng = np.random.default_rng(seed)
  X = rng.standard_normal((N, D)).astype(np.float32)
  w_true = (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32)
  eta = X.astype(np.float64) @ w_true
  if family == o.BERNOULLI_LOGIT:
    y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.int32)
  elif family == o.NORMAL_IDENTITY:
    y = (eta + lik_scale * rng.standard_normal(N)).astype(np.float32)
  else:
    y = rng.poisson(np.exp(np.clip(eta, -3, 3))).astype(np.int32)
  P = D + int(has_bias)
  spec = o.GLMSpec(D, has_bias, family, np.zeros(P, np.float32), np.full(P, prior_scale, np.float32), lik_scale)
  return X, y, spec


def _sampler(X, y, spec, **kw):
  from edward_b200 import engine
  es = engine.GLMSpec(spec.n_features, spec.has_bias, spec.family, spec.prior_loc, spec.prior_scale, spec.lik_scale)
  return engine.GLMSampler(es, X, y, **kw)


def _rel(a, b):
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


SHAPES = [
    (40, 1, True, o.BERNOULLI_LOGIT),      # cfg 1 shape
    (1024, 54, False, o.BERNOULLI_LOGIT),  # cfg 2 feature count
    (4096, 1000, False, o.BERNOULLI_LOGIT),  # cfg 4 feature count
    (1, 1, False, o.BERNOULLI_LOGIT),
    (3, 2, True, o.BERNOULLI_LOGIT),
    (33, 3, False, o.BERNOULLI_LOGIT),
    (777, 7, True, o.BERNOULLI_LOGIT),
    (5000, 100, False, o.BERNOULLI_LOGIT),
    (2500, 130, True, o.BERNOULLI_LOGIT),
    (1500, 257, False, o.BERNOULLI_LOGIT),
    (600, 2048, False, o.BERNOULLI_LOGIT),
    (300, 2047, True, o.BERNOULLI_LOGIT),
    (70001, 54, False, o.BERNOULLI_LOGIT),
    (1000, 10, True, o.NORMAL_IDENTITY),
    (50, 1, False, o.NORMAL_IDENTITY),
    (2000, 12, True, o.POISSON_LOG),
]


@pytest.mark.parametrize("N,D,bias,fam", SHAPES)
def test_logp_grad_matches_oracle(N, D, bias, fam):
  X, y, spec = _mk(N, D, bias, fam, seed=N + D, lik_scale=0.7)
  s = _sampler(X, y, spec)
  rng = np.random.default_rng(1)
  for scale in (0.0, 0.3):
    theta = (scale * rng.standard_normal(spec.n_params) / np.sqrt(D)).astype(np.float32)
    lp, g = s.logp_grad(theta)
    lp = float(lp.cpu()[0])
    g = g.cpu().numpy()
    lp64 = float(o.log_joint(X, y, theta, spec, np.float64))
    g64 = o.grad_log_joint(X, y, theta, spec, np.float64)
    assert abs(lp - lp64) <= REL_LOGP * abs(lp64), (lp, lp64)
    assert _rel(g, g64) <= REL_GRAD, _rel(g, g64)
    # the float32 restatement of the reference path is no closer to the truth than we are allowed to be
    lp32 = float(o.log_joint(X, y, theta, spec, np.float32))
    assert abs(lp - lp32) <= max(REL_LOGP * abs(lp32), 4 * abs(lp32 - lp64))
  s.close()


def _run_both(X, y, spec, T, L, eps, plan=0, z0=None, seed=3):
  import torch
  P = spec.n_params
  r0, u = o.synth_draws(T, P, seed=seed)
  init = np.zeros((T, P), np.float32)
  if z0 is not None:
    init[0] = z0
  p64 = init.astype(np.float64)
  infos, nacc = o.run(X, y, p64, r0, u, eps, L, spec, np.float64)
  s = _sampler(X, y, spec, plan=plan)
  params = torch.tensor(init, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_accept, logp_cur = s.read_state()
  return infos, nacc, p64, s, params.cpu().numpy(), sc.cpu().numpy(), pos.cpu().numpy(), n_accept, logp_cur


RUNS = [
    # N, D, bias, fam, T, L, eps
    (40, 1, True, o.BERNOULLI_LOGIT, 60, 2, 0.6),          # cfg 1 hyper-parameters
    (1024, 54, False, o.BERNOULLI_LOGIT, 25, 10, 0.5 / 1024),  # cfg 2 hyper-parameters at small N
    (1024, 54, False, o.BERNOULLI_LOGIT, 25, 5, 0.02),
    (4096, 1000, False, o.BERNOULLI_LOGIT, 8, 4, 0.003),
    (500, 10, True, o.NORMAL_IDENTITY, 30, 3, 0.01),
    (2000, 12, True, o.POISSON_LOG, 20, 3, 0.004),
    (9000, 54, False, o.BERNOULLI_LOGIT, 12, 10, 0.004),
]


@pytest.mark.parametrize("plan", [1, 2])
@pytest.mark.parametrize("N,D,bias,fam,T,L,eps", RUNS)
def test_run_matches_oracle(N, D, bias, fam, T, L, eps, plan):
  X, y, spec = _mk(N, D, bias, fam, seed=7 * N + D, lik_scale=0.5)
  infos, nacc, p64, s, params, sc, pos, n_accept, logp_cur = _run_both(X, y, spec, T, L, eps, plan=plan)
  ties = 0
  for i, info in enumerate(infos):
    # compare only while the two chains are in the same state (a legitimately flipped near-tie forks them)
    assert abs(sc[i, 0] - info.logp_old) <= REL_LOGP * abs(info.logp_old) + 1e-6, (i, sc[i], info)
    assert abs(sc[i, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (i, sc[i], info)
    assert _rel(pos[i], info.proposal) <= REL_POS, (i, _rel(pos[i], info.proposal))
    assert abs(sc[i, 4] - info.ratio) <= 1e-5 * max(abs(info.logp_new), 1.0) + 1e-4, (i, sc[i, 4], info.ratio)
    dev_accept = bool(sc[i, 6] > 0.5)
    if dev_accept != info.accept:
      assert info.margin < TIE_EPS, (i, info)
      ties += 1
      break
    assert _rel(params[i], p64[i]) <= REL_POS
  assert ties <= 1
  if ties == 0:
    assert n_accept == nacc
    assert abs(logp_cur - float(o.log_joint(X, y, p64[T - 1], spec))) <= REL_LOGP * abs(logp_cur) + 1e-6
  info = s.plan_info()
  assert info["plan_in_use"] == plan
  s.close()


def test_plans_agree_bitwise(monkeypatch):
  """Persistent and stepwise plans share the reduction tree and the zig-zag tile order (tied to the
  global leapfrog-step index), so they agree to the last bit, with the zig-zag on or off."""
  import torch
  X, y, spec = _mk(20000, 54, False, seed=5)
  T, L, eps = 6, 3, 0.003
  r0, u = o.synth_draws(T, 54)
  for zigzag in ("0", "1"):
    monkeypatch.setenv("EDHMC_ZIGZAG", zigzag)
    outs = []
    for plan in (1, 2):
      s = _sampler(X, y, spec, plan=plan)
      params = torch.zeros(T, 54, device="cuda")
      s.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
      outs.append((params.cpu().numpy(), s.read_state()))
      s.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]


def test_chunked_run_equals_single_run_and_cache_invalidation():
  """update()-style chunks (n_iter=1 launches) reproduce one long run; editing the current row of
  params between launches is honoured (the cached log joint / gradient are recomputed)."""
  import torch
  X, y, spec = _mk(3000, 54, False, seed=9)
  T, L, eps = 10, 3, 0.01
  r0, u = o.synth_draws(T, 54)
  r0t, ut = torch.tensor(r0, device="cuda"), torch.tensor(u, device="cuda")
  s1 = _sampler(X, y, spec)
  pa = torch.zeros(T, 54, device="cuda")
  s1.run(pa, 0, T, eps, L, r0=r0t, u=ut)
  s2 = _sampler(X, y, spec)
  pb = torch.zeros(T, 54, device="cuda")
  for t in range(T):
    s2.run(pb, t, 1, eps, L, r0=r0t[t:t + 1], u=ut[t:t + 1])
  assert torch.equal(pa, pb)
  assert s1.read_state() == s2.read_state()
  # now perturb the current state by hand and continue: must match an oracle started from that state
  T2 = 4
  pc = torch.zeros(T2, 54, device="cuda")
  pc[0] = 0.05
  s2.run(pc, 1, T2 - 1, eps, L, r0=r0t[:T2 - 1], u=ut[:T2 - 1])
  p64 = np.zeros((T2, 54))
  p64[0] = np.float32(0.05)
  o.run(X, y, p64, r0, u, eps, L, spec, np.float64, t0=1, n_iter=T2 - 1)
  assert _rel(pc.cpu().numpy(), p64) <= REL_POS
  s1.close()
  s2.close()


def test_error_conventions():
  import torch
  from edward_b200 import _C
  X, y, spec = _mk(100, 5)
  Xbad = X.copy()
  Xbad[17, 3] = np.inf
  with pytest.raises(_C.NonFiniteError):
    _sampler(Xbad, y, spec)
  s = _sampler(X, y, spec)
  params = torch.zeros(4, 5, device="cuda")
  with pytest.raises(_C.RangeError):
    s.run(params, 3, 2, 0.1, 2)
  with pytest.raises(IndexError):
    s.run(params, 4, 1, 0.1, 2)
  s.run(params, 0, 4, 0.1, 2)
  s.reset()
  assert s.read_state()[0] == 0
  s.close()


def test_device_rng_samples_the_reference_posterior():
  """tests/inferences/hmc_test.py:14-46 — Normal-Normal, 50 zeros, posterior N(0, 1/sqrt(51)); HMC defaults
  step_size=0.25, n_steps=2 (hmc.py:45); asserts |mean| <= 0.1, std within 0.1 of 0.140, n_accept > 0.1."""
  import torch
  N = 50
  X = np.ones((N, 1), np.float32)
  y = np.zeros(N, np.float32)
  spec = o.GLMSpec(1, False, o.NORMAL_IDENTITY, np.zeros(1, np.float32), np.ones(1, np.float32), 1.0)
  s = _sampler(X, y, spec)
  s.seed(42)
  T = 2000
  params = torch.ones(T, 1, device="cuda")
  s.run(params, 0, T, 0.25, 2)
  p = params.cpu().numpy()
  assert abs(p.mean()) <= 0.1
  assert abs(o.empirical_stddev(p)[0] - np.sqrt(1 / 51)) <= 0.1 * np.sqrt(1 / 51) + 0.1
  n_accept, _ = s.read_state()
  assert n_accept > 0.1 * T
  # and tighter than the reference asks: the chain is long enough for 3-sigma bounds
  assert abs(o.empirical_stddev(p)[0] - np.sqrt(1 / 51)) < 0.03
  s.close()


def test_full_size_cfg2_against_torch_float64():
  """cfg 2 shape (581,012 x 54): log joint and gradient against an independent float64 evaluation on the
  device (the numpy oracle would take too long here), plus the size-independent identity
  grad(theta) - grad_prior(theta) = X^T (y - sigmoid(X theta))."""
  import torch
  N, D = 581012, 54
  g = torch.Generator(device="cuda").manual_seed(0)
  X = torch.randn(N, D, device="cuda", generator=g)
  w_true = torch.randn(D, device="cuda", generator=g) / D ** 0.5
  y = (torch.rand(N, device="cuda", generator=g) < torch.sigmoid(X @ w_true)).to(torch.int32)
  spec = o.GLMSpec(D)
  s = _sampler(X, y, spec)
  theta = (0.1 * torch.randn(D, device="cuda", generator=g)).float()
  lp, gr = s.logp_grad(theta)
  X64, th64, y64 = X.double(), theta.double(), y.double()
  eta = X64 @ th64
  lik = -(torch.clamp(eta, min=0) - eta * y64 + torch.log1p(torch.exp(-eta.abs())))
  prior = (-0.5 * th64 ** 2 - 0.5 * np.log(2 * np.pi)).sum()
  lp64 = float(lik.sum() + prior)
  g64 = X64.T @ (y64 - torch.sigmoid(eta)) - th64
  assert abs(float(lp[0]) - lp64) <= REL_LOGP * abs(lp64)
  assert float((gr.double() - g64).abs().max() / g64.abs().max()) <= REL_GRAD
  info = s.plan_info()
  assert info["grid_ctas"] == torch.cuda.get_device_properties(0).multi_processor_count
  s.close()
