"""Parity at BASELINE.json's FULL sizes, where the numpy oracle is too slow: the CUDA paths against a float64 torch
restatement of the same transition (hmc.py:81-130,195-210; TF densities of oracle/hmc_oracle.py) evaluated on the same
device data, with injected momentum / uniform draws. Tolerances of the north star: log joint and gradient 1e-5 relative,
positions after L leapfrog steps 1e-4 relative, accept decisions identical (a proposal whose |log u - ratio| < 1e-3 is a
near-tie and is not counted).

  cfg 2        581,012 x 54, one chain: T=3 x L=10 trajectory
  cfg 3        581,012 x 54 x 256 chains (tcgen05 3xTF32, fp32 TMEM accumulation over a CTA's whole row range):
               chains 0 / 127 / 255, gradient + T=2 x L=10 trajectory
  cfg 4 shard  1,250,000 x 1000 (one of eight row shards), one chain: gradient + T=2 x L=3 trajectory
  cfg 5 shard  1,250,000 x 1000 x 1024 chains (two-GEMM path): chains 0 / 511 / 1023, gradient + T=1 x L=2 trajectory
The small-size parity against the oracle proper is in test_gpu_engine_parity.py / test_gpu_chains.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_LOGP, REL_GRAD, REL_POS, TIE_EPS = 1e-5, 1e-5, 1e-4, 1e-3
GEN_BLOCK = 65536


def _gen(N, D, seed=42):
  import torch
  dev = "cuda"
  gw = torch.Generator(device=dev).manual_seed(seed + (1 << 40))
  w_true = torch.randn(D, device=dev, generator=gw) / D ** 0.5
  X = torch.empty(N, D, device=dev)
  y = torch.empty(N, device=dev, dtype=torch.int32)
  for b, lo in enumerate(range(0, N, GEN_BLOCK)):
    n = min(GEN_BLOCK, N - lo)
    g = torch.Generator(device=dev).manual_seed(seed + b)
    xb = torch.randn(n, D, device=dev, generator=g)
    X[lo:lo + n] = xb
    y[lo:lo + n] = (torch.rand(n, device=dev, generator=g) < torch.sigmoid(xb @ w_true)).to(torch.int32)
  return X, y


def _logp_grad64(X, y, th):
  """float64 log joint (Normal(0,1) priors, Bernoulli-logit likelihood in TF's stable form) and gradient, chunked."""
  import torch
  th = th.double()
  D = X.shape[1]
  g = torch.zeros(D, dtype=torch.float64, device=X.device)
  ll = torch.zeros((), dtype=torch.float64, device=X.device)
  for lo in range(0, X.shape[0], 4 * GEN_BLOCK):
    xb = X[lo:lo + 4 * GEN_BLOCK].double()
    yb = y[lo:lo + 4 * GEN_BLOCK].double()
    eta = xb @ th
    ll += -(torch.clamp(eta, min=0) - eta * yb + torch.log1p(torch.exp(-eta.abs()))).sum()
    g += xb.t() @ (yb - torch.sigmoid(eta))
  lp = ll + (-0.5 * th * th - 0.5 * np.log(2 * np.pi)).sum()
  return lp, g - th


def _hmc64(X, y, z0, r0, u, eps, L):
  """T transitions in float64: returns (rows [T, D], accepts [T], margins [T])."""
  import torch
  T = r0.shape[0]
  z_cur = z0.double().clone()
  rows, accepts, margins = [], [], []
  for t in range(T):
    z, r = z_cur.clone(), r0[t].double().clone()
    lp_old, g = _logp_grad64(X, y, z)
    for _ in range(L):
      r = r + 0.5 * eps * g
      z = z + eps * r
      lp_new, g = _logp_grad64(X, y, z)
      r = r + 0.5 * eps * g
    ratio = 0.5 * (r0[t].double() ** 2).sum() - 0.5 * (r ** 2).sum() + lp_new - lp_old
    log_u = torch.log(u[t].double())
    acc = bool(log_u < ratio)
    if acc:
      z_cur = z
    rows.append(z_cur.clone())
    accepts.append(acc)
    margins.append(abs(float(log_u - ratio)))
  return torch.stack(rows), accepts, margins


def _check_rows(got, want, accepts_got, accepts_want, margins, what):
  import torch
  for t in range(want.shape[0]):
    if margins[t] < TIE_EPS:
      return  # a genuine near-tie: later rows may legitimately differ
    assert bool(accepts_got[t]) == bool(accepts_want[t]), (what, t, margins[t])
    rel = float((got[t].double() - want[t]).abs().max() / want[t].abs().max().clamp(min=1e-30))
    assert rel <= REL_POS, (what, t, rel)


def _draws(T, shape, seed):
  import torch
  g = torch.Generator(device="cuda").manual_seed(seed)
  r0 = torch.randn((T,) + tuple(shape), device="cuda", generator=g)
  u = torch.rand((T,) + tuple(shape[:-1]), device="cuda", generator=g).clamp(1e-6, 1 - 1e-6)
  return r0, u


def _single_chain_case(N, D, T, L, what):
  import torch
  from edward_b200 import engine
  X, y = _gen(N, D)
  s = engine.GLMSampler(engine.GLMSpec(D), X, y)
  g = torch.Generator(device="cuda").manual_seed(3)
  for th in (torch.zeros(D, device="cuda"), torch.randn(D, device="cuda", generator=g) / D ** 0.5):
    lp, gr = s.logp_grad(th)
    lp64, g64 = _logp_grad64(X, y, th)
    assert abs(float(lp[0]) - float(lp64)) <= REL_LOGP * abs(float(lp64)), what
    assert float((gr.double() - g64).abs().max() / g64.abs().max()) <= REL_GRAD, what
  eps = 0.1 / np.sqrt(N / 4.0)  # a tenth of the posterior scale: real dynamics, not a near-zero displacement
  r0, u = _draws(T, (D,), 17)
  params = torch.zeros(T, D, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=r0, u=u)
  want, acc, margins = _hmc64(X, y, torch.zeros(D, device="cuda"), r0, u, eps, L)
  _check_rows(params, want, (sc[:, 6] > 0.5).tolist(), acc, margins, what)
  s.close()


def test_cfg2_full_size_trajectory():
  _single_chain_case(581012, 54, 3, 10, "cfg2")


def test_cfg4_shard_full_size_gradient_and_trajectory():
  _single_chain_case(1250000, 1000, 2, 3, "cfg4 shard")


def _chains_case(N, D, C, T, L, chains, what):
  import torch
  from edward_b200 import engine
  X, y = _gen(N, D)
  s = engine.GLMSampler(engine.GLMSpec(D), X, y, n_chains=C)
  g = torch.Generator(device="cuda").manual_seed(4)
  theta = torch.randn(C, D, device="cuda", generator=g) / D ** 0.5
  theta[0] = 0.0
  lp, gr = s.logp_grad_chains(theta)
  for c in chains:
    lp64, g64 = _logp_grad64(X, y, theta[c])
    assert abs(float(lp[c]) - float(lp64)) <= REL_LOGP * abs(float(lp64)), (what, c)
    assert float((gr[c].double() - g64).abs().max() / g64.abs().max()) <= REL_GRAD, (what, c)
  eps = 0.1 / np.sqrt(N / 4.0)
  r0, u = _draws(T, (C, D), 23)
  params = torch.zeros(T, C, D, device="cuda")
  tr = s.set_chain_trace(T)
  s.run_chains(params, 0, T, eps, L, r0=r0, u=u)
  for c in chains:
    want, acc, margins = _hmc64(X, y, torch.zeros(D, device="cuda"), r0[:, c], u[:, c], eps, L)
    _check_rows(params[:, c], want, (tr[:, c, 6] > 0.5).tolist(), acc, margins, (what, c))
  s.close()


def test_cfg3_full_size_chains():
  _chains_case(581012, 54, 256, 2, 10, (0, 127, 255), "cfg3")


def test_cfg5_shard_full_size_chains():
  _chains_case(1250000, 1000, 1024, 1, 2, (0, 511, 1023), "cfg5 shard")
