"""C vectorised chains (extension): the tcgen05 / 3xTF32 contraction path and its CUDA-core twin, each chain
checked against an independent oracle run of that chain (north_star: "checked against independent reference
runs per chain") and against the single-chain CUDA path."""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu

REL_LOGP = 1e-5
REL_GRAD = 1e-5
REL_POS = 1e-4
TIE_EPS = 1e-3


def _data(N, D, seed):
  rng = np.random.default_rng(seed)
  X = rng.standard_normal((N, D)).astype(np.float32)
  wt = (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32)
  y = (rng.random(N) < 1 / (1 + np.exp(-(X.astype(np.float64) @ wt)))).astype(np.int32)
  return X, y


def _sampler(X, y, D, C, impl, monkeypatch):
  from edward_b200 import engine
  monkeypatch.setenv("EDHMC_MC_IMPL", impl)
  return engine.GLMSampler(engine.GLMSpec(D), X, y, n_chains=C)


@pytest.mark.parametrize("impl", ["simple", "tc"])
@pytest.mark.parametrize("N,D,C", [(1000, 54, 128), (4133, 54, 256), (257, 8, 128), (3000, 64, 128), (20000, 33, 128)])
def test_chain_logp_grad_matches_oracle(N, D, C, impl, monkeypatch):
  X, y = _data(N, D, N + D)
  s = _sampler(X, y, D, C, impl, monkeypatch)
  rng = np.random.default_rng(5)
  theta = (0.3 * rng.standard_normal((C, D)) / np.sqrt(D)).astype(np.float32)
  theta[0] = 0.0
  lp, g = s.logp_grad_chains(theta)
  lp, g = lp.cpu().numpy(), g.cpu().numpy()
  spec = o.GLMSpec(D)
  for c in list(range(0, C, 37)) + [C - 1]:
    lp64 = float(o.log_joint(X, y, theta[c], spec))
    g64 = o.grad_log_joint(X, y, theta[c], spec)
    assert abs(lp[c] - lp64) <= REL_LOGP * abs(lp64), (impl, c, lp[c], lp64)
    rel = np.max(np.abs(g[c] - g64)) / np.max(np.abs(g64))
    assert rel <= REL_GRAD, (impl, c, rel)
  s.close()


@pytest.mark.parametrize("impl", ["simple", "tc"])
def test_chain_run_matches_independent_oracle_runs(impl, monkeypatch):
  import torch
  N, D, C, T, L, eps = 2000, 54, 128, 6, 5, 0.02
  X, y = _data(N, D, 11)
  s = _sampler(X, y, D, C, impl, monkeypatch)
  rng = np.random.Generator(np.random.Philox(key=99))
  r0 = rng.standard_normal((T, C, D), dtype=np.float32)
  u = np.clip(rng.random((T, C), dtype=np.float32), 1e-7, 1 - 1e-7).astype(np.float32)
  z0 = (0.05 * rng.standard_normal((C, D))).astype(np.float32)
  params = torch.zeros(T, C, D, device="cuda")
  params[0] = torch.tensor(z0)
  tr = s.set_chain_trace(T)
  s.run_chains(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_acc, logp = s.read_chain_state()
  got = params.cpu().numpy()
  tr = tr.cpu().numpy()
  spec = o.GLMSpec(D)
  ties = 0
  for c in [0, 1, 17, 63, 64, 100, 127]:
    p64 = np.zeros((T, D))
    p64[0] = z0[c]
    infos, nacc = o.run(X, y, p64, r0[:, c], u[:, c], eps, L, spec)
    forked = False
    for i, info in enumerate(infos):
      assert abs(tr[i, c, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (impl, c, i)
      if bool(tr[i, c, 6] > 0.5) != info.accept:
        assert info.margin < TIE_EPS, (impl, c, i, info)
        ties += 1
        forked = True
        break
      assert np.max(np.abs(got[i, c] - p64[i])) <= REL_POS * max(np.max(np.abs(p64[i])), 1e-3), (impl, c, i)
    if not forked:
      assert n_acc[c] == nacc
  assert ties <= 1
  s.close()


def test_tc_and_simple_paths_agree_and_match_single_chain_kernel(monkeypatch):
  """The tensor-core path, the CUDA-core path and the single-chain streaming kernel evaluate the same log
  joint / gradient for the same theta."""
  import torch
  from edward_b200 import engine
  N, D, C = 9000, 54, 128
  X, y = _data(N, D, 3)
  theta = (0.2 * np.random.default_rng(1).standard_normal((C, D)) / np.sqrt(D)).astype(np.float32)
  outs = {}
  for impl in ("simple", "tc"):
    s = _sampler(X, y, D, C, impl, monkeypatch)
    lp, g = s.logp_grad_chains(theta)
    outs[impl] = (lp.cpu().numpy(), g.cpu().numpy())
    s.close()
  assert np.max(np.abs(outs["tc"][0] - outs["simple"][0]) / np.abs(outs["simple"][0])) < 2e-6
  assert np.max(np.abs(outs["tc"][1] - outs["simple"][1])) / np.max(np.abs(outs["simple"][1])) < 5e-6
  single = engine.GLMSampler(engine.GLMSpec(D), X, y)
  for c in (0, 77):
    lp1, g1 = single.logp_grad(theta[c])
    assert abs(float(lp1[0]) - outs["tc"][0][c]) <= 2e-6 * abs(float(lp1[0]))
    assert np.max(np.abs(g1.cpu().numpy() - outs["tc"][1][c])) <= 5e-6 * np.max(np.abs(outs["tc"][1][c]))
  single.close()


def test_chain_argument_checks():
  from edward_b200 import _C, engine
  X, y = _data(100, 8, 0)
  s = engine.GLMSampler(engine.GLMSpec(8), X, y, n_chains=100)   # any chain count: the kernels pad to whole 128-chain tiles
  with pytest.raises(TypeError):
    s.run_chains(__import__("torch").zeros(2, 128, 8, device="cuda"), 0, 2, 0.01, 2)  # params must be [T, 100, 8]
  s.close()


# ---- wide models / the two-GEMM path (chains_wide.cu): n_features > 64, or forced with EDHMC_MC_IMPL=wide ----
@pytest.mark.parametrize("N,D,C,impl", [(1000, 54, 128, "wide"), (5000, 200, 128, "tc"), (2111, 1000, 256, "tc"),
                                        (20000, 72, 128, "tc"), (300, 1000, 128, "tc")])
def test_wide_chain_logp_grad_matches_oracle(N, D, C, impl, monkeypatch):
  X, y = _data(N, D, N + D)
  s = _sampler(X, y, D, C, impl, monkeypatch)
  rng = np.random.default_rng(5)
  theta = (0.3 * rng.standard_normal((C, D)) / np.sqrt(D)).astype(np.float32)
  theta[0] = 0.0
  lp, g = s.logp_grad_chains(theta)
  lp, g = lp.cpu().numpy(), g.cpu().numpy()
  spec = o.GLMSpec(D)
  for c in list(range(0, C, 37)) + [C - 1]:
    lp64 = float(o.log_joint(X, y, theta[c], spec))
    g64 = o.grad_log_joint(X, y, theta[c], spec)
    assert abs(lp[c] - lp64) <= REL_LOGP * abs(lp64), (c, lp[c], lp64)
    rel = np.max(np.abs(g[c] - g64)) / np.max(np.abs(g64))
    assert rel <= REL_GRAD, (c, rel)
  s.close()


@pytest.mark.parametrize("D,impl", [(54, "wide"), (300, "tc")])
def test_wide_chain_run_matches_independent_oracle_runs(D, impl, monkeypatch):
  import torch
  N, C, T, L, eps = 2000, 128, 5, 4, 0.02
  X, y = _data(N, D, 11)
  s = _sampler(X, y, D, C, impl, monkeypatch)
  rng = np.random.Generator(np.random.Philox(key=99))
  r0 = rng.standard_normal((T, C, D), dtype=np.float32)
  u = np.clip(rng.random((T, C), dtype=np.float32), 1e-7, 1 - 1e-7).astype(np.float32)
  z0 = (0.05 * rng.standard_normal((C, D))).astype(np.float32)
  params = torch.zeros(T, C, D, device="cuda")
  params[0] = torch.tensor(z0)
  tr = s.set_chain_trace(T)
  s.run_chains(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_acc, logp = s.read_chain_state()
  got = params.cpu().numpy()
  tr = tr.cpu().numpy()
  spec = o.GLMSpec(D)
  ties = 0
  for c in [0, 1, 64, 127]:
    p64 = np.zeros((T, D))
    p64[0] = z0[c]
    infos, nacc = o.run(X, y, p64, r0[:, c], u[:, c], eps, L, spec)
    forked = False
    for i, info in enumerate(infos):
      assert abs(tr[i, c, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (c, i)
      if bool(tr[i, c, 6] > 0.5) != info.accept:
        assert info.margin < TIE_EPS, (c, i, info)
        ties += 1
        forked = True
        break
      assert np.max(np.abs(got[i, c] - p64[i])) <= REL_POS * max(np.max(np.abs(p64[i])), 1e-3), (c, i)
    if not forked:
      assert n_acc[c] == nacc
  assert ties <= 1
  s.close()


# ---- bias latent: one more column (of ones) of the pre-tiled operands, so both tensor-core paths and the CUDA-core twin
#      take models with an intercept (cfg 1's shape: Bernoulli(logits=ed.dot(X, w) + b), examples/...:60-62) ----
@pytest.mark.parametrize("N,D,C,impl", [(1000, 7, 128, "simple"), (1000, 7, 128, "tc"), (4133, 53, 256, "tc"), (3000, 63, 128, "tc"),
                                        (900, 64, 128, "tc"), (2111, 200, 128, "tc")])
def test_chain_logp_grad_with_bias_latent(N, D, C, impl, monkeypatch):
  from edward_b200 import engine
  X, y = _data(N, D, N + D)
  monkeypatch.setenv("EDHMC_MC_IMPL", impl)
  P = D + 1
  ps = np.full(P, 1.5, np.float32)
  s = engine.GLMSampler(engine.GLMSpec(D, True, prior_scale=ps), X, y, n_chains=C)
  rng = np.random.default_rng(5)
  theta = (0.3 * rng.standard_normal((C, P)) / np.sqrt(D)).astype(np.float32)
  lp, g = s.logp_grad_chains(theta)
  lp, g = lp.cpu().numpy(), g.cpu().numpy()
  spec = o.GLMSpec(D, True, prior_scale=ps)
  for c in list(range(0, C, 41)) + [C - 1]:
    lp64 = float(o.log_joint(X, y, theta[c], spec))
    g64 = o.grad_log_joint(X, y, theta[c], spec)
    assert abs(lp[c] - lp64) <= REL_LOGP * abs(lp64), (impl, c, lp[c], lp64)
    assert np.max(np.abs(g[c] - g64)) / np.max(np.abs(g64)) <= REL_GRAD, (impl, c)
  s.close()


@pytest.mark.parametrize("D,impl", [(20, "tc"), (100, "tc")])
def test_chain_run_with_bias_latent_matches_independent_oracle_runs(D, impl, monkeypatch):
  import torch
  from edward_b200 import engine
  N, C, T, L, eps = 2000, 128, 5, 4, 0.02
  P = D + 1
  X, y = _data(N, D, 13)
  monkeypatch.setenv("EDHMC_MC_IMPL", impl)
  s = engine.GLMSampler(engine.GLMSpec(D, True), X, y, n_chains=C)
  rng = np.random.Generator(np.random.Philox(key=7))
  r0 = rng.standard_normal((T, C, P), dtype=np.float32)
  u = np.clip(rng.random((T, C), dtype=np.float32), 1e-7, 1 - 1e-7).astype(np.float32)
  params = torch.zeros(T, C, P, device="cuda")
  tr = s.set_chain_trace(T)
  s.run_chains(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_acc, _ = s.read_chain_state()
  got, tr = params.cpu().numpy(), tr.cpu().numpy()
  spec = o.GLMSpec(D, True)
  ties = 0
  for c in [0, 5, 64, 127]:
    p64 = np.zeros((T, P))
    infos, nacc = o.run(X, y, p64, r0[:, c], u[:, c], eps, L, spec)
    forked = False
    for i, info in enumerate(infos):
      assert abs(tr[i, c, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (c, i)
      if bool(tr[i, c, 6] > 0.5) != info.accept:
        assert info.margin < TIE_EPS, (c, i, info)
        ties += 1
        forked = True
        break
      assert np.max(np.abs(got[i, c] - p64[i])) <= REL_POS * max(np.max(np.abs(p64[i])), 1e-3), (c, i)
    if not forked:
      assert n_acc[c] == nacc
  assert ties <= 1
  s.close()


# ---- chain counts that are not multiples of 128: the pass kernels run whole 128-chain tiles, the padding chains live in
#      the handle's own state arrays only; every caller-owned array has exactly C chains ----
@pytest.mark.parametrize("N,D,C,impl", [(1000, 54, 100, "tc"), (1000, 54, 100, "simple"), (3000, 20, 130, "tc"), (900, 8, 2, "tc"),
                                        (1500, 200, 200, "tc"), (1000, 54, 70, "wide")])
def test_chain_counts_not_multiple_of_128(N, D, C, impl, monkeypatch):
  import torch
  X, y = _data(N, D, N + C)
  s = _sampler(X, y, D, C, impl, monkeypatch)
  rng = np.random.Generator(np.random.Philox(key=C))
  theta = (0.3 * rng.standard_normal((C, D)) / np.sqrt(D)).astype(np.float32)
  lp, g = s.logp_grad_chains(theta)
  lp, g = lp.cpu().numpy(), g.cpu().numpy()
  assert lp.shape == (C,) and g.shape == (C, D)
  spec = o.GLMSpec(D)
  for c in sorted({0, C // 2, C - 1}):
    lp64 = float(o.log_joint(X, y, theta[c], spec))
    g64 = o.grad_log_joint(X, y, theta[c], spec)
    assert abs(lp[c] - lp64) <= REL_LOGP * abs(lp64), (c, lp[c], lp64)
    assert np.max(np.abs(g[c] - g64)) / np.max(np.abs(g64)) <= REL_GRAD, c
  T, L, eps = 4, 3, 0.02
  r0 = rng.standard_normal((T, C, D), dtype=np.float32)
  u = np.clip(rng.random((T, C), dtype=np.float32), 1e-7, 1 - 1e-7).astype(np.float32)
  params = torch.zeros(T, C, D, device="cuda")
  tr = s.set_chain_trace(T)
  s.run_chains(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_acc, _ = s.read_chain_state()
  assert len(n_acc) == C
  got, tr = params.cpu().numpy(), tr.cpu().numpy()
  for c in sorted({0, C - 1}):
    p64 = np.zeros((T, D))
    infos, nacc = o.run(X, y, p64, r0[:, c], u[:, c], eps, L, spec)
    forked = False
    for i, info in enumerate(infos):
      assert abs(tr[i, c, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (c, i)
      if bool(tr[i, c, 6] > 0.5) != info.accept:
        assert info.margin < TIE_EPS, (c, i, info)
        forked = True
        break
      assert np.max(np.abs(got[i, c] - p64[i])) <= REL_POS * max(np.max(np.abs(p64[i])), 1e-3), (c, i)
    if not forked:
      assert n_acc[c] == nacc
  s.close()
