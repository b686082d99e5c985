"""Pins the CPU oracle (oracle/hmc_oracle.py, oracle/hmc_ref.c) — CPU only.

The reference holds no golden vectors for this path and TensorFlow cannot run here ("parity unpinned"),
so the oracle is anchored on: independent densities (scipy.stats), finite differences, the reference's
in-tree closed forms (edward/inferences/conjugacy/conjugate_log_probs.py:21-24,134-141), np.dot
(tests/util/dot_test.py), the reference's statistical HMC tests (tests/inferences/hmc_test.py), and the
committed golden fixtures (regression)."""
import glob
import os

import numpy as np
import pytest
from scipy import stats

import hmc_oracle as o
import ref_c

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("ref_"))


def _spec_from(d):
  return o.GLMSpec(d["X"].shape[1], bool(d["has_bias"]), int(d["family"]), d["prior_loc"], d["prior_scale"],
                   float(d["lik_scale"]))


def test_densities_match_scipy():
  rng = np.random.default_rng(0)
  x = rng.standard_normal(100) * 3
  loc, scale = rng.standard_normal(100), rng.random(100) + 0.1
  np.testing.assert_allclose(o.normal_log_prob(x, loc, scale, np.float64), stats.norm.logpdf(x, loc, scale), rtol=1e-12)
  l = rng.standard_normal(200) * 8
  y = rng.integers(0, 2, 200)
  p = 1 / (1 + np.exp(-l))
  np.testing.assert_allclose(o.bernoulli_logit_log_prob(l, y, np.float64), stats.bernoulli.logpmf(y, p), rtol=1e-9, atol=1e-12)
  k = rng.integers(0, 9, 200)
  np.testing.assert_allclose(o.poisson_log_log_prob(l / 4, k, np.float64), stats.poisson.logpmf(k, np.exp(l / 4)), rtol=1e-10)
  # extreme logits stay finite (the stable form of sigmoid_cross_entropy_with_logits)
  assert np.all(np.isfinite(o.bernoulli_logit_log_prob(np.array([-200., 200., 0.]), np.array([1, 0, 1]), np.float32)))


def test_densities_match_reference_closed_forms():
  """conjugate_log_probs.py:21-24: x*log(p) + (1-x)*log1p(-p);  :134-141: expanded-square Normal."""
  rng = np.random.default_rng(1)
  l = rng.standard_normal(50) * 3
  y = rng.integers(0, 2, 50).astype(np.float64)
  p = 1 / (1 + np.exp(-l))
  np.testing.assert_allclose(o.bernoulli_logit_log_prob(l, y, np.float64), y * np.log(p) + (1 - y) * np.log1p(-p), rtol=1e-10)
  x, loc, scale = rng.standard_normal(50), rng.standard_normal(50), rng.random(50) + 0.2
  var = scale ** 2
  closed = -0.5 * np.log(2 * np.pi * var) - 0.5 * loc ** 2 / var + x * loc / var - 0.5 * x ** 2 / var
  np.testing.assert_allclose(o.normal_log_prob(x, loc, scale, np.float64), closed, rtol=1e-9, atol=1e-12)


def test_dot_matches_numpy_and_raises_on_inf():
  """tests/util/dot_test.py:13-31."""
  a = np.arange(25, dtype=np.float32).reshape(5, 5) * 0.1
  b = np.arange(5, dtype=np.float32)
  np.testing.assert_allclose(o.dot(a, b), np.dot(a, b), rtol=1e-6)
  np.testing.assert_allclose(o.dot(b, a), np.dot(b, a), rtol=1e-6)
  a2 = a.copy()
  a2[0, 0] = np.inf
  with pytest.raises(ValueError):
    o.dot(a2, b)
  b2 = b.copy()
  b2[1] = np.inf
  with pytest.raises(ValueError):
    o.dot(a, b2)


@pytest.mark.parametrize("family,bias", [(o.BERNOULLI_LOGIT, True), (o.NORMAL_IDENTITY, True), (o.POISSON_LOG, False)])
def test_gradient_matches_finite_differences(family, bias):
  rng = np.random.default_rng(2)
  N, D = 200, 6
  X = rng.standard_normal((N, D)).astype(np.float32)
  y = rng.integers(0, 2, N) if family != o.NORMAL_IDENTITY else rng.standard_normal(N).astype(np.float32)
  P = D + int(bias)
  spec = o.GLMSpec(D, bias, family, rng.standard_normal(P).astype(np.float32), (rng.random(P) + 0.5).astype(np.float32), 0.8)
  th = 0.3 * rng.standard_normal(P)
  g = o.grad_log_joint(X, y, th, spec, np.float64)
  h = 1e-6
  for i in range(P):
    e = np.zeros(P)
    e[i] = h
    fd = (o.log_joint(X, y, th + e, spec) - o.log_joint(X, y, th - e, spec)) / (2 * h)
    assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_leapfrog_is_reversible_and_conserves_energy():
  X, y, _ = o.synth_data(500, 8)
  spec = o.GLMSpec(8)
  z0 = np.zeros(8)
  r0 = np.random.default_rng(3).standard_normal(8)
  z1, r1 = o.leapfrog(X, y, z0, r0, 0.005, 20, spec)
  zb, rb = o.leapfrog(X, y, z1, -r1, 0.005, 20, spec)
  np.testing.assert_allclose(zb, z0, atol=1e-10)
  np.testing.assert_allclose(-rb, r0, atol=1e-10)
  h0 = -o.log_joint(X, y, z0, spec) + 0.5 * r0 @ r0
  h1 = -o.log_joint(X, y, z1, spec) + 0.5 * r1 @ r1
  assert abs(h1 - h0) < 0.05


def test_oracle_samples_reference_posterior_normal_normal():
  """tests/inferences/hmc_test.py:14-46 on the oracle: posterior N(0, 1/sqrt(51)); reference tolerances."""
  N, T = 50, 2000
  X = np.ones((N, 1), np.float32)
  y = np.zeros(N, np.float32)
  spec = o.GLMSpec(1, False, o.NORMAL_IDENTITY, np.zeros(1, np.float32), np.ones(1, np.float32), 1.0)
  r0, u = o.synth_draws(T, 1, seed=11)
  params = np.ones((T, 1))
  infos, nacc = o.run(X, y, params, r0, u, 0.25, 2, spec)
  np.testing.assert_allclose(o.empirical_mean(params), 0, rtol=1e-1, atol=1e-1)
  np.testing.assert_allclose(o.empirical_stddev(params), np.sqrt(1 / 51), rtol=1e-1, atol=1e-1)
  assert nacc > 0.1


def test_oracle_samples_reference_posterior_linear_regression():
  """tests/inferences/hmc_test.py:48-91 on the oracle (N=40, D=10, scale 0.1, step_size 0.01)."""
  rng = np.random.RandomState(0)
  N, D, T = 40, 10, 2000
  w_true = rng.randn(D)
  X = rng.randn(N, D).astype(np.float32)
  y = (X @ w_true + rng.normal(0, 0.1, size=N)).astype(np.float32)
  spec = o.GLMSpec(D, True, o.NORMAL_IDENTITY, np.zeros(D + 1, np.float32), np.ones(D + 1, np.float32), 0.1)
  r0, u = o.synth_draws(T, D + 1, seed=12)
  params = np.zeros((T, D + 1))
  infos, nacc = o.run(X, y, params, r0, u, 0.01, 2, spec)
  np.testing.assert_allclose(o.empirical_mean(params)[:D], w_true, rtol=5e-1, atol=5e-1)
  np.testing.assert_allclose(o.empirical_mean(params)[D:], [0.0], rtol=5e-1, atol=5e-1)
  assert nacc > 0.1


def test_transition_row_semantics():
  """Transition t reads row max(t-1,0), writes row t (hmc.py:81-85,121-126); running past T raises."""
  X, y, _ = o.synth_data(100, 3)
  spec = o.GLMSpec(3)
  r0, u = o.synth_draws(4, 3)
  params = np.zeros((3, 3))
  params[0] = [0.1, -0.2, 0.3]
  start = params[0].copy()
  info = o.transition(X, y, params, 0, r0[0], u[0], 0.05, 2, spec)
  assert np.allclose(params[0], info.proposal if info.accept else start)
  with pytest.raises(IndexError):
    o.run(X, y, params, r0, u, 0.05, 2, spec, t0=0, n_iter=4)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
  d = np.load(path)
  spec = _spec_from(d)
  X, y, T, L, eps = d["X"], d["y"], int(d["T"]), int(d["L"]), float(d["eps"])
  for tag, dt, tol in (("f64", np.float64, 1e-12), ("f32", np.float32, 2e-5)):
    np.testing.assert_allclose(o.log_joint(X, y, d["theta"], spec, dt), d["logp_theta_" + tag], rtol=tol)
    g = o.grad_log_joint(X, y, d["theta"], spec, dt)
    assert np.max(np.abs(g - d["grad_theta_" + tag])) <= tol * np.max(np.abs(g)) + 1e-12
  params = np.zeros((T, spec.n_params))
  infos, nacc = o.run(X, y, params, d["r0"], d["u"], eps, L, spec, np.float64)
  np.testing.assert_allclose(params, d["params_f64"], rtol=1e-9, atol=1e-12)
  assert nacc == int(d["n_accept_f64"])
  np.testing.assert_array_equal([float(i.accept) for i in infos], d["trace_f64"][:, 6])


@pytest.mark.parametrize("path", [p for p in GOLDEN if "poisson" not in p],
                         ids=[os.path.basename(p)[:-4] for p in GOLDEN if "poisson" not in p])
def test_c_port_matches_golden(path):
  """oracle/hmc_ref.c (float32, reference schedule) against the float64 golden values."""
  d = np.load(path)
  spec = _spec_from(d)
  X, y, T, L, eps = d["X"], d["y"], int(d["T"]), int(d["L"]), float(d["eps"])
  lp, g = ref_c.logp_grad(X, y, d["theta"], spec)
  assert abs(lp - float(d["logp_theta_f64"])) <= 2e-5 * abs(float(d["logp_theta_f64"]))
  assert np.max(np.abs(g - d["grad_theta_f64"])) <= 2e-5 * np.max(np.abs(d["grad_theta_f64"]))
  params = np.zeros((T, spec.n_params), np.float32)
  nacc, tr = ref_c.run(X, y, params, d["r0"], d["u"], eps, L, spec)
  t64 = d["trace_f64"]
  margin = np.abs(t64[:, 5] - t64[:, 4])
  same = tr[:, 6] == t64[:, 6]
  assert np.all(same | (margin < 1e-2)), (tr[:, 6], t64[:, 6], margin)
  if np.all(same):
    assert np.max(np.abs(params - d["params_f64"])) <= 2e-4 * max(np.max(np.abs(d["params_f64"])), 1e-3)
    assert nacc == int(d["n_accept_f64"])


def test_c_port_checknumerics_and_range():
  X, y, _ = o.synth_data(64, 4)
  spec = o.GLMSpec(4)
  r0, u = o.synth_draws(2, 4)
  Xb = X.copy()
  Xb[5, 2] = np.inf
  with pytest.raises(ValueError):
    ref_c.run(Xb, y, np.zeros((2, 4), np.float32), r0, u, 0.1, 2, spec)
  with pytest.raises(IndexError):
    ref_c.run(X, y, np.zeros((1, 4), np.float32), r0, u, 0.1, 2, spec, n_iter=2)


def test_synthetic_data_is_shard_invariant():
  """rows generated in 65,536-row blocks seeded by block index → identical under any
  block-aligned sharding."""
  N, D = 65536 * 2 + 1000, 5
  X, y, w = o.synth_data(N, D)
  X2, y2, _ = o.synth_data(N - 65536, D, row_start=65536)
  assert np.array_equal(X[65536:], X2) and np.array_equal(y[65536:], y2)


def test_beta_logit_prior_matches_autodiff_of_the_literal_tf_expression():
  """A Beta(a, b) latent moved to the real line by ed.transform (util/random_variables.py:895-897): the oracle's closed
  form a log sigmoid(u) + b log sigmoid(-u) - lbeta against the literal composition [TF 1.5] Beta._log_prob(sigmoid(u)) +
  Sigmoid.forward_log_det_jacobian(u) = -softplus(-u) - softplus(u), and its gradient against torch.autograd."""
  import torch
  from scipy.special import betaln
  for a, b in ((1.0, 1.0), (4.0, 8.0), (0.5, 2.5)):
    u = torch.linspace(-6, 6, 25, dtype=torch.float64, requires_grad=True)
    z = torch.sigmoid(u)
    lp = (a - 1.0) * torch.log(z) + (b - 1.0) * torch.log1p(-z) - float(betaln(a, b))
    lp = lp + (-torch.nn.functional.softplus(-u) - torch.nn.functional.softplus(u))
    (g,) = torch.autograd.grad(lp.sum(), u)
    un = u.detach().numpy()
    np.testing.assert_allclose(o.beta_logit_log_prob(un, a, b, np.float64), lp.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(o.beta_logit_log_prob_grad(un, a, b, np.float64), g.numpy(), rtol=1e-12, atol=1e-12)
    # it is a density on the real line: integrates to one
    grid = np.linspace(-40, 40, 400001)
    assert abs(np.trapezoid(np.exp(o.beta_logit_log_prob(grid, a, b, np.float64)), grid) - 1.0) < 1e-6


def test_oracle_beta_bernoulli_posterior_through_the_sigmoid():
  """tests/inferences/inference_auto_transform_test.py:163-188 on the oracle: z ~ Beta(1,1), 10 Bernoulli(probs=z)
  observations, HMC in logit space; the mapped-back samples have the moments of the exact Beta(4, 8) posterior."""
  x_obs = np.asarray([0, 0, 1, 1, 0, 0, 0, 0, 0, 1], np.int32)
  spec = o.GLMSpec(1, False, o.BERNOULLI_LOGIT, np.ones(1, np.float32), np.ones(1, np.float32), 1.0, prior_kind=np.array([1]))
  X = np.ones((10, 1), np.float32)
  T = 3000
  r0, u = o.synth_draws(T, 1, seed=5)
  params = np.zeros((T, 1))
  infos, nacc = o.run(X, x_obs, params, r0, u, 1.0, 5, spec)
  z = 1 / (1 + np.exp(-params[500:, 0]))
  a, b = 1.0 + x_obs.sum(), 1.0 + (1 - x_obs).sum()
  assert abs(z.mean() - a / (a + b)) < 0.02
  assert abs(z.var() - a * b / ((a + b) ** 2 * (a + b + 1))) < 0.004
  assert nacc > 0.5 * T


def test_predictive_oracle_against_scipy():
  """oracle.predictive (the restatement of evaluate.py:132-143,158-162,222-227) against scipy.stats densities."""
  import scipy.stats as st
  rng = np.random.default_rng(11)
  N, D, S = 40, 5, 7
  X = rng.standard_normal((N, D))
  W = rng.standard_normal((S, D)) / 2
  B = rng.standard_normal(S) / 3
  eta = X @ W.T + B[None, :]
  yb = (rng.random(N) < 0.5).astype(np.int32)
  m, l = o.predictive(X, yb, W, B, o.BERNOULLI_LOGIT)
  p = 1 / (1 + np.exp(-eta))
  assert np.allclose(m, p.mean(axis=1), rtol=1e-12)
  assert np.allclose(l, st.bernoulli.logpmf(yb[:, None], p).sum(axis=1), rtol=1e-10)
  yn = rng.standard_normal(N)
  m, l = o.predictive(X, yn, W, B, o.NORMAL_IDENTITY, 0.7)
  assert np.allclose(m, eta.mean(axis=1), rtol=1e-12)
  assert np.allclose(l, st.norm.logpdf(yn[:, None], eta, 0.7).sum(axis=1), rtol=1e-10)
  yp = rng.poisson(1.0, N)
  m, l = o.predictive(X, yp, W, None, o.POISSON_LOG)
  eta0 = X @ W.T
  assert np.allclose(m, np.exp(eta0).mean(axis=1), rtol=1e-12)
  assert np.allclose(l, st.poisson.logpmf(yp[:, None], np.exp(eta0)).sum(axis=1), rtol=1e-10)
