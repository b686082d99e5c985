"""Constrained latents under auto_transform: the reference's own tests
tests/inferences/inference_auto_transform_test.py:114-188 re-expressed against edward_b200 (same models, settings and
tolerances), and engine-level parity of the Beta-through-sigmoid prior against the oracle with injected draws."""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fresh_graph():
  from edward_b200 import graph as g
  g.reset_default_graph()
  yield


def _imports():
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Beta, Empirical, Normal, TransformedDistribution
  return ed, tf, Bernoulli, Beta, Empirical, Normal, TransformedDistribution


def test_beta_prior_engine_matches_oracle():
  """log joint, gradient, proposals and accept decisions of the Beta-Bernoulli model in logit space."""
  import torch
  from edward_b200 import engine
  x_obs = np.asarray([0, 0, 1, 1, 0, 0, 0, 0, 0, 1], np.int32)
  spec = o.GLMSpec(1, False, o.BERNOULLI_LOGIT, np.full(1, 2.0, np.float32), np.full(1, 3.5, np.float32), 1.0,
                   prior_kind=np.array([1]))
  X = np.ones((10, 1), np.float32)
  s = engine.GLMSampler(engine.GLMSpec(1, False, 0, spec.prior_loc, spec.prior_scale), X, x_obs)
  s.set_prior_kinds([1])
  for u in (-2.0, 0.0, 0.7, 3.0):
    lp, g = s.logp_grad(np.array([u], np.float32))
    lp64 = float(o.log_joint(X, x_obs, np.array([u]), spec))
    g64 = o.grad_log_joint(X, x_obs, np.array([u]), spec)
    assert abs(float(lp[0]) - lp64) <= 1e-5 * abs(lp64), (u, float(lp[0]), lp64)
    assert abs(float(g[0]) - g64[0]) <= 1e-5 * max(abs(g64[0]), 1.0), (u, float(g[0]), g64[0])
  T, L, eps = 60, 5, 0.7
  r0, uu = o.synth_draws(T, 1, seed=3)
  p64 = np.zeros((T, 1))
  infos, nacc = o.run(X, x_obs, p64, r0, uu, eps, L, spec)
  for plan in (1, 2):
    s2 = engine.GLMSampler(engine.GLMSpec(1, False, 0, spec.prior_loc, spec.prior_scale), X, x_obs, plan=plan)
    s2.set_prior_kinds([1])
    params = torch.zeros(T, 1, device="cuda")
    sc, pos = s2.set_trace(T)
    s2.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(uu))
    sc = sc.cpu().numpy()
    ties = [i for i in infos if i.margin < 1e-3]
    if not ties:
      np.testing.assert_array_equal(sc[:, 6] > 0.5, np.array([i.accept for i in infos]))
      assert np.max(np.abs(params.cpu().numpy() - p64)) <= 1e-4 * max(np.max(np.abs(p64)), 1e-6)
      assert s2.read_state()[0] == nacc
    s2.close()
  s.close()


def test_hmc_custom():
  """inference_auto_transform_test.py:114-137."""
  ed, tf, Bernoulli, Beta, Empirical, Normal, TransformedDistribution = _imports()
  sess = ed.get_session()
  x = TransformedDistribution(distribution=Normal(1.0, 1.0), bijector=tf.contrib.distributions.bijectors.Softplus())
  x.support = 'nonnegative'
  qx = Empirical(tf.Variable(tf.random_normal([1000])))
  inference = ed.HMC({x: qx})
  inference.initialize(auto_transform=True, step_size=0.8)
  tf.global_variables_initializer().run()
  for _ in range(inference.n_iter):
    inference.update()
  n_samples = 10000
  x_unconstrained = inference.transformations[x]
  qx_constrained_params = x_unconstrained.bijector.inverse(qx.params)
  x_mean, x_var = tf.nn.moments(x.sample(n_samples), 0)
  qx_mean, qx_var = tf.nn.moments(qx_constrained_params[500:], 0)
  stats = sess.run([x_mean, qx_mean, x_var, qx_var])
  np.testing.assert_allclose(stats[0], stats[1], rtol=1e-1, atol=1e-1)
  np.testing.assert_allclose(stats[2], stats[3], rtol=1e-1, atol=1e-1)


def test_hmc_default():
  """inference_auto_transform_test.py:139-161: ed.HMC([x]) builds the Empirical; latent_vars[x] is the constrained view."""
  ed, tf, Bernoulli, Beta, Empirical, Normal, TransformedDistribution = _imports()
  sess = ed.get_session()
  x = TransformedDistribution(distribution=Normal(1.0, 1.0), bijector=tf.contrib.distributions.bijectors.Softplus())
  x.support = 'nonnegative'
  inference = ed.HMC([x])
  inference.initialize(auto_transform=True, step_size=0.8, n_print=0)
  tf.global_variables_initializer().run()
  for _ in range(3000):   # the default store has 10,000 rows; the moments settle long before
    inference.update()
  qx_constrained = inference.latent_vars[x]
  x_mean, x_var = tf.nn.moments(x.sample(1000), 0)
  qx_mean, qx_var = tf.nn.moments(qx_constrained.params[500:3000], 0)
  stats = sess.run([x_mean, qx_mean, x_var, qx_var])
  np.testing.assert_allclose(stats[0], stats[1], rtol=1e-1, atol=1e-1)
  np.testing.assert_allclose(stats[2], stats[3], rtol=1e-1, atol=1e-1)


def test_hmc_betabernoulli():
  """inference_auto_transform_test.py:163-188: do we correctly handle dependencies of transformed variables?"""
  ed, tf, Bernoulli, Beta, Empirical, Normal, TransformedDistribution = _imports()
  sess = ed.get_session()
  z = Beta(1., 1., name="z")
  xs = Bernoulli(probs=z, sample_shape=10)
  x_obs = np.asarray([0, 0, 1, 1, 0, 0, 0, 0, 0, 1], dtype=np.int32)
  qz_samples = tf.Variable(tf.random_uniform(shape=(1000,)))
  qz = ed.models.Empirical(params=qz_samples, name="z_posterior")
  inference_hmc = ed.inferences.HMC({z: qz}, data={xs: x_obs})
  inference_hmc.run(step_size=1.0, n_steps=5, auto_transform=True, n_print=0)
  z_unconstrained = inference_hmc.transformations[z]
  qz_constrained = z_unconstrained.bijector.inverse(qz_samples)
  qz_mean, qz_var = sess.run(tf.nn.moments(qz_constrained, 0))
  true_posterior = Beta(1. + np.sum(x_obs), 1. + np.sum(1 - x_obs))
  pz_mean, pz_var = sess.run((true_posterior.mean(), true_posterior.variance()))
  np.testing.assert_allclose(qz_mean, pz_mean, rtol=5e-2, atol=5e-2)
  np.testing.assert_allclose(qz_var, pz_var, rtol=1e-2, atol=1e-2)
