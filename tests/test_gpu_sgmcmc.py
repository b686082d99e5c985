"""SGLD / SGHMC on the fused gradient kernel: parity against the oracle's restatement of
sgld.py:52-87 / sghmc.py:58-96 with injected noise, mini-batches and `scale`, plus the reference's own
statistical tests (tests/inferences/sgld_test.py:14-46, sghmc_test.py)."""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fresh_graph():
  from edward_b200 import graph as g
  g.reset_default_graph()
  yield


def _sampler(X, y, spec):
  from edward_b200 import engine
  return engine.GLMSampler(engine.GLMSpec(spec.n_features, spec.has_bias, spec.family, spec.prior_loc, spec.prior_scale,
                                          spec.lik_scale), X, y)


@pytest.mark.parametrize("kind", ["sgld", "sghmc"])
@pytest.mark.parametrize("batch", [0, 256])
def test_sgmcmc_matches_oracle(kind, batch):
  import torch
  N, D, T = 2048, 20, 40
  X, y, _ = o.synth_data(N, D)
  spec = o.GLMSpec(D, True, o.BERNOULLI_LOGIT, np.zeros(D + 1, np.float32), np.full(D + 1, 2.0, np.float32))
  P = D + 1
  noise = np.random.default_rng(1).standard_normal((T, P)).astype(np.float32)
  lik = N / batch if batch else 1.0
  pf = np.full(P, 0.5, np.float32)
  p64 = np.zeros((T, P))
  p64[0] = 0.01
  s = _sampler(X, y, spec)
  params = torch.zeros(T, P, device="cuda")
  params[0] = 0.01
  if kind == "sgld":
    o.sgld_run(X, y, p64, noise, 1e-3, spec, lik_factor=lik, prior_factor=pf, batch_rows=batch)
    s.sgmcmc_run("sgld", params, 0, T, 1e-3, lik_factor=lik, prior_factor=pf, noise=torch.tensor(noise), batch_rows=batch)
  else:
    v64 = o.sghmc_run(X, y, p64, noise, 1e-2, 0.2, spec, lik_factor=lik, prior_factor=pf, batch_rows=batch)
    vel = torch.zeros(P, device="cuda")
    s.sgmcmc_run("sghmc", params, 0, T, 1e-2, friction=0.2, lik_factor=lik, prior_factor=pf, velocity=vel,
                 noise=torch.tensor(noise), batch_rows=batch)
    assert np.max(np.abs(vel.cpu().numpy() - v64)) <= 1e-4 * max(np.max(np.abs(v64)), 1e-3)
  got = params.cpu().numpy()
  assert np.max(np.abs(got - p64)) <= 1e-4 * np.max(np.abs(p64))
  assert s.read_state()[0] == T
  s.close()


@pytest.mark.parametrize("cls,step", [("SGLD", 0.10), ("SGHMC", 0.025)])
def test_reference_statistical_test_normal_normal(cls, step):
  """sgld_test.py:14-46 / sghmc_test.py: 50 zeros, posterior N(0, 1/sqrt(51)); reference tolerances."""
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Empirical, Normal
  sess = ed.get_session()
  x_data = np.array([0.0] * 50, dtype=np.float32)
  mu = Normal(loc=tf.constant(0.0), scale=tf.constant(1.0))
  x = Normal(loc=mu, scale=tf.constant(1.0), sample_shape=50)
  n_samples = 2000
  qmu = Empirical(params=tf.Variable(tf.ones(n_samples)))
  inference = getattr(ed, cls)({mu: qmu}, data={x: x_data})
  inference.run(step_size=step, n_print=0)
  np.testing.assert_allclose(qmu.mean().eval(), 0, rtol=1e-1, atol=1e-1)
  np.testing.assert_allclose(qmu.stddev().eval(), np.sqrt(1 / 51), rtol=1.5e-1, atol=1.5e-1)
  old_t, old_n_accept = sess.run([inference.t, inference.n_accept])
  assert old_t == n_samples and old_n_accept > 0.1
  sess.run(inference.reset)
  assert sess.run([inference.t, inference.n_accept]) == [0, 0]


def test_sgld_minibatch_front_end_with_scale():
  """Mini-batch SGLD with scale={y: N/B} through the front-end recovers the logistic-regression weights."""
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Empirical, Normal
  N, D, T, B = 20000, 8, 600, 2000
  X_train, y_train, w_true = o.synth_data(N, D)
  X = tf.placeholder(tf.float32, [N, D])
  w = Normal(loc=tf.zeros(D), scale=tf.ones(D))
  y = Bernoulli(logits=ed.dot(X, w))
  qw = Empirical(params=tf.Variable(tf.zeros([T, D])))
  inference = ed.SGLD({w: qw}, data={X: X_train, y: y_train})
  inference.run(step_size=1e-3, n_print=0, scale={y: float(N) / B}, batch_size=B)
  est = qw.params.eval()[T // 2:].mean(axis=0)
  spec = o.GLMSpec(D)
  z = np.zeros(D)
  for _ in range(200):  # posterior mode by gradient ascent on the oracle
    z = z + 2e-4 * o.grad_log_joint(X_train, y_train, z, spec)
  assert np.max(np.abs(est - z)) < 0.15, (est, z)


def test_sghmc_checkpoint_resume_is_bit_identical():
  """the SGHMC velocity is part of the resumable state."""
  import edward_b200 as ed
  from edward_b200 import graph as g
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Empirical, Normal

  def build(T):
    g.reset_default_graph()
    ed.set_seed(5)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((256, 5)).astype(np.float32)
    y = (rng.random(256) < 0.5).astype(np.int32)
    xs = tf.placeholder(tf.float32, [256, 5])
    w = Normal(loc=tf.zeros(5), scale=tf.ones(5))
    yrv = Bernoulli(logits=ed.dot(xs, w))
    qw = Empirical(params=tf.Variable(tf.zeros([T, 5])))
    inf = ed.SGHMC({w: qw}, data={xs: X, yrv: y})
    inf.initialize(step_size=0.05, friction=0.2, n_print=0)
    tf.global_variables_initializer().run()
    return inf, qw

  T = 20
  full, qfull = build(T)
  for _ in range(T):
    full.update()
  want = qfull.params.eval().copy()
  a, qa = build(T)
  for _ in range(9):
    a.update()
  state = a.state_dict()
  assert "velocity" in state and np.any(state["velocity"] != 0)
  b, qb = build(T)
  b.load_state_dict(state)
  for _ in range(T - 9):
    b.update()
  np.testing.assert_array_equal(qb.params.eval(), want)
