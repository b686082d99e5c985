"""SURVEY §8f ranks 2 and 4 on the GPU: posterior-predictive criticism over the device-resident sample store
(ed.copy / ed.evaluate / ed.ppc, criticisms/evaluate.py:20-235, ppc.py:13) and checkpoint / resume."""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fresh_graph():
  from edward_b200 import graph as g
  g.reset_default_graph()
  yield


def _fit(N=4000, D=6, T=400, seed=3):
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Empirical, Normal
  ed.set_seed(seed)
  Xv, yv, w_true = o.synth_data(N, D)
  X = tf.placeholder(tf.float32, [N, D])
  w = Normal(loc=tf.zeros(D), scale=tf.ones(D))
  b = Normal(loc=tf.zeros([]), scale=tf.ones([]))
  y = Bernoulli(logits=ed.dot(X, w) + b)
  qw = Empirical(params=tf.Variable(tf.zeros([T, D])))
  qb = Empirical(params=tf.Variable(tf.zeros([T])))
  inference = ed.HMC({w: qw, b: qb}, data={X: Xv, y: yv})
  inference.run(step_size=0.02, n_steps=5, n_print=0)
  return ed, (X, w, b, y, qw, qb, inference), (Xv, yv, w_true)


def test_evaluate_binary_accuracy_and_log_lik_match_host_computation():
  ed, (X, w, b, y, qw, qb, inference), (Xv, yv, w_true) = _fit()
  y_post = ed.copy(y, {w: qw, b: qb})
  acc, ll = ed.evaluate(['binary_accuracy', 'log_lik'], data={X: Xv, y_post: yv}, n_samples=300, seed=1)
  W, B = qw.params.eval()[100:], qb.params.eval()[100:]
  eta = Xv @ W.T + B[None, :]
  probs = (1 / (1 + np.exp(-eta))).mean(axis=1)
  acc_host = np.mean((probs > 0.5) == (yv > 0.5))
  ll_host = np.mean(-(np.maximum(eta, 0) - eta * yv[:, None] + np.log1p(np.exp(-np.abs(eta)))))
  assert abs(acc - acc_host) < 0.02, (acc, acc_host)
  assert abs(ll - ll_host) < 0.05, (ll, ll_host)
  bayes = np.mean(((Xv @ w_true) > 0) == (yv > 0.5))
  assert acc > bayes - 0.03
  with pytest.raises(KeyError):
    ed.evaluate('binary_accuracy', data={X: Xv, y_post: yv, y: yv})
  with pytest.raises(NotImplementedError):
    ed.evaluate('hinge_loss_that_does_not_exist', data={X: Xv, y_post: yv})


def test_evaluate_mse_for_linear_regression():
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Empirical, Normal
  rng = np.random.RandomState(0)
  N, D, T = 200, 5, 600
  w_true = rng.randn(D)
  Xv = rng.randn(N, D).astype(np.float32)
  yv = (Xv @ w_true + rng.normal(0, 0.1, N)).astype(np.float32)
  X = tf.placeholder(tf.float32, [N, D])
  w = Normal(loc=tf.zeros(D), scale=tf.ones(D))
  y = Normal(loc=ed.dot(X, w), scale=0.1 * tf.ones(N))
  qw = Empirical(params=tf.Variable(tf.zeros([T, D])))
  ed.HMC({w: qw}, data={X: Xv, y: yv}).run(step_size=0.005, n_steps=5, n_print=0)
  y_post = ed.copy(y, {w: qw})
  mse, mae = ed.evaluate(['mse', 'mae'], data={X: Xv, y_post: yv}, n_samples=400)
  assert mse < 0.1 and mae < 0.25, (mse, mae)


def test_ppc_runs_on_posterior_predictive():
  ed, (X, w, b, y, qw, qb, inference), (Xv, yv, w_true) = _fit(N=500, T=100)
  y_post = ed.copy(y, {w: qw, b: qb})
  Trep, Tobs = ed.ppc(lambda xs, zs: np.mean(xs[y_post]), data={X: Xv, y_post: yv}, latent_vars={w: qw, b: qb}, n_samples=20)
  assert Trep.shape == (20,) and Tobs.shape == (20,)
  assert np.allclose(Tobs, yv.mean())
  assert abs(Trep[5:].mean() - yv.mean()) < 0.2


def test_checkpoint_resume_is_bit_identical():
  import edward_b200 as ed
  from edward_b200 import graph as g
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Empirical, Normal
  N, D, T = 3000, 10, 40
  Xv, yv, _ = o.synth_data(N, D)

  def build():
    g.reset_default_graph()
    ed.set_seed(11)
    X = tf.placeholder(tf.float32, [N, D])
    w = Normal(loc=tf.zeros(D), scale=tf.ones(D))
    y = Bernoulli(logits=ed.dot(X, w))
    qw = Empirical(params=tf.Variable(tf.zeros([T, D])))
    inf = ed.HMC({w: qw}, data={X: Xv, y: yv})
    inf.initialize(step_size=0.02, n_steps=4, n_print=0)
    tf.global_variables_initializer().run()
    return inf, qw

  inf, qw = build()
  for _ in range(T):
    inf.update()
  full = qw.params.eval().copy()
  n_full = int(inf.n_accept.eval())

  inf, qw = build()
  for _ in range(T // 2):
    inf.update()
  ckpt = inf.state_dict()
  inf2, qw2 = build()
  inf2.load_state_dict(ckpt)
  assert int(inf2.t.eval()) == T // 2
  for _ in range(T - T // 2):
    inf2.update()
  assert np.array_equal(qw2.params.eval(), full)
  assert int(inf2.n_accept.eval()) == n_full


PRED = [
    # N, D, S, bias, family
    (1, 1, 1, False, o.BERNOULLI_LOGIT),
    (63, 3, 5, True, o.BERNOULLI_LOGIT),
    (1000, 54, 300, True, o.BERNOULLI_LOGIT),   # cfg 2 feature count, evaluate's default order of draws
    (777, 200, 65, False, o.BERNOULLI_LOGIT),   # several column chunks, a ragged draw chunk
    (500, 17, 64, True, o.NORMAL_IDENTITY),
    (300, 12, 130, True, o.POISSON_LOG),
]


@pytest.mark.parametrize("N,D,S,bias,fam", PRED)
def test_predictive_kernel_matches_oracle(N, D, S, bias, fam):
  """edhmc_predictive (the fused contraction + link + reduction over draws) against the float64 restatement of
  evaluate.py:132-143,158-162,222-227 on the same draws: 1e-5 relative."""
  import torch
  from edward_b200.criticisms.evaluate import predictive
  rng = np.random.default_rng(N + D + S)
  X = rng.standard_normal((N, D)).astype(np.float32)
  W = (rng.standard_normal((S, D)) / np.sqrt(D)).astype(np.float32)
  B = (0.3 * rng.standard_normal(S)).astype(np.float32) if bias else None
  if fam == o.BERNOULLI_LOGIT:
    y = (rng.random(N) < 0.5).astype(np.int32)
  elif fam == o.NORMAL_IDENTITY:
    y = rng.standard_normal(N).astype(np.float32)
  else:
    y = rng.poisson(1.0, N).astype(np.int32)
  dev = torch.device("cuda")
  mean, ll = predictive(torch.tensor(X, device=dev), torch.tensor(y, device=dev), torch.tensor(W, device=dev),
                        torch.tensor(B, device=dev) if bias else None, fam, 0.7)
  mean_o, ll_o = o.predictive(X, y, W, B, fam, 0.7)
  assert np.max(np.abs(mean.cpu().numpy() - mean_o)) <= 1e-5 * max(1.0, np.max(np.abs(mean_o)))
  assert np.max(np.abs(ll.cpu().numpy() - ll_o)) <= 1e-5 * np.max(np.abs(ll_o)) + 1e-6
  # a strided design matrix (a column slice of a wider array) gives the same numbers
  Xw = torch.zeros(N, D + 3, device=dev)
  Xw[:, :D] = torch.tensor(X, device=dev)
  mean2, ll2 = predictive(Xw[:, :D], torch.tensor(y, device=dev), torch.tensor(W, device=dev),
                          torch.tensor(B, device=dev) if bias else None, fam, 0.7)
  assert torch.equal(mean, mean2) and torch.equal(ll, ll2)
