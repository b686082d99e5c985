"""float64 models (reference: tests/inferences/hmc_test.py:93-103 runs test_normal_normal and test_linear_regression in
tf.float32 AND tf.float64). The float64 path is the compact device path of csrc/f64.cuh behind the typed entry points
edhmc_*_f64; the oracle is the same restatement evaluated in np.float64, so the tolerance is that of two different
summation orders in double (1e-11), far inside the 1e-5 / 1e-4 of north_star.
"""
import numpy as np
import pytest

import hmc_oracle as o

pytestmark = pytest.mark.gpu


def _mk(N, D, has_bias, family, seed, lik_scale=0.5):
  rng = np.random.default_rng(seed)
  X = rng.standard_normal((N, D))
  w_true = rng.standard_normal(D) / np.sqrt(D)
  eta = X @ w_true
  if family == o.BERNOULLI_LOGIT:
    y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.int32)
  elif family == o.NORMAL_IDENTITY:
    y = eta + lik_scale * rng.standard_normal(N)
  else:
    y = rng.poisson(np.exp(np.clip(eta, -3, 3))).astype(np.int32)
  P = D + int(has_bias)
  # prior parameters and the likelihood scale cross the C ABI as float32 (edhmc_cfg); exactly representable values here
  spec = o.GLMSpec(D, has_bias, family, np.zeros(P, np.float32), np.full(P, 1.5, np.float32), lik_scale)
  return X, y, spec


def _sampler(X, y, spec, **kw):
  import torch
  from edward_b200 import engine
  es = engine.GLMSpec(spec.n_features, spec.has_bias, spec.family, spec.prior_loc, spec.prior_scale, spec.lik_scale)
  return engine.GLMSampler(es, X, y, dtype=torch.float64, **kw)


def _rel(a, b):
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


SHAPES = [
    (50, 1, False, o.NORMAL_IDENTITY),    # hmc_test.py:14-46 (ones design)
    (40, 10, True, o.NORMAL_IDENTITY),    # hmc_test.py:48-91
    (1, 1, False, o.BERNOULLI_LOGIT),
    (777, 7, True, o.BERNOULLI_LOGIT),
    (5000, 54, False, o.BERNOULLI_LOGIT),
    (300, 131, True, o.BERNOULLI_LOGIT),
    (2000, 12, True, o.POISSON_LOG),
]


@pytest.mark.parametrize("N,D,bias,fam", SHAPES)
def test_f64_logp_grad_matches_oracle(N, D, bias, fam):
  import torch
  X, y, spec = _mk(N, D, bias, fam, seed=N + D)
  s = _sampler(X, y, spec)
  rng = np.random.default_rng(1)
  for scale in (0.0, 0.3):
    theta = scale * rng.standard_normal(spec.n_params) / np.sqrt(D)
    lp, g = s.logp_grad(theta)
    assert g.dtype == torch.float64
    lp = float(lp.cpu()[0])
    lp64 = float(o.log_joint(X, y, theta, spec, np.float64))
    g64 = o.grad_log_joint(X, y, theta, spec, np.float64)
    assert abs(lp - lp64) <= 1e-11 * abs(lp64), (lp, lp64)
    assert _rel(g.cpu().numpy(), g64) <= 1e-11
  s.close()


RUNS = [
    (50, 1, False, o.NORMAL_IDENTITY, 40, 2, 0.25),         # the reference's defaults (hmc.py:45)
    (40, 10, True, o.NORMAL_IDENTITY, 40, 2, 0.01),         # hmc_test.py:78 step_size
    (1024, 54, False, o.BERNOULLI_LOGIT, 20, 10, 0.5 / 1024),
    (2000, 12, True, o.POISSON_LOG, 20, 3, 0.004),
]


@pytest.mark.parametrize("N,D,bias,fam,T,L,eps", RUNS)
def test_f64_run_matches_oracle(N, D, bias, fam, T, L, eps):
  import torch
  X, y, spec = _mk(N, D, bias, fam, seed=7 * N + D)
  P = spec.n_params
  r0, u = o.synth_draws(T, P, seed=3)
  r0 = r0.astype(np.float64)
  u = u.astype(np.float64)
  p64 = np.zeros((T, P), np.float64)
  infos, nacc = o.run(X, y, p64, r0, u, eps, L, spec, np.float64)
  s = _sampler(X, y, spec)
  params = torch.zeros(T, P, dtype=torch.float64, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  n_accept, logp_cur = s.read_state()
  sc = sc.cpu().numpy()
  pos = pos.cpu().numpy()
  for i, info in enumerate(infos):
    assert abs(sc[i, 0] - info.logp_old) <= 1e-10 * abs(info.logp_old) + 1e-10
    assert abs(sc[i, 1] - info.logp_new) <= 1e-10 * abs(info.logp_new) + 1e-10
    assert _rel(pos[i], info.proposal) <= 1e-10
    assert abs(sc[i, 4] - info.ratio) <= 1e-8 * max(abs(info.logp_new), 1.0)
    assert bool(sc[i, 6] > 0.5) == info.accept, (i, info)
  assert n_accept == nacc
  assert _rel(params.cpu().numpy(), p64) <= 1e-10
  # chunked launches continue the same chain bit for bit (state cached on the device across calls)
  params2 = torch.zeros(T, P, dtype=torch.float64, device="cuda")
  s2 = _sampler(X, y, spec)
  r0t, ut = torch.tensor(r0, device="cuda"), torch.tensor(u, device="cuda")
  for t in range(T):
    s2.run(params2, t, 1, eps, L, r0=r0t[t:t + 1], u=ut[t:t + 1])
  assert torch.equal(params, params2)
  s.close()
  s2.close()


def test_f64_handle_refuses_f32_entry_points_and_back():
  import ctypes as C
  import torch
  from edward_b200 import _C
  X, y, spec = _mk(64, 4, False, o.BERNOULLI_LOGIT, seed=2)
  s64 = _sampler(X, y, spec)
  buf = torch.zeros(8, 4, dtype=torch.float64, device="cuda")
  lib = _C.lib()
  assert lib.edhmc_run(s64._h, buf.data_ptr(), 4, 8, 0, 1, C.c_float(0.1), 2, None, None, None) == _C.ERR_INVALID
  assert b"edhmc_run_f64" in lib.edhmc_last_error()
  with pytest.raises(_C.EdhmcError):
    s64.sgmcmc_run("sgld", buf.float(), 0, 1, 0.1)
  s64.close()


def test_f64_nonfinite_data_raises_at_bind():
  from edward_b200 import _C
  X, y, spec = _mk(100, 3, False, o.NORMAL_IDENTITY, seed=5)
  X[17, 1] = np.nan
  with pytest.raises(_C.NonFiniteError):
    _sampler(X, y, spec)


# ---- the reference's own float64 cases through the front-end (hmc_test.py:14-103) ----------------------------------
@pytest.mark.parametrize("default", [True, False])
def test_reference_normal_normal_float64(default):
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Empirical, Normal
  tf.reset_default_graph()
  ed.set_seed(42)
  dtype = tf.float64
  x_data = np.array([0.0] * 50, dtype=np.float32)
  mu = Normal(loc=tf.constant(0.0, dtype=dtype), scale=tf.constant(1.0, dtype=dtype))
  x = Normal(loc=mu, scale=tf.constant(1.0, dtype=dtype), sample_shape=50)
  n_samples = 2000
  if not default:
    qmu = Empirical(params=tf.Variable(tf.ones(n_samples, dtype=dtype)))
    inference = ed.HMC({mu: qmu}, data={x: x_data})
  else:
    inference = ed.HMC([mu], data={x: x_data})
    qmu = inference.latent_vars[mu]
  inference.run(n_print=0 if default else None)
  assert qmu.params.dtype == tf.float64
  assert inference._sampler.dtype.is_floating_point and str(inference._sampler.dtype) == "torch.float64"
  np.testing.assert_allclose(qmu.mean().eval(), 0, rtol=1e-1, atol=1e-1)
  np.testing.assert_allclose(qmu.stddev().eval(), np.sqrt(1 / 51), rtol=1e-1, atol=1e-1)
  sess = ed.get_session()
  old_t, old_n_accept = sess.run([inference.t, inference.n_accept])
  assert old_t == (n_samples if not default else 1e4)
  assert old_n_accept > 0.1
  sess.run(inference.reset)
  new_t, new_n_accept = sess.run([inference.t, inference.n_accept])
  assert new_t == 0 and new_n_accept == 0


@pytest.mark.parametrize("default", [True, False])
def test_reference_linear_regression_float64(default):
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Empirical, Normal
  tf.reset_default_graph()
  ed.set_seed(42)
  dtype = tf.float64
  rng = np.random.RandomState(42)

  def build_toy_dataset(N, w, noise_std=0.1):
    x = rng.randn(N, len(w))
    return x, np.dot(x, w) + rng.normal(0, noise_std, size=N)

  N, D = 40, 10
  w_true = rng.randn(D)
  X_train, y_train = build_toy_dataset(N, w_true)
  X = tf.placeholder(dtype, [N, D])
  w = Normal(loc=tf.zeros(D, dtype=dtype), scale=tf.ones(D, dtype=dtype))
  b = Normal(loc=tf.zeros(1, dtype=dtype), scale=tf.ones(1, dtype=dtype))
  y = Normal(loc=ed.dot(X, w) + b, scale=0.1 * tf.ones(N, dtype=dtype))
  n_samples = 2000
  if not default:
    qw = Empirical(tf.Variable(tf.zeros([n_samples, D], dtype=dtype)))
    qb = Empirical(tf.Variable(tf.zeros([n_samples, 1], dtype=dtype)))
    inference = ed.HMC({w: qw, b: qb}, data={X: X_train, y: y_train})
  else:
    inference = ed.HMC([w, b], data={X: X_train, y: y_train})
    qw = inference.latent_vars[w]
    qb = inference.latent_vars[b]
  inference.run(step_size=0.01, n_print=0 if default else None)
  assert qw.params.dtype == tf.float64
  np.testing.assert_allclose(qw.mean().eval(), w_true, rtol=5e-1, atol=5e-1)
  np.testing.assert_allclose(qb.mean().eval(), [0.0], rtol=5e-1, atol=5e-1)
  sess = ed.get_session()
  old_t, old_n_accept = sess.run([inference.t, inference.n_accept])
  assert old_t == (n_samples if not default else 1e4)
  assert old_n_accept > 0.1
