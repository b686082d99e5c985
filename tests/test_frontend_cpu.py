"""Host-side logic of the drop-in surface — CPU only: argument checking with the reference's error
conventions, the GLM recogniser, tfshim, Progbar, ed.dot, Empirical read-outs."""
import io
import sys

import numpy as np
import pytest


@pytest.fixture(autouse=True)
def fresh_graph():
  from edward_b200 import graph as g
  g.reset_default_graph()
  yield


def _imports():
  import edward_b200 as ed
  from edward_b200 import tfshim as tf
  from edward_b200.models import Bernoulli, Empirical, Normal, Poisson
  return ed, tf, Bernoulli, Empirical, Normal, Poisson


def test_dot_equals_numpy_and_raises_on_inf():
  """tests/util/dot_test.py:13-31."""
  ed, tf, *_ = _imports()
  a = tf.constant(np.arange(5, dtype=np.float32))
  b = tf.constant(np.diag(np.ones(5, dtype=np.float32)) * 2)
  np.testing.assert_allclose(ed.dot(a, b).eval(), np.dot(a.eval(), b.eval()))
  np.testing.assert_allclose(ed.dot(b, a).eval(), np.dot(b.eval(), a.eval()))
  binf = b.eval().copy()
  binf[0, 0] = np.inf
  with pytest.raises(ValueError, match="Inf"):
    ed.dot(a, tf.constant(binf)).eval()
  with pytest.raises(ValueError, match="Inf"):
    ed.dot(tf.constant(binf), a).eval()


def test_check_data_and_latent_vars_error_conventions():
  """util/random_variables.py:21-83."""
  ed, tf, Bernoulli, Empirical, Normal, _ = _imports()
  x = Normal(loc=tf.zeros(3), scale=tf.ones(3))
  ph = tf.placeholder(tf.float32, [3])
  with pytest.raises(TypeError):
    ed.check_data([1, 2])
  with pytest.raises(TypeError):
    ed.check_data({ph: x})                       # feed value cannot be a RandomVariable
  with pytest.raises(TypeError):
    ed.check_data({ph: tf.constant([1., 2., 3.])})  # nor a tf.Tensor
  with pytest.raises(TypeError):
    ed.check_data({x: np.zeros(4, np.float32)})  # shape mismatch
  with pytest.raises(TypeError):
    ed.check_data({x: Normal(loc=tf.zeros(3, dtype=tf.float64), scale=tf.ones(3, dtype=tf.float64))})  # dtype
  with pytest.raises(TypeError):
    ed.check_data({"x": 1.0})
  ed.check_data({x: np.zeros(3, np.float32), ph: [1., 2., 3.]})
  with pytest.raises(TypeError):
    ed.check_latent_vars([x])
  with pytest.raises(TypeError):
    ed.check_latent_vars({x: 3.0})
  with pytest.raises(TypeError):
    ed.check_latent_vars({x: Normal(loc=tf.zeros(2), scale=tf.ones(2))})
  ed.check_latent_vars({x: Empirical(params=tf.Variable(tf.zeros([10, 3])))})


def test_montecarlo_constructor_conventions():
  """monte_carlo.py:61-93."""
  ed, tf, Bernoulli, Empirical, Normal, _ = _imports()
  w = Normal(loc=tf.zeros(3), scale=tf.ones(3))
  with pytest.raises(TypeError, match="Empirical"):
    ed.HMC({w: Normal(loc=tf.zeros(3), scale=tf.ones(3))})
  inf = ed.HMC([w])
  qw = inf.latent_vars[w]
  assert isinstance(qw, Empirical) and tuple(qw.params.shape) == (10000, 3)
  inf2 = ed.HMC({w: Empirical(params=tf.Variable(tf.zeros([7, 3])))})
  assert list(inf2.latent_vars.values())[0].n == 7


def test_recogniser_accepts_hot_path_models():
  ed, tf, Bernoulli, Empirical, Normal, Poisson = _imports()
  from edward_b200 import _C
  from edward_b200.glm import recognize
  X = tf.placeholder(tf.float32, [40, 3])
  w = Normal(loc=tf.zeros(3), scale=3.0 * tf.ones(3))
  b = Normal(loc=tf.zeros([]), scale=2.0 * tf.ones([]))
  qw = Empirical(params=tf.Variable(tf.zeros([5, 3])))
  qb = Empirical(params=tf.Variable(tf.zeros([5])))
  y = Bernoulli(logits=ed.dot(X, w) + b)
  m = recognize({w: qw, b: qb}, {X: np.zeros((40, 3)), y: np.zeros(40)})
  assert (m.spec.n_features, m.spec.has_bias, m.spec.family) == (3, True, _C.BERNOULLI_LOGIT)
  np.testing.assert_array_equal(m.spec.prior_scale, [3, 3, 3, 2])
  assert [(s.offset, s.size, s.scalar) for s in m.slots] == [(0, 3, False), (3, 1, True)]
  y2 = Bernoulli(logits=b + ed.dot(X, w))       # either order of the sum
  assert recognize({b: qb, w: qw}, {X: np.zeros((40, 3)), y2: np.zeros(40)}).spec.has_bias
  y3 = Normal(loc=ed.dot(X, w), scale=0.5 * tf.ones(40))
  m3 = recognize({w: qw}, {X: np.zeros((40, 3)), y3: np.zeros(40, np.float32)})
  assert m3.spec.family == _C.NORMAL_IDENTITY and abs(m3.spec.lik_scale - 0.5) < 1e-7
  y4 = Poisson(log_rate=ed.dot(X, w))
  assert recognize({w: qw}, {X: np.zeros((40, 3)), y4: np.zeros(40)}).spec.family == _C.POISSON_LOG
  mu = Normal(loc=tf.constant(0.0), scale=tf.constant(1.0))
  x = Normal(loc=mu, scale=tf.constant(1.0), sample_shape=50)
  m5 = recognize({mu: Empirical(params=tf.Variable(tf.zeros(9)))}, {x: np.zeros(50, np.float32)})
  assert (m5.spec.n_features, m5.n_rows, m5.x_node) == (1, 50, None)


def test_recogniser_rejects_everything_else():
  ed, tf, Bernoulli, Empirical, Normal, Poisson = _imports()
  from edward_b200.glm import recognize
  X = tf.placeholder(tf.float32, [10, 2])
  w = Normal(loc=tf.zeros(2), scale=tf.ones(2))
  qw = Empirical(params=tf.Variable(tf.zeros([5, 2])))
  data = {X: np.zeros((10, 2))}

  def with_obs(rv):
    d = dict(data)
    d[rv] = np.zeros(10)
    return d
  with pytest.raises(NotImplementedError):
    recognize({w: qw}, with_obs(Bernoulli(probs=tf.sigmoid(ed.dot(X, w)))))
  with pytest.raises(NotImplementedError):
    recognize({w: qw}, with_obs(Bernoulli(logits=ed.dot(X, w) * 2.0)))
  with pytest.raises(NotImplementedError):
    recognize({w: qw}, data)  # no observed variable
  w64 = Normal(loc=tf.zeros(2, dtype=tf.float64), scale=tf.ones(2, dtype=tf.float64))
  X64 = tf.placeholder(tf.float64, [10, 2])
  # float64 models are recognised (hmc_test.py:93-97) ...
  m64 = recognize({w64: Empirical(params=tf.Variable(tf.zeros([5, 2], dtype=tf.float64)))},
                  {X64: np.zeros((10, 2)), Bernoulli(logits=ed.dot(X64, w64)): np.zeros(10)})
  assert m64.dtype == "float64" and m64.spec.n_features == 2
  # ... but latents of mixed dtypes are not
  b32 = Normal(loc=tf.zeros(1), scale=tf.ones(1))
  with pytest.raises(NotImplementedError, match="share one dtype"):
    recognize({w64: Empirical(params=tf.Variable(tf.zeros([5, 2], dtype=tf.float64))),
               b32: Empirical(params=tf.Variable(tf.zeros([5, 1])))},
              {X64: np.zeros((10, 2)), Bernoulli(logits=ed.dot(X64, w64) + b32): np.zeros(10)})
  extra = Normal(loc=tf.zeros(2), scale=tf.ones(2))
  with pytest.raises(NotImplementedError):
    recognize({w: qw, extra: Empirical(params=tf.Variable(tf.zeros([5, 2])))},
              with_obs(Bernoulli(logits=ed.dot(X, w))))


def test_initialize_argument_conventions_cpu():
  ed, tf, Bernoulli, Empirical, Normal, _ = _imports()
  X = tf.placeholder(tf.float32, [10, 2])
  w = Normal(loc=tf.zeros(2), scale=tf.ones(2))
  y = Bernoulli(logits=ed.dot(X, w))
  qw = Empirical(params=tf.Variable(tf.zeros([5, 2])))
  inf = ed.HMC({w: qw}, data={X: np.zeros((10, 2), np.float32), y: np.zeros(10)})
  assert inf.data[y].dtype == np.int32          # cast to the random variable's dtype (inference.py:88-95)
  with pytest.raises(TypeError, match="scale must be a dict"):
    inf.initialize(scale=3.0)
  with pytest.raises(ValueError, match="auto_transform=True"):
    inf.initialize(auto_transform=False)


def test_set_seed_guard_and_empirical_readouts():
  ed, tf, Bernoulli, Empirical, Normal, _ = _imports()
  ed.set_seed(3)
  p = np.random.RandomState(0).randn(100, 4).astype(np.float32)
  q = Empirical(params=tf.Variable(p))
  np.testing.assert_allclose(q.mean().eval(), p.mean(0), rtol=1e-6)
  np.testing.assert_allclose(q.stddev().eval(), p.std(0), rtol=1e-5)       # population std (empirical.py:90-93)
  np.testing.assert_allclose(q.variance().eval(), p.var(0), rtol=1e-5)
  assert q.sample().eval().shape == (4,) and q.sample(7).eval().shape == (7, 4)   # empirical_sample_test.py:13-31
  assert tuple(q.event_shape) == (4,) and q.n == 100
  with pytest.raises(RuntimeError):
    ed.set_seed(4)  # after part of the graph exists (util/graphs.py:66-70)


def test_tfshim_symbols_used_by_the_example():
  ed, tf, *_ = _imports()
  assert tf.zeros([2, 3]).eval().shape == (2, 3) and tf.ones([]).eval().shape == ()
  assert (3.0 * tf.ones(2)).eval().tolist() == [3.0, 3.0]
  v = tf.get_variable("qw/params", [6, 2])
  assert v.eval().shape == (6, 2) and np.all(np.abs(v.eval()) <= np.sqrt(6.0 / 8) + 1e-6)
  with pytest.raises(ValueError):
    tf.get_variable("qw/params", [6, 2])
  v.load(np.ones((6, 2)))
  tf.global_variables_initializer().run()
  np.testing.assert_array_equal(v.eval(), v.initial_value)
  ph = tf.placeholder(tf.float32, [2])
  np.testing.assert_allclose(tf.sigmoid(ph).eval({ph: [0.0, 100.0]}), [0.5, 1.0])
  with pytest.raises(ValueError):
    ph.eval()
  tf.flags.DEFINE_integer("some_flag", default=7, help="")
  assert tf.flags.FLAGS.some_flag == 7


def test_progbar_output_format():
  """util/progbar.py:38-115."""
  ed, *_ = _imports()
  buf, old = io.StringIO(), sys.stdout
  sys.stdout = buf
  try:
    bar = ed.Progbar(50, interval=0)
    bar.update(1, {'Acceptance Rate': 0.5})
    bar.update(50, {'Acceptance Rate': 0.25})
  finally:
    sys.stdout = old
  out = buf.getvalue()
  assert " 1/50 [  2%]" in out and "ETA:" in out
  assert "50/50 [100%]" in out and "Elapsed:" in out and "Acceptance Rate: 0.250" in out
  assert out.endswith("\n")


def test_host_mirror_of_adopted_variables_follows_the_device_epoch():
  """graph.Variable.host_view: one device-to-host copy per device-write epoch. A CPU torch tensor stands in for the
  device store here; engine.GLMSampler.run / sgmcmc_run / run_chains, Variable.load / rebind and checkpoint restore bump
  the epoch on the real path."""
  import torch
  ed, tf, Bernoulli, Empirical, Normal, Poisson = _imports()
  from edward_b200 import graph as g
  v = tf.Variable(tf.zeros([6, 2]))
  assert v.host_view() is v._host                       # not adopted: the host array itself
  store = torch.arange(12, dtype=torch.float32).reshape(6, 2)
  v.rebind(store)                                       # adoption copies the current contents (zeros) in
  assert torch.equal(store, torch.zeros(6, 2))
  a = v.host_view()
  assert v.host_view() is a                             # cached
  store += 1.0                                          # a "kernel" writes the store ...
  assert np.all(v.host_view() == 0.0)                   # ... the mirror is stale until the launch site bumps the epoch
  g.bump_device_epoch()
  assert np.all(v.host_view() == 1.0) and v.host_view() is not a
  v.load(np.full((6, 2), 3.0, np.float32))              # load() writes through and bumps
  assert np.all(v.numpy() == 3.0) and torch.equal(store, torch.full((6, 2), 3.0))
  out = v.numpy()
  out[:] = -1.0                                         # numpy() hands out a copy
  assert np.all(v.host_view() == 3.0)
  q = Empirical(params=v)
  s = q.sample(5).eval()                                # draws rows from the mirror
  assert s.shape == (5, 2) and np.all(s == 3.0)
