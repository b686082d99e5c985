"""Normal, Bernoulli, Poisson — the random variables of the HMC hot path. In the reference these are
generated wrappers over tf.contrib.distributions (edward/models/random_variables.py:13-25); only their
constructor signatures, `support` tags (:27-58) and eager read-outs are mirrored here. The densities the
sampler needs are evaluated by libedhmc, not by these classes."""
from __future__ import annotations

import math

import numpy as np

from .. import graph as _g
from .random_variable import RandomVariable


def _bshape(*ts):
  shape = ()
  for t in ts:
    shape = np.broadcast_shapes(shape, tuple(d if d is not None else 1 for d in t.shape))
  return shape


class Normal(RandomVariable):
  support = 'real'

  def __init__(self, loc, scale, validate_args=False, allow_nan_stats=True, name="Normal", **kwargs):
    self.loc = _g.convert_to_tensor(loc)
    self.scale = _g.convert_to_tensor(scale, self.loc.dtype if not isinstance(scale, _g.Tensor) else None)
    if self.loc.dtype != self.scale.dtype:
      raise TypeError("loc and scale must have the same dtype: %r vs %r" % (self.loc.dtype, self.scale.dtype))
    self._args, self._kwargs = (loc, scale), dict(kwargs)
    super(Normal, self).__init__(_bshape(self.loc, self.scale), (), self.loc.dtype, name=name, **kwargs)

  def _sample_np(self, sample_shape, feed=None):
    loc, scale = _g.eval_in(self.loc, feed), _g.eval_in(self.scale, feed)
    shape = tuple(sample_shape) + tuple(self.batch_shape)
    return (loc + scale * np.random.standard_normal(shape)).astype(self.dtype.np)

  def log_prob(self, value):
    v = _g.convert_to_tensor(value, self.dtype)

    def fn():
      x, loc, scale = _g.evaluate(v), _g.evaluate(self.loc), _g.evaluate(self.scale)
      return -0.5 * np.square((x - loc) / scale) - (0.5 * math.log(2.0 * math.pi) + np.log(scale))
    return _g.Lazy(fn, _bshape(v, self.loc, self.scale), self.dtype, "NormalLogProb")

  def mean(self):
    return self.loc

  def stddev(self):
    return self.scale


class Bernoulli(RandomVariable):
  support = 'binary'

  def __init__(self, logits=None, probs=None, dtype=_g.int32, validate_args=False, allow_nan_stats=True,
               name="Bernoulli", **kwargs):
    if (logits is None) == (probs is None):
      raise ValueError("Must pass probs or logits, but not both.")
    self.logits = _g.convert_to_tensor(logits) if logits is not None else None
    self._probs = _g.convert_to_tensor(probs) if probs is not None else None
    self._args, self._kwargs = (), dict(kwargs, logits=logits, probs=probs)
    param = self.logits if self.logits is not None else self._probs
    super(Bernoulli, self).__init__(tuple(param.shape), (), dtype, name=name, **kwargs)

  def _probs_np(self, feed=None):
    if self._probs is not None:
      return _g.eval_in(self._probs, feed)
    return 1.0 / (1.0 + np.exp(-_g.eval_in(self.logits, feed)))

  @property
  def probs(self):
    if self._probs is not None:
      return self._probs
    return _g.Lazy(self._probs_np, tuple(self.logits.shape), self.logits.dtype, "Sigmoid", wants_feed=True)

  def _sample_np(self, sample_shape, feed=None):
    p = self._probs_np(feed)
    return (np.random.uniform(size=tuple(sample_shape) + p.shape) < p).astype(self.dtype.np)

  def log_prob(self, value):
    v = _g.convert_to_tensor(value)
    fdt = self.logits.dtype if self.logits is not None else self._probs.dtype

    def fn():
      y = _g.evaluate(v).astype(fdt.np)
      if self.logits is not None:
        l = _g.evaluate(self.logits)
        return -(np.maximum(l, 0) - l * y + np.log1p(np.exp(-np.abs(l))))
      p = _g.evaluate(self._probs)
      return y * np.log(p) + (1.0 - y) * np.log1p(-p)
    return _g.Lazy(fn, tuple(self.shape), fdt, "BernoulliLogProb")

  def mean(self):
    return self.probs


class Poisson(RandomVariable):
  support = 'countable'

  def __init__(self, rate=None, log_rate=None, validate_args=False, allow_nan_stats=True, name="Poisson", **kwargs):
    if (rate is None) == (log_rate is None):
      raise ValueError("Must specify exactly one of `rate` and `log_rate`.")
    self.log_rate = _g.convert_to_tensor(log_rate) if log_rate is not None else None
    self._rate = _g.convert_to_tensor(rate) if rate is not None else None
    self._args, self._kwargs = (), dict(kwargs, rate=rate, log_rate=log_rate)
    param = self.log_rate if self.log_rate is not None else self._rate
    super(Poisson, self).__init__(tuple(param.shape), (), param.dtype, name=name, **kwargs)

  def _rate_np(self, feed=None):
    return _g.eval_in(self._rate, feed) if self._rate is not None else np.exp(_g.eval_in(self.log_rate, feed))

  def _sample_np(self, sample_shape, feed=None):
    lam = self._rate_np(feed)
    return np.random.poisson(lam, size=tuple(sample_shape) + lam.shape).astype(self.dtype.np)

  def log_prob(self, value):
    from scipy.special import gammaln
    v = _g.convert_to_tensor(value)

    def fn():
      y = _g.evaluate(v).astype(self.dtype.np)
      lam = self._rate_np()
      return y * np.log(lam) - lam - gammaln(y + 1.0)
    return _g.Lazy(fn, tuple(self.shape), self.dtype, "PoissonLogProb")


class Beta(RandomVariable):
  """Beta(concentration1, concentration0) — a latent with support (0, 1) (edward/models/random_variables.py:27-58 tags it
  '01'); under auto_transform HMC samples it through a sigmoid (inference.py:223-264)."""
  support = '01'

  def __init__(self, concentration1=None, concentration0=None, validate_args=False, allow_nan_stats=True, name="Beta",
               **kwargs):
    self.concentration1 = _g.convert_to_tensor(concentration1)
    self.concentration0 = _g.convert_to_tensor(concentration0, self.concentration1.dtype
                                               if not isinstance(concentration0, _g.Tensor) else None)
    self._args, self._kwargs = (concentration1, concentration0), dict(kwargs)
    super(Beta, self).__init__(_bshape(self.concentration1, self.concentration0), (), self.concentration1.dtype, name=name,
                               **kwargs)

  def _ab(self, feed=None):
    return _g.eval_in(self.concentration1, feed), _g.eval_in(self.concentration0, feed)

  def _sample_np(self, sample_shape, feed=None):
    a, b = self._ab(feed)
    return np.random.beta(a, b, size=tuple(sample_shape) + tuple(self.batch_shape)).astype(self.dtype.np)

  def log_prob(self, value):
    from scipy.special import betaln
    v = _g.convert_to_tensor(value, self.dtype)

    def fn():
      x = _g.evaluate(v)
      a, b = self._ab()
      return (a - 1.0) * np.log(x) + (b - 1.0) * np.log1p(-x) - betaln(a, b)
    return _g.Lazy(fn, _bshape(v, self.concentration1), self.dtype, "BetaLogProb")

  def mean(self):
    return _g.Lazy(lambda: (lambda a, b: a / (a + b))(*self._ab()), tuple(self.batch_shape), self.dtype, "BetaMean")

  def variance(self):
    return _g.Lazy(lambda: (lambda a, b: a * b / ((a + b) ** 2 * (a + b + 1.0)))(*self._ab()), tuple(self.batch_shape),
                   self.dtype, "BetaVariance")


class TransformedDistribution(RandomVariable):
  """TransformedDistribution(distribution, bijector): Y = bijector.forward(X), X ~ distribution
  (tf.contrib.distributions.TransformedDistribution as Edward wraps it; the result of `ed.transform`,
  util/random_variables.py:856-917). `support` is an instance attribute the caller may set."""

  def __init__(self, distribution, bijector=None, validate_args=False, name="TransformedDistribution", **kwargs):
    if bijector is None:
      raise ValueError("TransformedDistribution needs a bijector")
    self.distribution = distribution
    self.bijector = bijector
    self._args, self._kwargs = (distribution, bijector), dict(kwargs)
    super(TransformedDistribution, self).__init__(tuple(distribution.batch_shape), tuple(distribution.event_shape),
                                                  distribution.dtype, name=name, **kwargs)

  def _sample_np(self, sample_shape, feed=None):
    return np.asarray(self.bijector._forward_np(self.distribution._sample_np(sample_shape, feed)), self.dtype.np)
