"""RandomVariable — the API shape of edward/models/random_variable.py:18-314 for the distributions the
HMC hot path uses. A random variable is a graph node standing for one sample of itself, so model code
can write `ed.dot(X, w) + b` with `w`, `b` random variables (random_variable.py:261-314)."""
from __future__ import annotations

import numpy as np

from .. import graph as _g


class RandomVariable(_g.Tensor):
  support = None
  op_type = "RandomVariable"

  def __init__(self, batch_shape, event_shape, dtype, sample_shape=(), value=None, name=None, collections=None):
    if isinstance(sample_shape, (int, np.integer)):
      sample_shape = (int(sample_shape),)
    self._sample_shape = _g.TensorShape(tuple(int(s) for s in sample_shape))
    self._batch_shape = _g.TensorShape(tuple(batch_shape))
    self._event_shape = _g.TensorShape(tuple(event_shape))
    self.name = name or type(self).__name__
    shape = tuple(self._sample_shape) + tuple(self._batch_shape) + tuple(self._event_shape)
    super(RandomVariable, self).__init__(shape, dtype)
    if value is not None:
      t_value = _g.convert_to_tensor(value, self.dtype)
      if not t_value.shape.is_compatible_with(shape):
        raise ValueError("Incompatible shape for initialization argument 'value'. Expected %s, got %s."
                         % (shape, tuple(t_value.shape)))
      self._value = t_value
    else:
      self._value = _g.Lazy(lambda feed: self._sample_np(tuple(self._sample_shape), feed), shape, self.dtype, "Sample",
                            wants_feed=True)
    _g.get_default_graph().random_variables.append(self)

  # shapes (random_variable.py:140-170)
  @property
  def sample_shape(self):
    return self._sample_shape

  @property
  def batch_shape(self):
    return self._batch_shape

  @property
  def event_shape(self):
    return self._event_shape

  def value(self):
    """The tensor this random variable stands for (random_variable.py:253-259)."""
    return self._value

  def _eval(self, feed):
    if id(self) in feed:
      return np.asarray(feed[id(self)], self.dtype.np)
    return self._value._eval(feed)

  def get_ancestors(self, collection=None):
    raise NotImplementedError("graph traversal utilities are outside the HMC hot path")

  # to be provided by the distribution
  def _sample_np(self, sample_shape, feed=None):
    raise NotImplementedError("sample is not implemented for {0}".format(type(self).__name__))

  def log_prob(self, value):
    raise NotImplementedError

  def sample(self, sample_shape=(), seed=None):
    if isinstance(sample_shape, (int, np.integer)):
      sample_shape = (int(sample_shape),)
    sample_shape = tuple(sample_shape)
    shape = sample_shape + tuple(self._batch_shape) + tuple(self._event_shape)
    return _g.Lazy(lambda feed: self._sample_np(sample_shape, feed), shape, self.dtype, "Sample", wants_feed=True)

  def __hash__(self):
    return id(self)

  def __eq__(self, other):
    return self is other

  def __ne__(self, other):
    return self is not other
