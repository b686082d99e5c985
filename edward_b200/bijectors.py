"""The bijectors Edward's `transform` uses (tf.contrib.distributions.bijectors.{Sigmoid, Softplus, Invert},
edward/util/random_variables.py:856-917): forward / inverse as lazy graph nodes, plus numpy kernels for read-outs.
The sampler itself never calls these: the Jacobian term of a transformed latent is folded into the device log joint
(include/edhmc.h edhmc_set_prior_kinds); the front-end uses them to map the stored unconstrained samples back."""
from __future__ import annotations

import numpy as np

from . import graph as _g


class Bijector(object):
  name = "bijector"

  def _forward_np(self, x):
    raise NotImplementedError

  def _inverse_np(self, y):
    raise NotImplementedError

  def forward(self, x):
    x = _g.convert_to_tensor(x)
    return _g.Unary(x, self._forward_np, type(self).__name__ + "Forward")

  def inverse(self, y):
    y = _g.convert_to_tensor(y)
    return _g.Unary(y, self._inverse_np, type(self).__name__ + "Inverse")


class Sigmoid(Bijector):
  """Y = 1 / (1 + exp(-X)): real line -> (0, 1)."""

  def _forward_np(self, x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, np.float64)))

  def _inverse_np(self, y):
    y = np.asarray(y, np.float64)
    return np.log(y) - np.log1p(-y)


class Softplus(Bijector):
  """Y = log(1 + exp(X)): real line -> (0, inf)."""

  def _forward_np(self, x):
    return np.logaddexp(0.0, np.asarray(x, np.float64))

  def _inverse_np(self, y):
    y = np.asarray(y, np.float64)
    return y + np.log(-np.expm1(-y))


class Invert(Bijector):
  def __init__(self, bijector):
    self.bijector = bijector

  def _forward_np(self, x):
    return self.bijector._inverse_np(x)

  def _inverse_np(self, y):
    return self.bijector._forward_np(y)
