"""check_data / check_latent_vars with the reference's error conventions
(edward/util/random_variables.py:21-83) and `transform` (:856-917). copy() is a TF-graph rewrite and is not part of this
path: HMC conditions the model through the GLM recogniser instead (edward_b200/glm.py)."""
from __future__ import annotations

import numpy as np

from .. import graph as _g
from ..models.random_variable import RandomVariable


def _is_array_like(v):
  """numpy values as in the reference, plus torch tensors (host or device) so data can be handed over
  without a detour through numpy."""
  if isinstance(v, (float, list, int, np.ndarray, np.number, str)):
    return True
  try:
    import torch
    return isinstance(v, torch.Tensor)
  except ImportError:
    return False


def _is_placeholder(t):
  return isinstance(t, _g.Tensor) and "Placeholder" in t.op_type


def check_data(data):
  """random_variables.py:21-61."""
  if not isinstance(data, dict):
    raise TypeError("data must have type dict.")
  for key, value in data.items():
    if _is_placeholder(key):
      if isinstance(value, RandomVariable):
        raise TypeError("The value of a feed cannot be a ed.RandomVariable object. "
                        "Acceptable feed values include Python scalars, strings, lists, numpy ndarrays, "
                        "or TensorHandles.")
      elif isinstance(value, _g.Tensor):
        raise TypeError("The value of a feed cannot be a tf.Tensor object. "
                        "Acceptable feed values include Python scalars, strings, lists, numpy ndarrays, "
                        "or TensorHandles.")
    elif isinstance(key, (RandomVariable, _g.Tensor)):
      if isinstance(value, (RandomVariable, _g.Tensor)):
        if not key.shape.is_compatible_with(value.shape):
          raise TypeError("Key-value pair in data does not have same shape: {}, {}".format(key.shape, value.shape))
        elif key.dtype != value.dtype:
          raise TypeError("Key-value pair in data does not have same dtype: {}, {}".format(key.dtype, value.dtype))
      elif _is_array_like(value):
        if not key.shape.is_compatible_with(tuple(np.shape(value))):
          raise TypeError("Key-value pair in data does not have same shape: {}, {}".format(key.shape, np.shape(value)))
        elif isinstance(value, (np.ndarray, np.number)) and \
                not np.issubdtype(value.dtype, np.floating) and \
                not np.issubdtype(value.dtype, np.integer) and \
                not np.issubdtype(value.dtype, np.str_):
          raise TypeError("Data value has an invalid dtype: {}".format(value.dtype))
      else:
        raise TypeError("Data value has an invalid type: {}".format(type(value)))
    else:
      raise TypeError("Data key has an invalid type: {}".format(type(key)))


def check_latent_vars(latent_vars):
  """random_variables.py:64-83."""
  if not isinstance(latent_vars, dict):
    raise TypeError("latent_vars must have type dict.")
  for key, value in latent_vars.items():
    if not isinstance(key, (RandomVariable, _g.Tensor)):
      raise TypeError("Latent variable key has an invalid type: {}".format(type(key)))
    elif not isinstance(value, (RandomVariable, _g.Tensor)):
      raise TypeError("Latent variable value has an invalid type: {}".format(type(value)))
    elif not key.shape.is_compatible_with(value.shape):
      raise TypeError("Key-value pair in latent_vars does not have same shape: {}, {}".format(key.shape, value.shape))
    elif key.dtype != value.dtype:
      raise TypeError("Key-value pair in latent_vars does not have same dtype: {}, {}".format(key.dtype, value.dtype))


def transform(x, *args, **kwargs):
  """util/random_variables.py:856-917: the default map of a continuous random variable to the unconstrained space —
  (0,1) through the inverse sigmoid, (0,inf) through the inverse softplus, the real line unchanged."""
  from .. import bijectors as tfb
  from ..models.random_variables import TransformedDistribution
  if len(args) != 0 or kwargs.get('bijector', None) is not None:
    return TransformedDistribution(x, *args, **kwargs)
  try:
    support = x.support
  except AttributeError:
    raise AttributeError("'{}' object has no 'support' so cannot be transformed.".format(type(x).__name__))
  if support == '01':
    bij, new_support = tfb.Invert(tfb.Sigmoid()), 'real'
  elif support == 'nonnegative':
    bij, new_support = tfb.Invert(tfb.Softplus()), 'real'
  elif support in ('real', 'multivariate_real'):
    return x
  elif support == 'simplex':
    raise NotImplementedError("simplex-valued latents are outside the HMC/GLM path built here")
  else:
    raise ValueError("'transform' does not handle supports of type '{}'".format(support))
  new_x = TransformedDistribution(x, bij, *args, **kwargs)
  new_x.support = new_support
  return new_x
