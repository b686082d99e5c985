"""`tfshim` — the handful of TensorFlow 1.x symbols that scripts on the HMC hot path touch
(examples/bayesian_logistic_regression.py:16-20,41-48,65,76,98; docs/tex/iclr2017.tex:248-264;
tests/inferences/hmc_test.py), so such a script runs with `from edward_b200 import tfshim as tf` as its
only change. It is NOT TensorFlow: no sessions, no autodiff, no kernels — see graph.py.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import graph as _g
from .graph import (Constant, Placeholder, Tensor, TensorShape, Variable, constant, convert_to_tensor,  # noqa: F401
                    float32, float64, get_default_graph, global_variables_initializer, int32, int64,
                    reset_default_graph, variables_initializer)


def placeholder(dtype, shape=None, name=None):
  return Placeholder(dtype, shape)


def zeros(shape, dtype=float32, name=None):
  dtype = _g.as_dtype(dtype)
  return Constant(np.zeros([int(s) for s in np.atleast_1d(shape)] if np.ndim(shape) else ([int(shape)] if shape != [] and shape != () else []), dtype.np), dtype)


def ones(shape, dtype=float32, name=None):
  dtype = _g.as_dtype(dtype)
  return Constant(np.ones([int(s) for s in np.atleast_1d(shape)] if np.ndim(shape) else ([int(shape)] if shape != [] and shape != () else []), dtype.np), dtype)


def zeros_like(t, dtype=None):
  t = convert_to_tensor(t)
  return Constant(np.zeros(tuple(t.shape), (dtype or t.dtype).np), dtype or t.dtype)


def ones_like(t, dtype=None):
  t = convert_to_tensor(t)
  return Constant(np.ones(tuple(t.shape), (dtype or t.dtype).np), dtype or t.dtype)


_named_variables = {}


def get_variable(name, shape=None, dtype=float32, initializer=None, trainable=True):
  """tf.get_variable with TF's default glorot_uniform initializer (drawn from numpy's global RNG, which
  ed.set_seed seeds)."""
  if name in _named_variables and _named_variables[name][0] is get_default_graph():
    raise ValueError("Variable %s already exists, disallowed." % name)
  shape = [int(s) for s in (shape or [])]
  dtype = _g.as_dtype(dtype)
  if initializer is None:
    if len(shape) == 0:
      fan_in = fan_out = 1
    elif len(shape) == 1:
      fan_in = fan_out = shape[0]
    else:
      fan_in, fan_out = shape[-2], shape[-1]
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    init = np.random.uniform(-limit, limit, size=shape).astype(dtype.np)
  elif callable(initializer):
    init = np.asarray(initializer(shape), dtype.np)
  else:
    init = np.asarray(_g.evaluate(initializer) if isinstance(initializer, Tensor) else initializer, dtype.np)
  v = Variable(init, trainable=trainable, dtype=dtype, name=name)
  _named_variables[name] = (get_default_graph(), v)
  return v


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
  dtype = _g.as_dtype(dtype)
  shape = tuple(int(s) for s in shape)
  return _g.Lazy(lambda: mean + stddev * np.random.standard_normal(shape), shape, dtype, "RandomStandardNormal")


def sigmoid(x):
  return _g.Unary(x, lambda v: 1.0 / (1.0 + np.exp(-v)), "Sigmoid")


def exp(x):
  return _g.Unary(x, np.exp, "Exp")


def log(x):
  return _g.Unary(x, np.log, "Log")


def square(x):
  return _g.Unary(x, np.square, "Square")


def sqrt(x):
  return _g.Unary(x, np.sqrt, "Sqrt")


def stack(values, axis=0):
  return _g.Stack(values)


def reduce_mean(x, axis=None):
  x = convert_to_tensor(x)
  shape = () if axis is None else tuple(d for i, d in enumerate(x.shape) if i != axis)
  return _g.Lazy(lambda: np.mean(_g.evaluate(x), axis=axis), shape, x.dtype, "Mean")


def reduce_sum(x, axis=None):
  x = convert_to_tensor(x)
  shape = () if axis is None else tuple(d for i, d in enumerate(x.shape) if i != axis)
  return _g.Lazy(lambda: np.sum(_g.evaluate(x), axis=axis), shape, x.dtype, "Sum")


def cast(x, dtype):
  x = convert_to_tensor(x)
  dtype = _g.as_dtype(dtype)
  return _g.Lazy(lambda: _g.evaluate(x), tuple(x.shape), dtype, "Cast")


def set_random_seed(seed):
  get_default_graph().seed = seed


class _Flags(object):
  """tf.flags: DEFINE_* + FLAGS, parsed from sys.argv on first attribute access."""

  def __init__(self):
    object.__setattr__(self, "_parser", argparse.ArgumentParser(add_help=False))
    object.__setattr__(self, "_values", None)
    object.__setattr__(self, "_defaults", {})

  def _define(self, name, default, help, type_):
    self._parser.add_argument("--" + name, default=default, type=type_, help=help)
    self._defaults[name] = default
    object.__setattr__(self, "_values", None)

  def __getattr__(self, name):
    if self._values is None:
      vals, _ = self._parser.parse_known_args(sys.argv[1:])
      object.__setattr__(self, "_values", vals)
    try:
      return getattr(self._values, name)
    except AttributeError:
      raise AttributeError(name)

  def __setattr__(self, name, value):
    if self._values is None:
      self.__getattr__(name)
    setattr(self._values, name, value)


class _FlagsModule(object):
  def __init__(self):
    self.FLAGS = _Flags()

  def DEFINE_integer(self, name, default, help=""):
    self.FLAGS._define(name, default, help, int)

  def DEFINE_float(self, name, default, help=""):
    self.FLAGS._define(name, default, help, float)

  def DEFINE_string(self, name, default, help=""):
    self.FLAGS._define(name, default, help, str)

  def DEFINE_boolean(self, name, default, help=""):
    self.FLAGS._define(name, default, help, lambda s: str(s).lower() in ("1", "true", "yes"))

  DEFINE_bool = DEFINE_boolean


flags = _FlagsModule()


class _App(object):
  @staticmethod
  def run(main=None, argv=None):
    main = main or sys.modules["__main__"].main
    sys.exit(main(argv or sys.argv))


app = _App()


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None):
  dtype = _g.as_dtype(dtype)
  shape = tuple(int(s) for s in shape)
  return _g.Lazy(lambda: np.random.uniform(minval, maxval, size=shape), shape, dtype, "RandomUniform")


class _Nn(object):
  sigmoid = staticmethod(sigmoid)

  @staticmethod
  def moments(x, axes, name=None):
    """tf.nn.moments: (mean, population variance) over `axes`."""
    x = convert_to_tensor(x)
    ax = tuple(np.atleast_1d(axes).tolist())
    shape = tuple(d for i, d in enumerate(x.shape) if i not in ax)
    return (_g.Lazy(lambda: np.mean(_g.evaluate(x), axis=ax), shape, x.dtype, "Mean"),
            _g.Lazy(lambda: np.var(_g.evaluate(x), axis=ax), shape, x.dtype, "Variance"))

  @staticmethod
  def softplus(x):
    return _g.Unary(x, lambda v: np.logaddexp(0.0, v), "Softplus")


nn = _Nn()


class _Namespace(object):
  pass


def _contrib():
  from . import bijectors as _bij
  c = _Namespace()
  c.distributions = _Namespace()
  c.distributions.bijectors = _bij
  return c


contrib = _contrib()  # tf.contrib.distributions.bijectors.{Softplus, Sigmoid, Invert}
