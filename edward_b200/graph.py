"""A deliberately tiny symbolic layer: just enough expression nodes to write the models of the HMC hot
path the way an Edward script does (`ed.dot(X, w) + b`, constants, placeholders, variables) and to let
`ed.HMC` recognise them. It replaces TensorFlow's graph only as a *description*: no arithmetic of the
sampler runs through it (that is libedhmc's job), and anything the recogniser does not know raises
NotImplementedError instead of falling back to a slow path.

Nodes evaluate eagerly with numpy for read-outs outside the hot path (e.g. the example's posterior
predictive plot, examples/bayesian_logistic_regression.py:74-77).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


class DType(object):
  def __init__(self, name, np_dtype):
    self.name = name
    self.np = np.dtype(np_dtype)

  def __repr__(self):
    return "tf." + self.name

  def __eq__(self, other):
    if isinstance(other, DType):
      return self.np == other.np
    try:
      return self.np == np.dtype(other)
    except TypeError:
      return False

  def __ne__(self, other):
    return not self.__eq__(other)

  def __hash__(self):
    return hash(self.np)

  @property
  def as_numpy_dtype(self):
    return self.np.type


float32 = DType("float32", np.float32)
float64 = DType("float64", np.float64)
int32 = DType("int32", np.int32)
int64 = DType("int64", np.int64)
_DTYPES = [float32, float64, int32, int64]


def as_dtype(d) -> DType:
  if isinstance(d, DType):
    return d
  for t in _DTYPES:
    if t == d:
      return t
  raise TypeError("unsupported dtype %r" % (d,))


class TensorShape(tuple):
  """Static shape; None marks an unknown dimension."""

  def as_list(self):
    return list(self)

  def is_compatible_with(self, other):
    other = tuple(other)
    if len(self) != len(other):
      return False
    return all(a is None or b is None or a == b for a, b in zip(self, other))

  def concatenate(self, other):
    return TensorShape(tuple(self) + tuple(other))

  @property
  def ndims(self):
    return len(self)


class _Graph(object):
  """The default 'graph': a registry of variables and random variables (for set_seed's guard and
  global_variables_initializer)."""

  def __init__(self):
    self.variables = []
    self.random_variables = []
    self.n_nodes = 0
    self.seed = None


_default_graph = _Graph()


def get_default_graph() -> _Graph:
  return _default_graph


def reset_default_graph():
  global _default_graph
  _default_graph = _Graph()


class Tensor(object):
  """Base expression node."""
  op_type = "Tensor"

  def __init__(self, shape, dtype):
    self._shape = TensorShape(tuple(shape))
    self._dtype = as_dtype(dtype)
    get_default_graph().n_nodes += 1

  @property
  def shape(self):
    return self._shape

  def get_shape(self):
    return self._shape

  @property
  def dtype(self):
    return self._dtype

  def _eval(self, feed):
    raise NotImplementedError

  def eval(self, feed_dict=None, session=None):
    return evaluate(self, feed_dict)

  # operator overloads used by model code (edward/models/random_variable.py:261-314 forwards them to TF)
  def __add__(self, other):
    return Add(self, other)

  def __radd__(self, other):
    return Add(other, self)

  def __sub__(self, other):
    return Sub(self, other)

  def __rsub__(self, other):
    return Sub(other, self)

  def __mul__(self, other):
    return Mul(self, other)

  def __rmul__(self, other):
    return Mul(other, self)

  def __truediv__(self, other):
    return Div(self, other)

  def __neg__(self):
    return Mul(constant(-1.0, self.dtype), self)

  def __getitem__(self, key):
    probe = np.empty(tuple(d if d is not None else 1 for d in self.shape), np.bool_)[key]
    return Lazy(lambda feed: self._eval(feed)[key], probe.shape, self.dtype, "StridedSlice", wants_feed=True)

  __array_priority__ = 100


def evaluate(node, feed_dict=None):
  feed = {}
  if feed_dict:
    for k, v in feed_dict.items():
      feed[id(convert_to_tensor(k) if not isinstance(k, Tensor) else k)] = v
  return convert_to_tensor(node)._eval(feed)


class Constant(Tensor):
  op_type = "Const"

  def __init__(self, value, dtype=None):
    arr = np.asarray(value)
    if dtype is None:
      dtype = float32 if arr.dtype.kind == "f" else (int32 if arr.dtype.kind in "iub" else arr.dtype)
    dtype = as_dtype(dtype)
    self.value = np.asarray(arr, dtype.np)
    super(Constant, self).__init__(self.value.shape, dtype)

  def _eval(self, feed):
    return self.value


def constant(value, dtype=None):
  return Constant(value, dtype)


class Placeholder(Tensor):
  op_type = "Placeholder"

  def __init__(self, dtype, shape=None):
    super(Placeholder, self).__init__(tuple(shape) if shape is not None else (), dtype)

  def _eval(self, feed):
    if id(self) not in feed:
      raise ValueError("You must feed a value for placeholder tensor with dtype %r and shape %r"
                       % (self.dtype, tuple(self.shape)))
    return np.asarray(feed[id(self)], self.dtype.np)


# Host mirrors of device-resident variables are cached between device writes: every launch that can write a sample
# store (engine.GLMSampler.run / sgmcmc_run / run_chains, Variable.load / rebind, checkpoint restore) bumps this epoch.
# The reference's example draws 100 single rows from the Empirical store every 10 iterations
# (examples/bayesian_logistic_regression.py:91-96); without the mirror each draw is a host-device round trip.
_device_epoch = [0]


def bump_device_epoch():
  _device_epoch[0] += 1


def device_epoch():
  return _device_epoch[0]


class Variable(Tensor):
  """Mutable tensor. Until a sampler adopts it, it lives on the host; `ed.HMC.initialize` re-homes the
  Empirical parameter variables into device memory (`rebind`) so that the sample store stays on the
  GPU (hmc.py:121-126 writes it with scatter_update)."""
  op_type = "VariableV2"

  def __init__(self, initial_value=None, trainable=True, dtype=None, name=None, collections=None):
    if isinstance(initial_value, Tensor):
      init = np.array(evaluate(initial_value))
      if dtype is None:
        dtype = initial_value.dtype
    else:
      arr = np.asarray(initial_value)
      if dtype is None:
        dtype = float32 if arr.dtype == np.float32 else (arr.dtype if arr.dtype.kind != "b" else int32)
      init = np.array(arr)
    dtype = as_dtype(dtype)
    self.initial_value = np.asarray(init, dtype.np)
    self.name = name
    self.trainable = trainable
    self._storage = None  # torch tensor (device) once adopted by a sampler
    self._mirror, self._mirror_epoch = None, -1  # host copy of _storage, valid while the device epoch is unchanged
    self._host = self.initial_value.copy()
    super(Variable, self).__init__(self.initial_value.shape, dtype)
    if collections is None or len(collections) > 0:
      get_default_graph().variables.append(self)

  # -- storage -------------------------------------------------------------------------------
  def rebind(self, storage):
    """Adopt device storage (a torch tensor view of the right shape); current contents are copied in."""
    import torch
    cur = self.numpy()  # the CURRENT contents: after an earlier adoption they live on the device, not in _host
    storage.copy_(torch.as_tensor(np.ascontiguousarray(cur)).to(storage.device, storage.dtype).reshape(storage.shape))
    self._storage = storage
    bump_device_epoch()

  def value_tensor(self):
    """Device tensor if adopted, else None."""
    return self._storage

  def host_view(self):
    """Current contents as a host array that must not be modified (shared with the mirror cache)."""
    if self._storage is None:
      return self._host
    if self._mirror_epoch != _device_epoch[0]:
      m = self._storage.detach().cpu().numpy().reshape(tuple(self.shape)).astype(self.dtype.np, copy=False)
      self._mirror = m.copy() if self._storage.device.type == "cpu" else m  # (.cpu() of a CPU tensor aliases it)
      self._mirror_epoch = _device_epoch[0]
    return self._mirror

  def numpy(self):
    if self._storage is not None:
      return self.host_view().copy()
    return self._host

  def load(self, value):
    value = np.asarray(value, self.dtype.np).reshape(tuple(self.shape))
    if self._storage is not None:
      import torch
      self._storage.copy_(torch.as_tensor(value).to(self._storage.device, self._storage.dtype).reshape(self._storage.shape))
      bump_device_epoch()
    else:
      self._host = value.copy()

  def assign(self, value):
    self.load(evaluate(value) if isinstance(value, Tensor) else value)
    return self

  @property
  def initializer(self):
    return _InitOp([self])

  def _eval(self, feed):
    return self.numpy()


class _InitOp(object):
  def __init__(self, variables, extra=()):
    self.variables = list(variables)
    self.extra = list(extra)

  def run(self, feed_dict=None, session=None):
    for v in self.variables:
      v.load(v.initial_value)
    for fn in self.extra:
      fn()


def global_variables_initializer():
  return _InitOp(get_default_graph().variables)


def variables_initializer(var_list):
  return _InitOp([v for v in var_list if isinstance(v, Variable)],
                 [v.reset for v in var_list if not isinstance(v, Variable) and hasattr(v, "reset")])


def convert_to_tensor(x, dtype=None):
  if isinstance(x, Tensor):
    return x
  if hasattr(x, "value") and callable(getattr(x, "value")) and hasattr(x, "log_prob"):
    return x.value()  # a RandomVariable stands for its sample tensor
  return Constant(x, dtype)


def _bshape(a, b):
  try:
    return np.broadcast_shapes(tuple(d if d is not None else 1 for d in a), tuple(d if d is not None else 1 for d in b))
  except ValueError:
    raise ValueError("Dimensions must be equal: %r vs %r" % (tuple(a), tuple(b)))


class _Binary(Tensor):
  fn = None

  def __init__(self, a, b):
    self.a = convert_to_tensor(a)
    self.b = convert_to_tensor(b, self.a.dtype if not isinstance(b, Tensor) else None)
    if not isinstance(a, Tensor) and isinstance(b, Tensor):
      self.a = convert_to_tensor(a, self.b.dtype)
    super(_Binary, self).__init__(_bshape(self.a.shape, self.b.shape), self.a.dtype)

  def _eval(self, feed):
    return type(self).fn(self.a._eval(feed), self.b._eval(feed)).astype(self.dtype.np)


class Add(_Binary):
  op_type = "Add"
  fn = staticmethod(np.add)


class Sub(_Binary):
  op_type = "Sub"
  fn = staticmethod(np.subtract)


class Mul(_Binary):
  op_type = "Mul"
  fn = staticmethod(np.multiply)


class Div(_Binary):
  op_type = "RealDiv"
  fn = staticmethod(np.divide)


class Dot(Tensor):
  """`ed.dot` node (util/tensorflow.py:10-45): matrix·vector or vector·matrix, result 1-D."""
  op_type = "Dot"

  def __init__(self, x, y):
    self.x = convert_to_tensor(x)
    self.y = convert_to_tensor(y)
    if len(self.x.shape) == 1:
      n = self.y.shape[1] if len(self.y.shape) == 2 else None
    else:
      n = self.x.shape[0]
    super(Dot, self).__init__((n,), self.x.dtype)

  def _eval(self, feed):
    x = self.x._eval(feed)
    y = self.y._eval(feed)
    if not np.all(np.isfinite(x)) or not np.all(np.isfinite(y)):
      raise ValueError("InvalidArgumentError: Tensor had NaN or Inf values")
    if x.ndim == 1:
      return np.matmul(x[None, :], y).reshape(-1)
    return np.matmul(x, y[:, None]).reshape(-1)


class Unary(Tensor):
  def __init__(self, a, fn, op_type):
    self.a = convert_to_tensor(a)
    self.fn = fn
    self.op_type = op_type
    super(Unary, self).__init__(self.a.shape, self.a.dtype)

  def _eval(self, feed):
    return self.fn(self.a._eval(feed)).astype(self.dtype.np)


class Stack(Tensor):
  op_type = "Pack"

  def __init__(self, values):
    self.values = [convert_to_tensor(v) for v in values]
    super(Stack, self).__init__((len(self.values),) + tuple(self.values[0].shape), self.values[0].dtype)

  def _eval(self, feed):
    return np.stack([v._eval(feed) for v in self.values])


class Lazy(Tensor):
  """A node whose value is produced by a Python callable at evaluation time (random draws, Empirical
  statistics read back from the device store)."""

  def __init__(self, fn, shape, dtype, op_type="Lazy", wants_feed=False):
    self.fn = fn
    self.op_type = op_type
    self.wants_feed = wants_feed  # fn(feed) receives the internal {id(node): value} feed of this evaluation
    super(Lazy, self).__init__(shape, dtype)

  def _eval(self, feed):
    return np.asarray(self.fn(feed) if self.wants_feed else self.fn(), self.dtype.np)


def eval_in(node, feed):
  """Evaluates `node` inside an ongoing evaluation (feed = internal {id(node): value} dict)."""
  return convert_to_tensor(node)._eval(feed or {})
