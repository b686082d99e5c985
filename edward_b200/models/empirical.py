"""Empirical(params) — the sample store HMC writes (edward/models/empirical.py:16-125). `params` is
normally a Variable; once an `ed.HMC` adopts it the storage is a device tensor and mean/stddev/sample are
computed on the device from it."""
from __future__ import annotations

import numpy as np

from .. import graph as _g
from .random_variable import RandomVariable


class Empirical(RandomVariable):
  support = 'points'

  def __init__(self, params, validate_args=False, allow_nan_stats=True, name="Empirical", **kwargs):
    self._params = _g.convert_to_tensor(params)
    self._args, self._kwargs = (params,), dict(kwargs)
    shape = tuple(self._params.shape)
    self._n = shape[0] if len(shape) > 0 else 1
    super(Empirical, self).__init__((), shape[1:], self._params.dtype, name=name, **kwargs)

  @property
  def params(self):
    """Distribution parameter (empirical.py:65-68)."""
    return self._params

  @property
  def n(self):
    return self._n

  def get_variables(self, collection=None):
    """Variables this random variable depends on (util/random_variables.py:726); HMC writes the first."""
    return [self._params] if isinstance(self._params, _g.Variable) else []

  def _device_params(self):
    if isinstance(self._params, _g.Variable):
      return self._params.value_tensor()
    return None

  def mean(self):
    """empirical.py:87-88 — reduce_mean(params, 0); on the device when the store lives there."""
    def fn():
      t = self._device_params()
      if t is not None:
        return t.mean(dim=0).cpu().numpy()
      return np.mean(_g.evaluate(self._params), axis=0)
    return _g.Lazy(fn, tuple(self.event_shape), self.dtype, "Mean")

  def stddev(self):
    """empirical.py:90-93 — sqrt(reduce_mean(square(params - mean), 0)) (population standard deviation)."""
    def fn():
      t = self._device_params()
      if t is not None:
        return (t - t.mean(dim=0)).square().mean(dim=0).sqrt().cpu().numpy()
      p = _g.evaluate(self._params)
      return np.sqrt(np.mean(np.square(p - np.mean(p, axis=0)), axis=0))
    return _g.Lazy(fn, tuple(self.event_shape), self.dtype, "Stddev")

  def variance(self):
    sd = self.stddev()
    return _g.Lazy(lambda: np.square(_g.evaluate(sd)), tuple(self.event_shape), self.dtype, "Variance")

  def _sample_np(self, sample_shape, feed=None):
    """empirical.py:98-110 — rows gathered at uniformly drawn indices."""
    n = int(np.prod(sample_shape)) if len(sample_shape) else 1
    t = self._device_params()
    if len(self._params.shape) == 0:
      return np.tile(_g.evaluate(self._params), sample_shape)
    idx = np.random.randint(0, self._n, size=n)
    if t is not None and n > 4096:
      import torch
      rows = t[torch.as_tensor(idx, device=t.device)].cpu().numpy()  # large draws: gather on the device
    elif t is not None:
      rows = self._params.host_view()[idx]  # host mirror, refreshed after the next device write (graph.device_epoch)
    else:
      rows = _g.evaluate(self._params)[idx]
    return rows.reshape(tuple(sample_shape) + tuple(self.event_shape)).astype(self.dtype.np)
