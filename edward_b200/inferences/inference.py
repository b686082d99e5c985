"""Inference base class — the drop-in surface of edward/inferences/inference.py:20-338: data binding,
`run()` driver, the `t` counter, n_print / Progbar bookkeeping, reset ops, the auto-transform identity
map for real-valued latents."""
from __future__ import annotations

import abc
import os
from datetime import datetime

import numpy as np

from .. import graph as _g
from ..models.random_variable import RandomVariable
from ..util.progbar import Progbar
from ..util.random_variables import _is_array_like, _is_placeholder, check_data, check_latent_vars


class Counter(_g.Tensor):
  """Integer state of the inference (`t`, `n_accept`): a tf.Variable in the reference (inference.py:211,
  monte_carlo.py:100). `getter`/`resetter` connect it to wherever the value really lives."""
  op_type = "VariableV2"

  def __init__(self, getter, resetter, name):
    self._getter, self._resetter, self.name = getter, resetter, name
    super(Counter, self).__init__((), _g.int32)

  def _eval(self, feed):
    return np.int32(self._getter())

  def reset(self):
    self._resetter()

  def __int__(self):
    return int(self._getter())


class Inference(abc.ABC):
  def __init__(self, latent_vars=None, data=None):
    """inference.py:54-95."""
    if latent_vars is None:
      latent_vars = {}
    if data is None:
      data = {}
    check_latent_vars(latent_vars)
    self.latent_vars = latent_vars
    check_data(data)
    self.data = {}
    for key, value in data.items():
      if _is_placeholder(key):
        self.data[key] = value
      elif isinstance(key, (RandomVariable, _g.Tensor)):
        if isinstance(value, (RandomVariable, _g.Tensor)):
          self.data[key] = value
        elif _is_array_like(value):
          # stored once with the key's dtype (inference.py:88-95: placeholder(key.dtype) → Variable)
          if isinstance(value, (float, list, int, np.ndarray, np.number, str)):
            self.data[key] = np.asarray(value).astype(key.dtype.np)
          else:
            self.data[key] = value  # torch tensor: cast on the device when the sampler binds it

  def run(self, variables=None, use_coordinator=True, *args, **kwargs):
    """inference.py:97-154: initialize, (re-)initialise variables, n_iter × (update, print_progress),
    finalize."""
    self.initialize(*args, **kwargs)
    if variables is None:
      init = _g.global_variables_initializer()
    else:
      init = _g.variables_initializer(variables)
    init.run()
    self._run_loop()
    self.finalize()

  def _run_loop(self):
    for _ in range(self.n_iter):
      info_dict = self.update()
      self.print_progress(info_dict)

  @abc.abstractmethod
  def initialize(self, n_iter=1000, n_print=None, scale=None, auto_transform=True, logdir=None,
                 log_timestamp=True, log_vars=None, debug=False):
    """inference.py:157-285."""
    self.n_iter = int(n_iter)
    if n_print is None:
      self.n_print = int(n_iter / 100)
    else:
      self.n_print = n_print
    self.progbar = Progbar(self.n_iter)
    self._t = 0
    self.t = Counter(lambda: self._t, self._reset_t, "iteration")

    if scale is None:
      scale = {}
    elif not isinstance(scale, dict):
      raise TypeError("scale must be a dict object.")
    self.scale = scale

    # inference.py:223-264. map from original latent vars to unconstrained versions: a latent whose support differs
    # from its posterior's is moved to the unconstrained space by ed.transform; an Empirical ('points') posterior is
    # taken to already live there, and a constrained view of it (params pushed through the inverse map) is what
    # `self.latent_vars[z]` returns afterwards.
    self.transformations = {}
    if auto_transform:
      from ..models.empirical import Empirical
      from ..util.random_variables import transform
      latent_vars = self.latent_vars.copy()
      self.latent_vars = {}
      self.latent_vars_unconstrained = {}
      for z, qz in latent_vars.items():
        if hasattr(z, 'support') and hasattr(qz, 'support') and z.support != qz.support and qz.support != 'point':
          z_unconstrained = transform(z)
          self.transformations[z] = z_unconstrained
          if qz.support == "points":
            qz_unconstrained = qz
          else:
            raise NotImplementedError("only Empirical posteriors are supported on this path")
          self.latent_vars_unconstrained[z_unconstrained] = qz_unconstrained
          if z_unconstrained is not z:
            qz_constrained = Empirical(params=z_unconstrained.bijector.inverse(qz_unconstrained.params))
          else:
            qz_constrained = qz_unconstrained
          self.latent_vars[z] = qz_constrained
        else:
          self.latent_vars[z] = qz
          self.latent_vars_unconstrained[z] = qz
      del latent_vars

    if logdir is not None:
      # TensorBoard event files, as the reference's tf.summary.FileWriter writes them (inference.py:266-277): scalars
      # registered by the subclass (MonteCarlo: "n_accept", monte_carlo.py:106-109) and, per log_vars, one
      # "parameter/<name>" scalar or histogram per variable (inference.py:340-376), every n_print iterations.
      self.logging = True
      logdir = os.path.expanduser(logdir)
      if log_timestamp:
        logdir = os.path.join(logdir, datetime.strftime(datetime.utcnow(), "%Y%m%d_%H%M%S"))
      os.makedirs(logdir, exist_ok=True)
      from torch.utils.tensorboard import SummaryWriter
      self.logdir = logdir
      self.train_writer = SummaryWriter(log_dir=logdir)
      self._set_log_variables(log_vars)
    else:
      self.logging = False
    self.debug = debug
    self.reset = [_g.variables_initializer([self.t])]

  def _reset_t(self):
    self._t = 0

  def _set_log_variables(self, log_vars=None):
    """inference.py:340-376: None = every variable of the data and of the latent variables / their posteriors
    (for MonteCarlo: the Empirical stores), [] = none."""
    if log_vars is None:
      log_vars = []
      for key, value in self.latent_vars.items():
        for rv in (key, value):
          if hasattr(rv, "get_variables"):
            log_vars += list(rv.get_variables())
      seen, uniq = set(), []
      for v in log_vars:
        if id(v) not in seen:
          seen.add(id(v))
          uniq.append(v)
      log_vars = uniq
    self._log_vars = list(log_vars)

  def _summary_scalars(self):
    """Scalars a subclass adds to every summary (name -> value)."""
    return {}

  def _log_scalars(self, t):
    """What `sess.run(self.summarize)` + `train_writer.add_summary(summary, t)` do (monte_carlo.py:143-146)."""
    if not self.logging or self.n_print == 0 or not (t == 1 or t % self.n_print == 0):
      return
    for name, value in self._summary_scalars().items():
      self.train_writer.add_scalar(name, value, global_step=t)
    for i, var in enumerate(self._log_vars):
      name = (getattr(var, "name", None) or "Variable_%d" % i).replace(':', '/')
      val = np.asarray(var.numpy() if hasattr(var, "numpy") else _g.evaluate(var))
      if val.ndim == 0 or (val.ndim == 1 and val.shape[0] == 1):
        self.train_writer.add_scalar("parameter/{}".format(name), float(val.reshape(-1)[0]), global_step=t)
      else:
        self.train_writer.add_histogram("parameter/{}".format(name), val, global_step=t)

  @abc.abstractmethod
  def update(self, feed_dict=None):
    raise NotImplementedError

  def print_progress(self, info_dict):
    """inference.py:322-332."""
    if self.n_print != 0:
      t = info_dict['t']
      if t == 1 or t % self.n_print == 0:
        self.progbar.update(t)

  def finalize(self):
    """inference.py:334-338."""
    if self.logging:
      self.train_writer.close()
