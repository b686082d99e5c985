"""ed.HMC — Hamiltonian Monte Carlo over Empirical posteriors, edward/inferences/hmc.py:14-210, with the
transition executed by libedhmc on the GPU instead of a TensorFlow graph.

    inference = ed.HMC({w: qw, b: qb}, data={X: X_train, y: y_train})
    inference.run(step_size=0.6)          # or initialize() + update() loop

keeps the reference's surface: `initialize(step_size=0.25, n_steps=2, ...)` (hmc.py:45), `update()` →
{'t', 'accept_rate'} (monte_carlo.py:111-150), `run()` (inference.py:97-154), attributes `t`, `n_accept`,
`n_accept_over_t`, `n_iter`, `n_print`, `step_size`, `n_steps`, `latent_vars`, `data`, `reset`, `progbar`,
`train`. Transition t reads row max(t-1,0) of each Empirical's params and writes row t (hmc.py:81-85,121-126).
"""
from __future__ import annotations

import os

import numpy as np

from .. import _C
from .. import graph as _g
from ..glm import recognize
from ..util.graphs import get_seed
from .monte_carlo import MonteCarlo


class HMC(MonteCarlo):
  def __init__(self, *args, **kwargs):
    super(HMC, self).__init__(*args, **kwargs)
    self._sampler = None

  def initialize(self, step_size=0.25, n_steps=2, *args, **kwargs):
    """hmc.py:45-59. Extra keyword-only options of this implementation (all optional):
    `device` (CUDA device, default cuda:$LOCAL_RANK or the current device), `plan` ('auto' | 'persistent' |
    'stepwise'), `row_sharded` (data passed to this rank is one shard of the rows; defaults to True when
    torch.distributed is initialised with world_size > 1)."""
    self.step_size = step_size
    self.n_steps = n_steps
    self._device = kwargs.pop('device', None)
    self._plan = {'auto': _C.PLAN_AUTO, 'persistent': _C.PLAN_PERSISTENT, 'stepwise': _C.PLAN_STEPWISE}[
        kwargs.pop('plan', 'auto')]
    self._row_sharded = kwargs.pop('row_sharded', None)
    return super(HMC, self).initialize(*args, **kwargs)

  # ------------------------------------------------------------------------------------------
  def build_update(self):
    """hmc.py:61-130, built as a device-resident sampler rather than a graph."""
    try:
      self.latent_vars_unconstrained
    except AttributeError:
      raise ValueError("This implementation of HMC requires that all "
                       "variables have unconstrained support. Please "
                       "initialize with auto_transform=True to ensure "
                       "this. (if your variables already have unconstrained "
                       "support then doing this is a no-op).")
    import torch

    # the sampler works in the unconstrained space (hmc.py:81-97,132-159): every original latent is paired with the
    # Empirical store that holds its unconstrained samples
    stores = {z: self.latent_vars_unconstrained[self.transformations.get(z, z)] for z in self.latent_vars}
    model = recognize(stores, self.data)
    self._model = model
    if model.y_rv is None:
      y_val = np.zeros(0, np.int32)
    else:
      y_val = self.data[model.y_rv]
      if isinstance(y_val, _g.Tensor):
        y_val = _g.evaluate(y_val)
    self._y_value = y_val
    self._x_value = self._current_x({})
    self._x_key = self._x_identity(self._x_value)
    dev = self._device
    if dev is None:
      dev = "cuda:%d" % int(os.environ["LOCAL_RANK"]) if "LOCAL_RANK" in os.environ else "cuda"
    self._dev = dev
    self._seed_value = None
    self._sampler = self._make_sampler(self._x_value)

    # Re-home the Empirical stores: one packed [T_max, P] device buffer, each latent's Variable a view of its own
    # first rows (the reference allows stores of different lengths; n_iter is the shortest, monte_carlo.py:96-97).
    P = model.spec.n_params
    rows_max = self.n_iter
    for slot in model.slots:
      variables = slot.qz.get_variables()
      if not variables:
        raise TypeError("Empirical random variables must be directly parameterized by a tf.Variable "
                        "for HMC to update them (hmc.py:66-70).")
      rows_max = max(rows_max, int(variables[0].shape[0]))
    self._packed = torch.zeros(rows_max, P, dtype=self._sampler.dtype, device=self._sampler.dev)
    for slot in model.slots:
      var = slot.qz.get_variables()[0]
      rows = int(var.shape[0])
      view = self._packed[:rows, slot.offset] if slot.scalar and len(var.shape) == 1 \
          else self._packed[:rows, slot.offset:slot.offset + slot.size]
      var.rebind(view)
    return self._train

  def _make_sampler(self, x):
    """Builds the device-side sampler for design matrix `x` with everything build_update decided: device, plan, debug,
    the global row count and communicator when the rows are sharded, and the Philox seed. Used by build_update and by
    _maybe_rebind, so that a re-bound sampler keeps the seed and stays one shard of the same global problem."""
    import torch
    import torch.distributed as dist
    from ..engine import GLMSampler
    from ..util.graphs import sampler_seed

    dev = self._dev
    sharded = self._row_sharded
    if sharded is None:
      sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    n_global = None
    if sharded:
      cnt = torch.tensor([int(np.shape(x)[0])], dtype=torch.int64,
                         device=dev if dist.get_backend() == "nccl" else "cpu")
      dist.all_reduce(cnt)
      n_global = int(cnt.item())
    sampler = GLMSampler(self._model.spec, x, self._y_value, device=dev, plan=self._plan, debug=self.debug,
                         n_rows_global=n_global, dtype=getattr(torch, self._model.dtype))
    if sharded:
      sampler.init_comm(dist.get_world_size(), dist.get_rank())
    if self._model.prior_kinds is not None:
      sampler.set_prior_kinds(self._model.prior_kinds)
    if self._seed_value is None:
      bcast = None
      if sharded:
        def bcast(v):
          box = [v]
          dist.broadcast_object_list(box, src=0)
          return box[0]
      self._seed_value = sampler_seed(bcast)
    sampler.seed(self._seed_value)
    return sampler

  @staticmethod
  def _x_identity(x):
    """What makes two design matrices 'the same binding': the storage they live in, not the Python object."""
    try:
      import torch
      if isinstance(x, torch.Tensor):
        return ("torch", x.data_ptr(), tuple(x.shape), str(x.dtype))
    except ImportError:
      pass
    if isinstance(x, np.ndarray):
      return ("numpy", x.__array_interface__["data"][0], x.shape, str(x.dtype))
    return ("object", id(x))

  def _current_x(self, feed_dict):
    model = self._model
    if model.x_node is None:
      return np.ones((model.n_rows, 1), model.dtype)  # a scalar latent used directly as the predictor; 0 rows = no data
    node = model.x_node
    if node in feed_dict:
      return feed_dict[node]
    if node in self.data:
      return self.data[node]
    if isinstance(node, _g.Variable) and node.value_tensor() is not None:
      return node.value_tensor()
    return _g.evaluate(node)

  def _train(self, feed_dict=None):
    """One transition at the current `t` — the op `sess.run(self.train)` executes in the reference."""
    self._maybe_rebind(feed_dict or {})
    self._sampler.run(self._packed[:self.n_iter], self._t, 1, self.step_size, self.n_steps)

  def _maybe_rebind(self, feed_dict):
    node = self._model.x_node
    if node is None or not ("Placeholder" in node.op_type or node in feed_dict):
      return  # X is a Variable / constant / computed node bound at initialize(): nothing can have been re-fed
    x = self._current_x(feed_dict)
    key = self._x_identity(x)
    if key != self._x_key:
      # a different design matrix was fed for the placeholder (monte_carlo.py:134-136): upload it
      old = self._sampler
      self._x_value, self._x_key = x, key
      n_acc, _ = old.read_state()
      self._n_accept_base = getattr(self, "_n_accept_base", 0) + n_acc
      old.close()
      self._sampler = self._make_sampler(x)

  # counters ----------------------------------------------------------------------------------
  def _get_n_accept(self):
    if self._sampler is None:
      return 0
    return self._sampler.read_state()[0] + getattr(self, "_n_accept_base", 0)

  def _reset_n_accept(self):
    self._n_accept_base = 0
    if self._sampler is not None:
      self._sampler.reset()

  # driver ------------------------------------------------------------------------------------
  def _run_loop(self):
    """Inference.run's loop (inference.py:145-147) with the transitions between two progress reports
    executed as ONE device launch: the observable behaviour (params rows, t, n_accept, the progress lines
    at t == 1 and t % n_print == 0) is unchanged, the host round-trips per transition are gone."""
    t = self._t
    while t < self.n_iter:
      if self.n_print == 0:
        nxt = self.n_iter
      elif t == 0:
        nxt = 1
      else:
        nxt = min(self.n_iter, (t // self.n_print + 1) * self.n_print)
      self._maybe_rebind({})
      self._sampler.run(self._packed[:self.n_iter], t, nxt - t, self.step_size, self.n_steps)
      with np.errstate(divide='ignore', invalid='ignore'):
        accept_rate = np.float64(self._get_n_accept()) / np.float64(nxt - 1) if self.n_print != 0 else None
      t = nxt
      self._t = t
      if self.n_print != 0:
        self._log_scalars(t)
        self.print_progress({'t': t, 'accept_rate': accept_rate})
    if self.n_print == 0:
      self._sampler.read_state()  # synchronise: run() returns with the samples written

  # checkpoint / resume (the reference relies on tf.train.Saver over params/t/n_accept,
  # tests/inferences/saver_test.py) -----------------------------------------------------------------
  def state_dict(self):
    """The resumable chain state: the Empirical store, the iteration counter, n_accept and the Philox seed.
    Device draws are a pure function of (seed, t, element), so a resumed chain continues bit-identically."""
    return {"params": self._packed.detach().cpu().numpy().copy(), "t": int(self._t),
            "n_accept": int(self._get_n_accept()), "seed": self._seed_value,
            "step_size": float(self.step_size), "n_steps": int(self.n_steps)}

  def load_state_dict(self, state):
    import torch
    params = np.asarray(state["params"], self._model.dtype)
    if tuple(params.shape) != tuple(self._packed.shape):
      raise ValueError("checkpoint params have shape %s, expected %s" % (params.shape, tuple(self._packed.shape)))
    self._packed.copy_(torch.as_tensor(params).to(self._packed.device))
    _g.bump_device_epoch()
    self._sampler.reset()
    self._n_accept_base = int(state["n_accept"])
    self._t = int(state["t"])
    if state.get("seed") is not None:
      self._seed_value = int(state["seed"])
      self._sampler.seed(self._seed_value)

  def finalize(self):
    super(HMC, self).finalize()
