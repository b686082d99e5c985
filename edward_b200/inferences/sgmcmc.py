"""ed.SGLD and ed.SGHMC (edward/inferences/sgld.py:13-121, sghmc.py:13-130) on the same fused log-joint gradient
kernel as ed.HMC: one streaming pass over the (mini-batch of the) data per iteration, the O(P) update on the
device, samples written to the Empirical store in place. `scale` (inference.py:170-175, sgld.py:106-119) is
honoured: scale[x] multiplies the likelihood term, scale[z] the prior term of latent z.

Extension: `initialize(batch_size=B)` cycles through contiguous mini-batches of the bound data on the device
(iteration t uses batch t mod floor(N/B)) instead of feeding each batch from the host.
"""
from __future__ import annotations

import numpy as np

from .. import graph as _g
from .hmc import HMC


def _scalar(v, what):
  if isinstance(v, _g.Tensor):
    v = _g.evaluate(v)
  a = np.unique(np.asarray(v, np.float64))
  if a.size != 1:
    raise NotImplementedError("scale for %s must be a scalar on this path" % what)
  return float(a[0])


class _SGMCMC(HMC):
  """Shares HMC's model recognition, device binding and Empirical re-homing (HMC.build_update)."""
  _kind = None

  def _common_init(self, step_size, args, kwargs):
    self.step_size = step_size
    self._batch_size = kwargs.pop('batch_size', 0) or 0
    self.n_steps = 0
    self._device = kwargs.pop('device', None)
    from .. import _C
    self._plan = _C.PLAN_STEPWISE
    self._row_sharded = kwargs.pop('row_sharded', None)
    from .monte_carlo import MonteCarlo
    return MonteCarlo.initialize(self, *args, **kwargs)

  def build_update(self):
    train = super(_SGMCMC, self).build_update()
    model = self._model
    if model.dtype != "float32":
      raise NotImplementedError("SGLD / SGHMC run in float32 on this path (got %s latents)" % model.dtype)
    self._lik_factor = _scalar(self.scale.get(model.y_rv, 1.0), "the observed variable")
    pf = np.ones(model.spec.n_params, np.float32)
    for slot in model.slots:
      pf[slot.offset:slot.offset + slot.size] = _scalar(self.scale.get(slot.z, 1.0), slot.z.name)
    self._prior_factor = pf
    import torch
    self._velocity = torch.zeros(model.spec.n_params, dtype=torch.float32, device=self._sampler.dev)
    return train

  def _step(self, t, n):
    self._sampler.sgmcmc_run(self._kind, self._packed[:self.n_iter], t, n, self.step_size, getattr(self, "friction", 0.0),
                             self._lik_factor, self._prior_factor, self._velocity, None, self._batch_size)

  def _train(self, feed_dict=None):
    self._maybe_rebind(feed_dict or {})
    self._step(self._t, 1)

  def _run_loop(self):
    t = self._t
    while t < self.n_iter:
      if self.n_print == 0:
        nxt = self.n_iter
      elif t == 0:
        nxt = 1
      else:
        nxt = min(self.n_iter, (t // self.n_print + 1) * self.n_print)
      self._maybe_rebind({})
      self._step(t, nxt - t)
      with np.errstate(divide='ignore', invalid='ignore'):
        accept_rate = np.float64(self._get_n_accept()) / np.float64(nxt - 1) if self.n_print != 0 else None
      t = nxt
      self._t = t
      if self.n_print != 0:
        self._log_scalars(t)
        self.print_progress({'t': t, 'accept_rate': accept_rate})
    if self.n_print == 0:
      self._sampler.read_state()

  def state_dict(self):
    """HMC.state_dict plus what the stochastic-gradient samplers carry between iterations: the SGHMC velocity, the
    friction and the mini-batch size (the batch of iteration t is a function of t, so nothing else is needed)."""
    state = super(_SGMCMC, self).state_dict()
    state["velocity"] = self._velocity.detach().cpu().numpy().copy()
    state["friction"] = float(getattr(self, "friction", 0.0))
    state["batch_size"] = int(self._batch_size)
    return state

  def load_state_dict(self, state):
    import torch
    super(_SGMCMC, self).load_state_dict(state)
    if "velocity" in state:
      self._velocity.copy_(torch.as_tensor(np.asarray(state["velocity"], np.float32)).to(self._velocity.device))
    if "friction" in state and hasattr(self, "friction"):
      self.friction = float(state["friction"])
    if "batch_size" in state:
      self._batch_size = int(state["batch_size"])


class SGLD(_SGMCMC):
  """Stochastic gradient Langevin dynamics (sgld.py:13-87)."""
  _kind = "sgld"

  def initialize(self, step_size=0.25, *args, **kwargs):
    """sgld.py:43-50. step_size: constant scale factor of the learning rate step_size / (t+1)^0.55."""
    return self._common_init(step_size, args, kwargs)


class SGHMC(_SGMCMC):
  """Stochastic gradient Hamiltonian Monte Carlo (sghmc.py:13-96)."""
  _kind = "sghmc"

  def initialize(self, step_size=0.25, friction=0.1, *args, **kwargs):
    """sghmc.py:45-57."""
    self.friction = friction
    return self._common_init(step_size, args, kwargs)
