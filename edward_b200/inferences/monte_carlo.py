"""MonteCarlo base class — edward/inferences/monte_carlo.py:16-170."""
from __future__ import annotations

import abc

import numpy as np

from .. import graph as _g
from ..models import Empirical
from .inference import Counter, Inference


class MonteCarlo(Inference):
  def __init__(self, latent_vars=None, data=None):
    """monte_carlo.py:61-93: a list of latents gets Empirical posteriors of 10,000 zero rows; a dict must
    map to Empirical random variables of scalar sample shape."""
    if isinstance(latent_vars, list):
      latent_vars = {z: Empirical(params=_g.Variable(
          np.zeros([int(1e4)] + list(z.batch_shape) + list(z.event_shape), z.dtype.np))) for z in latent_vars}
    elif isinstance(latent_vars, dict):
      for qz in latent_vars.values():
        if not isinstance(qz, Empirical):
          raise TypeError("Posterior approximation must consist of only Empirical random variables.")
        elif len(qz.sample_shape) != 0:
          raise ValueError("Empirical posterior approximations must have a scalar sample shape.")
    super(MonteCarlo, self).__init__(latent_vars, data)

  def initialize(self, *args, **kwargs):
    """monte_carlo.py:95-109."""
    kwargs['n_iter'] = int(np.amin([qz.params.shape.as_list()[0] for qz in self.latent_vars.values()]))
    super(MonteCarlo, self).initialize(*args, **kwargs)
    self.n_accept = Counter(self._get_n_accept, self._reset_n_accept, "n_accept")
    self.n_accept_over_t = _g.Lazy(lambda: np.float64(self._get_n_accept()) / np.float64(self._t), (), _g.float64)
    self.train = self.build_update()
    self.reset.append(_g.variables_initializer([self.n_accept]))

  def _get_n_accept(self):
    return 0

  def _reset_n_accept(self):
    pass

  def update(self, feed_dict=None):
    """monte_carlo.py:111-150: one transition, then `t += 1`. Returns {'t', 'accept_rate'} where
    accept_rate = n_accept / t with t read BEFORE the increment, as the reference's two sess.run calls
    evaluate it (monte_carlo.py:139-140; inf/nan at the first call, as there)."""
    if feed_dict is None:
      feed_dict = {}
    for key, value in self.data.items():
      if isinstance(key, _g.Tensor) and "Placeholder" in key.op_type:
        feed_dict.setdefault(key, value)
    self.train(feed_dict)
    with np.errstate(divide='ignore', invalid='ignore'):
      accept_rate = np.float64(self._get_n_accept()) / np.float64(self._t)
    self._t += 1
    t = self._t
    self._log_scalars(t)
    return {'t': t, 'accept_rate': accept_rate}

  def _summary_scalars(self):
    """tf.summary.scalar("n_accept", self.n_accept) — monte_carlo.py:106-109."""
    return {"n_accept": float(self._get_n_accept())}

  def print_progress(self, info_dict):
    """monte_carlo.py:152-158."""
    if self.n_print != 0:
      t = info_dict['t']
      if t == 1 or t % self.n_print == 0:
        self.progbar.update(t, {'Acceptance Rate': info_dict['accept_rate']})

  @abc.abstractmethod
  def build_update(self):
    raise NotImplementedError()
