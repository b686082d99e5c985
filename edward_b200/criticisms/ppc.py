"""ed.ppc — posterior predictive checks (edward/criticisms/ppc.py:13-120): T(xrep, zrep) against T(x, zrep)."""
from __future__ import annotations

import numpy as np

from .. import graph as _g
from ..models.random_variable import RandomVariable
from ..util.copying import copy
from ..util.random_variables import check_data, check_latent_vars


def ppc(T, data, latent_vars=None, n_samples=100):
  """ppc.py:13-120. `T(xs, zs)` receives dicts keyed like `data` / `latent_vars` (numpy values). For each of
  n_samples replications: draw z from the posteriors in latent_vars, draw the observed variables given z, and
  evaluate T on the replicated and on the observed data. Returns [T(xrep, zrep) array, T(x, zrep) array]."""
  if not callable(T):
    raise TypeError("T must be a callable function.")
  check_data(data)
  if latent_vars is None:
    latent_vars = {}
  check_latent_vars(latent_vars)
  if not isinstance(n_samples, int):
    raise TypeError("n_samples must have type int.")
  feed = {k: v for k, v in data.items() if isinstance(k, _g.Tensor) and "Placeholder" in k.op_type}
  observed = [k for k in data.keys() if isinstance(k, RandomVariable)]
  Trep, Tobs = [], []
  for _ in range(n_samples):
    zrep = {z: np.asarray(_g.evaluate(qz.sample())) for z, qz in latent_vars.items()}
    swap = {z: _g.constant(v, z.dtype) for z, v in zrep.items()}
    xrep = {}
    for x in observed:
      xc = copy(x, swap)
      xrep[x] = np.asarray(_g.evaluate(xc.sample(), feed))
    xobs = {x: np.asarray(data[x]) for x in observed}
    Trep.append(T(xrep, zrep))
    Tobs.append(T(xobs, zrep))
  return [np.stack(Trep), np.stack(Tobs)]
