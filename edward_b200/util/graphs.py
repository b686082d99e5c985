"""set_seed / get_session / random_variables (edward/util/graphs.py:13-73)."""
from __future__ import annotations

import numpy as np

from .. import graph as _g

_seed = None


def set_seed(x):
  """Seed NumPy and the device Philox streams (graphs.py:59-73). As in the reference, seeding after part
  of the model has been built is an error (graphs.py:66-70)."""
  global _seed
  if _g.get_default_graph().n_nodes > 0:
    raise RuntimeError("Seeding is not supported after initializing part of the graph. "
                       "Please move set_seed to the beginning of your code.")
  np.random.seed(x)
  _g.get_default_graph().seed = int(x)
  _seed = int(x)


def get_seed():
  """The seed given to ed.set_seed for the CURRENT default graph, or None (as in the reference, where a reset graph is
  unseeded again: graphs.py:59-73 seeds tf's current default graph)."""
  return _g.get_default_graph().seed


def sampler_seed(rank0_broadcast=None):
  """64-bit Philox seed for a new sampler. Seeded graph: derived from the user's seed and the number of samplers this
  graph has seeded so far, so that several inference objects of one seeded program draw different, reproducible
  streams (the first one keeps the plain seed). Unseeded (the reference's default: random every run): 64 fresh bits
  from the OS, agreed across ranks through `rank0_broadcast` when the rows are sharded (every rank must draw the same
  momenta)."""
  import os
  seed = get_seed()
  graph = _g.get_default_graph()
  k = getattr(graph, "seed_uses", 0)
  graph.seed_uses = k + 1
  if seed is not None:
    x = (int(seed) & (2 ** 64 - 1)) ^ ((k * 0x9E3779B97F4A7C15) & (2 ** 64 - 1))
    return x if k else int(seed) & (2 ** 64 - 1)  # the first sampler of a seeded program keeps the plain seed
  fresh = int.from_bytes(os.urandom(8), "little")
  if rank0_broadcast is not None:
    fresh = int(rank0_broadcast(fresh))
  return fresh


class _Session(object):
  """There is no session: the device work is queued by libedhmc on the current CUDA stream. `run`
  evaluates nodes / executes reset ops so that reference-style test code keeps working."""

  def run(self, fetches, feed_dict=None):
    if isinstance(fetches, (list, tuple)):
      return [self.run(f, feed_dict) for f in fetches]
    if hasattr(fetches, "run") and not isinstance(fetches, _g.Tensor):
      return fetches.run(feed_dict)
    if hasattr(fetches, "eval"):
      return fetches.eval(feed_dict)
    return fetches

  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False

  def close(self):
    pass


_ED_SESSION = None


def get_session():
  global _ED_SESSION
  if _ED_SESSION is None:
    _ED_SESSION = _Session()
  return _ED_SESSION


def random_variables(graph=None):
  return list((graph or _g.get_default_graph()).random_variables)
