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
  s = _g.get_default_graph().seed
  return s if s is not None else _seed


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
