"""ed.dot (edward/util/tensorflow.py:10-45)."""
from __future__ import annotations

from .. import graph as _g


def dot(x, y):
  """Dot product between a 2-D tensor and a 1-D tensor: [M x N]·[N] → [M] or [M]·[M x N] → [N].

  Builds a `Dot` node. `ed.HMC` recognises it as the linear predictor of a GLM and carries out its
  finite check (util/tensorflow.py:33-36) once, when the sampler binds the data, instead of on every
  evaluation. Evaluating the node eagerly (`.eval()`) checks both operands and raises on NaN/Inf like
  the reference (tests/util/dot_test.py:22-31)."""
  return _g.Dot(x, y)
