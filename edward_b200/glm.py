"""Recognises the models of the HMC hot path in a symbolic description and turns them into a GLMSpec.

The reference conditions an arbitrary TF model on a sample by graph rewriting (`copy(z, dict_swap)`,
edward/util/random_variables.py:139-445, called from HMC._log_joint, hmc.py:172-190). This path only
needs "evaluate log p(data, z) at z" for

    latents     w ~ Normal(loc, scale)  [D]      (and optionally  b ~ Normal(loc, scale), scalar)
    likelihood  y ~ Bernoulli(logits=eta) | Normal(loc=eta, scale=s) | Poisson(log_rate=eta)
    predictor   eta = ed.dot(X, w) [+ b]         (or eta = mu for a scalar latent with sample_shape=N)

so the recogniser pattern-matches exactly that and raises NotImplementedError for anything else — there
is no generic (slow) evaluator to fall back to.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _C
from . import graph as _g
from .engine import GLMSpec
from .models import Bernoulli, Beta, Empirical, Normal, Poisson, TransformedDistribution
from .models.random_variable import RandomVariable


@dataclass
class LatentSlot:
  z: RandomVariable      # model latent
  qz: Empirical          # its Empirical posterior
  offset: int            # first column in the packed [T, P] store
  size: int              # number of columns
  scalar: bool           # event shape () → params is [T], else [T, size]


@dataclass
class GLMModel:
  spec: GLMSpec
  x_node: Optional[_g.Tensor]  # Placeholder / Variable / Constant holding X, or None for the ones-design
  y_rv: Optional[RandomVariable]  # None: no observed variable (the latents are sampled from their priors)
  slots: List[LatentSlot]
  n_rows: int
  prior_kinds: Optional[np.ndarray] = None  # per latent dimension (edhmc_set_prior_kinds), None = all Normal
  dtype: str = "float32"  # "float32" (the hot path) or "float64" (hmc_test.py:93-97, the compact device path)


def _unsupported(msg):
  raise NotImplementedError(
      "edward_b200.HMC covers GLM-style models only (Normal priors; Bernoulli-logit, Normal-identity or "
      "Poisson-log likelihood over ed.dot(X, w) [+ b]): " + msg)


def _const_value(t, what):
  """Evaluates a parameter tensor that must not depend on placeholders or random variables."""
  try:
    return np.asarray(_g.evaluate(t))
  except Exception:
    _unsupported("%s must be a constant" % what)


def _strip(node):
  return node


def _as_latent(node, latents):
  for z in latents:
    if node is z or (isinstance(node, RandomVariable) and node is z):
      return z
  return None


def _decompose(eta, latents):
  """eta → (X node | None, w latent | None, b latent | None)."""
  if isinstance(eta, _g.Add):
    for first, second in ((eta.a, eta.b), (eta.b, eta.a)):
      if isinstance(first, _g.Dot):
        xn, w, b0 = _decompose(first, latents)
        b = _as_latent(second, latents)
        if b is None or b0 is not None:
          _unsupported("the additive term of the linear predictor must be one latent Normal")
        return xn, w, b
    _unsupported("linear predictor must be ed.dot(X, w) + b")
  if isinstance(eta, _g.Dot):
    xs, ys = eta.x, eta.y
    w = _as_latent(ys, latents)
    if w is None or len(xs.shape) != 2:
      _unsupported("ed.dot must multiply a data matrix X [N, D] by a latent vector w [D]")
    if isinstance(xs, RandomVariable):
      _unsupported("X must be data, not a random variable")
    return xs, w, None
  z = _as_latent(eta, latents)
  if z is not None:
    return None, z, None
  _unsupported("cannot interpret the likelihood's parameter as a linear predictor")


def _prior_of(z):
  """(kind, p0, p1) of the prior of latent z IN THE UNCONSTRAINED SPACE the sampler works in (hmc.py:132-159):
  Normal(loc, scale) -> itself; Beta(a, b), moved to the real line through the inverse sigmoid -> kind 1 with the
  log-det-Jacobian folded in; TransformedDistribution(Normal(loc, scale), Softplus) with support 'nonnegative', moved back
  through the inverse softplus -> the base Normal again (the two Jacobians cancel)."""
  if isinstance(z, Normal):
    return _C.PRIOR_NORMAL, z.loc, z.scale
  if isinstance(z, Beta):
    return _C.PRIOR_BETA_LOGIT, z.concentration1, z.concentration0
  if isinstance(z, TransformedDistribution) and isinstance(z.distribution, Normal):
    from . import bijectors as tfb
    if isinstance(z.bijector, tfb.Softplus) and getattr(z, "support", None) == 'nonnegative':
      return _C.PRIOR_NORMAL, z.distribution.loc, z.distribution.scale
  _unsupported("latent %s must have a Normal prior, a Beta prior, or be a Softplus-transformed Normal" % z.name)


def recognize(latent_vars: dict, data: dict) -> GLMModel:
  """latent_vars maps each ORIGINAL latent to the Empirical store of its unconstrained samples."""
  latents = list(latent_vars.keys())
  observed = [k for k in data.keys() if isinstance(k, RandomVariable)]
  dtypes = {z.dtype.name for z in latents}
  if len(dtypes) != 1 or not dtypes <= {"float32", "float64"}:
    _unsupported("the latent variables must share one dtype, float32 or float64 (got %s)" % sorted(dtypes))
  dtype = dtypes.pop()
  for z in latents:
    if not isinstance(latent_vars[z], Empirical):
      raise TypeError("Posterior approximation must consist of only Empirical random variables.")
  if len(observed) == 0:
    # no data: one scalar latent sampled from its prior (tests/inferences/inference_auto_transform_test.py:114-161)
    if len(latents) != 1 or int(np.prod(tuple(latents[0].shape) or (1,))) != 1:
      _unsupported("without observed variables exactly one scalar latent is supported")
    z = latents[0]
    kind, p0, p1 = _prior_of(z)
    spec = GLMSpec(1, False, _C.BERNOULLI_LOGIT,
                   np.asarray(_const_value(p0, "prior parameter"), np.float32).reshape(1),
                   np.asarray(_const_value(p1, "prior parameter"), np.float32).reshape(1), 1.0)
    kinds = np.array([kind], np.int32)
    return GLMModel(spec, None, None, [LatentSlot(z, latent_vars[z], 0, 1, len(z.shape) == 0)], 0,
                    kinds if kind != _C.PRIOR_NORMAL else None, dtype)
  if len(observed) != 1:
    _unsupported("at most one observed random variable is supported, got %d" % len(observed))
  y_rv = observed[0]

  lik_scale = 1.0
  if isinstance(y_rv, Bernoulli):
    if y_rv.logits is not None:
      family, eta = _C.BERNOULLI_LOGIT, y_rv.logits
    else:
      # Bernoulli(probs=z) with z a (0,1)-valued latent sampled through the sigmoid: the logits ARE the unconstrained
      # latent (inference_auto_transform_test.py:163-188, Beta-Bernoulli)
      z = _as_latent(y_rv._probs, latents)
      if z is None or not isinstance(z, Beta):
        _unsupported("Bernoulli must be parameterised by logits, or by probs = a Beta latent")
      family, eta = _C.BERNOULLI_LOGIT, z
  elif isinstance(y_rv, Normal):
    family, eta = _C.NORMAL_IDENTITY, y_rv.loc
    sc = np.unique(_const_value(y_rv.scale, "the likelihood scale"))
    if sc.size != 1:
      _unsupported("the Normal likelihood needs one scale shared by all rows")
    lik_scale = float(sc[0])
  elif isinstance(y_rv, Poisson):
    if y_rv.log_rate is None:
      _unsupported("Poisson must be parameterised by log_rate")
    family, eta = _C.POISSON_LOG, y_rv.log_rate
  else:
    _unsupported("likelihood %s" % type(y_rv).__name__)

  x_node, w, b = _decompose(eta, latents)
  used = [z for z in (w, b) if z is not None]
  if len(used) != len(latents) or any(z not in used for z in latents):
    _unsupported("every latent variable must appear in the linear predictor (and nothing else may)")
  for z in used:
    if isinstance(z, Beta) and not (z is eta):
      _unsupported("a Beta latent can only be the success probability of a Bernoulli likelihood")

  n_rows = int(y_rv.shape[0]) if len(y_rv.shape) >= 1 else 1
  if x_node is None:
    # y_n ~ family(mu) for a scalar latent mu: design matrix of ones (tests/inferences/hmc_test.py:14-46)
    if int(np.prod(tuple(w.shape) or (1,))) != 1:
      _unsupported("a latent used directly as the predictor must be scalar")
    D = 1
  else:
    D = int(x_node.shape[1])
    if tuple(w.shape) != (D,):
      _unsupported("w must have shape [%d], got %s" % (D, tuple(w.shape)))
    if x_node.shape[0] is not None:
      n_rows = int(x_node.shape[0])
  if b is not None and int(np.prod(tuple(b.shape) or (1,))) != 1:
    _unsupported("the bias latent must be scalar")

  P = D + (1 if b is not None else 0)
  loc = np.zeros(P, np.float32)
  scale = np.ones(P, np.float32)
  kinds = np.zeros(P, np.int32)
  slots = []
  off = 0
  for z, size in ((w, D), (b, 1)):
    if z is None:
      continue
    kind, p0, p1 = _prior_of(z)
    loc[off:off + size] = np.broadcast_to(_const_value(p0, "prior loc"), tuple(z.shape) or ()).reshape(-1)
    scale[off:off + size] = np.broadcast_to(_const_value(p1, "prior scale"), tuple(z.shape) or ()).reshape(-1)
    kinds[off:off + size] = kind
    slots.append(LatentSlot(z, latent_vars[z], off, size, len(z.shape) == 0))
    off += size
  spec = GLMSpec(D, b is not None, family, loc, scale, lik_scale)
  return GLMModel(spec, x_node, y_rv, slots, n_rows, kinds if np.any(kinds != 0) else None, dtype)
