"""CPU oracle for Edward's HMC hot path (GLM likelihoods, Normal priors).

TEST INFRASTRUCTURE ONLY. Nothing under edward_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.

PARITY PIN. The reference (/root/reference, blei-lab/edward 1.3.5) delegates every arithmetic op on this path to
TensorFlow 1.x (`tensorflow>=1.2.0rc0`, setup.py:18; CI pins tensorflow==1.5.0, .travis.yml:40), which is neither
vendored nor installable here, and the reference's own tests hold no golden vector / known-answer value for this path
(tests/inferences/hmc_test.py:33-35,79-80 assert posterior moments only). TensorFlow output is therefore NOT available;
the pin is the next best thing: tests/golden/make_reference_golden.py + ref_exec.py EXECUTE the reference's own
`leapfrog` and `HMC.build_update` source (edward/inferences/hmc.py, read from /root/reference at generation time) with
torch standing in for the few tf.* calls they make and torch.autograd for tf.gradients on the literal TF density
expressions, and the resulting fixtures tests/golden/ref_*.npz (committed with the generator) are what this file is
held to (tests/test_reference_exec.py), float32 and float64. This file restates
  * the algorithm from the reference's own sources (cited per function), and
  * the TF 1.5 densities from their published definitions (cited as [TF 1.5]),
and is additionally pinned by: scipy.stats densities, finite differences, the in-tree closed forms
edward/inferences/conjugacy/conjugate_log_probs.py:21-24,134-141, and the reference's statistical
HMC tests (see tests/test_oracle.py).

Two arithmetic modes:
  dtype=np.float32 — follows the TF op order of the float32 reference path (separate mul/add, no FMA);
  dtype=np.float64 — the same formulas in double: the "truth" the float32 paths are judged against.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

BERNOULLI_LOGIT = 0
NORMAL_IDENTITY = 1
POISSON_LOG = 2


@dataclass
class GLMSpec:
  """Model of the hot path: theta=[w(D), b?] ~ Normal(prior_loc, prior_scale); y ~ family(X w + b)."""
  n_features: int
  has_bias: bool = False
  family: int = BERNOULLI_LOGIT
  prior_loc: Optional[np.ndarray] = None    # [P]
  prior_scale: Optional[np.ndarray] = None  # [P]
  lik_scale: float = 1.0                    # Normal-identity only
  prior_kind: Optional[np.ndarray] = None   # [P] 0 Normal(loc, scale); 1 Beta(a=loc, b=scale) latent sampled through a sigmoid

  def __post_init__(self):
    P = self.n_params
    if self.prior_loc is None:
      self.prior_loc = np.zeros(P, np.float32)
    if self.prior_scale is None:
      self.prior_scale = np.ones(P, np.float32)
    self.prior_loc = np.asarray(self.prior_loc, np.float32).reshape(P)
    self.prior_scale = np.asarray(self.prior_scale, np.float32).reshape(P)
    if self.prior_kind is not None:
      self.prior_kind = np.asarray(self.prior_kind, np.int32).reshape(P)

  @property
  def n_params(self) -> int:
    return self.n_features + (1 if self.has_bias else 0)


# --------------------------------------------------------------------------------------------
# ed.dot — edward/util/tensorflow.py:10-45
# --------------------------------------------------------------------------------------------
def dot(x, y):
  """`ed.dot`: finite-check both operands (util/tensorflow.py:33-36, raises on NaN/Inf as in
  tests/util/dot_test.py:22-31), then matmul + reshape(-1) (util/tensorflow.py:38-45)."""
  x = np.asarray(x)
  y = np.asarray(y)
  if not np.all(np.isfinite(x)) or not np.all(np.isfinite(y)):
    raise ValueError("InvalidArgumentError: Tensor had NaN or Inf values")
  if x.ndim == 1:
    return np.matmul(x[None, :], y).reshape(-1)
  return np.matmul(x, y[:, None]).reshape(-1)


# --------------------------------------------------------------------------------------------
# Densities [TF 1.5, not in the reference tree]
# --------------------------------------------------------------------------------------------
def normal_log_prob(x, loc, scale, dtype):
  """[TF 1.5] distributions/normal.py::_log_prob =
  -0.5*square((x-loc)/scale) - (0.5*log(2*pi) + log(scale)). Reached from hmc.py:185."""
  x = np.asarray(x, dtype)
  loc = np.asarray(loc, dtype)
  scale = np.asarray(scale, dtype)
  z = (x - loc) / scale
  log_unnorm = dtype(-0.5) * np.square(z)
  log_norm = dtype(0.5 * math.log(2.0 * math.pi)) + np.log(scale)
  return log_unnorm - log_norm


def normal_log_prob_grad(x, loc, scale, dtype):
  """Autodiff of normal_log_prob wrt x as tf.gradients produces it: -0.5*(2*z) / scale."""
  x = np.asarray(x, dtype)
  loc = np.asarray(loc, dtype)
  scale = np.asarray(scale, dtype)
  z = (x - loc) / scale
  return (dtype(-0.5) * (dtype(2.0) * z)) / scale


def beta_logit_log_prob(u, a, b, dtype):
  """Density, in the unconstrained space u = logit(z), of a latent z ~ Beta(a, b) moved there by `ed.transform`
  (util/random_variables.py:895-897: TransformedDistribution(z, Invert(Sigmoid()))), as HMC._log_joint_unconstrained
  evaluates it (hmc.py:132-159): [TF 1.5] Beta._log_prob(z) = (a-1) log z + (b-1) log1p(-z) - lbeta(a, b) at z = sigmoid(u)
  plus the bijector's inverse_log_det_jacobian = log sigmoid(u) + log sigmoid(-u). Collected:
  a log sigmoid(u) + b log sigmoid(-u) - lbeta(a, b)."""
  from scipy.special import betaln
  u = np.asarray(u, dtype)
  a = np.asarray(a, dtype)
  b = np.asarray(b, dtype)
  ls_pos = np.minimum(u, 0) - np.log1p(np.exp(-np.abs(u)))
  ls_neg = np.minimum(-u, 0) - np.log1p(np.exp(-np.abs(u)))
  return (a * ls_pos + b * ls_neg - betaln(a, b)).astype(dtype)


def beta_logit_log_prob_grad(u, a, b, dtype):
  u = np.asarray(u, dtype)
  a = np.asarray(a, dtype)
  b = np.asarray(b, dtype)
  return (a - (a + b) / (1 + np.exp(-u))).astype(dtype)


def prior_log_prob(theta, spec, dtype):
  """Sum over the latents of the prior log density in the unconstrained space (hmc.py:183-185 + :132-159)."""
  theta = np.asarray(theta, dtype)
  if spec.prior_kind is None:
    return normal_log_prob(theta, spec.prior_loc, spec.prior_scale, dtype)
  out = np.empty(spec.n_params, dtype)
  for i in range(spec.n_params):
    if spec.prior_kind[i] == 1:
      out[i] = beta_logit_log_prob(theta[i], spec.prior_loc[i], spec.prior_scale[i], dtype)
    else:
      out[i] = normal_log_prob(theta[i], spec.prior_loc[i], spec.prior_scale[i], dtype)
  return out


def prior_log_prob_grad(theta, spec, dtype):
  theta = np.asarray(theta, dtype)
  if spec.prior_kind is None:
    return normal_log_prob_grad(theta, spec.prior_loc, spec.prior_scale, dtype)
  out = np.empty(spec.n_params, dtype)
  for i in range(spec.n_params):
    if spec.prior_kind[i] == 1:
      out[i] = beta_logit_log_prob_grad(theta[i], spec.prior_loc[i], spec.prior_scale[i], dtype)
    else:
      out[i] = normal_log_prob_grad(theta[i], spec.prior_loc[i], spec.prior_scale[i], dtype)
  return out


def bernoulli_logit_log_prob(logits, y, dtype):
  """[TF 1.5] distributions/bernoulli.py::_log_prob = -nn.sigmoid_cross_entropy_with_logits(
  labels=cast(event, float), logits) with nn_impl.py's stable form
  add(where(l>=0,l,0) - l*y, log1p(exp(where(l>=0,-l,l)))). Reached from hmc.py:190."""
  l = np.asarray(logits, dtype)
  yv = np.asarray(y).astype(dtype)
  cond = l >= 0
  relu = np.where(cond, l, dtype(0))
  neg_abs = np.where(cond, -l, l)
  return -((relu - l * yv) + np.log1p(np.exp(neg_abs)))


def bernoulli_logit_log_prob_grad(logits, y, dtype):
  """d/dlogits of bernoulli_logit_log_prob, piecewise through the `where`s as autodiff does:
  -( [l>=0] - y + sign * e/(1+e) ), e = exp(-|l|), sign = -1 if l>=0 else +1.  Equals y - sigmoid(l)."""
  l = np.asarray(logits, dtype)
  yv = np.asarray(y).astype(dtype)
  cond = l >= 0
  e = np.exp(np.where(cond, -l, l))
  q = (dtype(1.0) / (dtype(1.0) + e)) * e
  return np.where(cond, (yv - dtype(1.0)) + q, yv - q)


def poisson_log_log_prob(eta, y, dtype):
  """[TF 1.5] distributions/poisson.py::_log_prob with log_rate=eta: y*eta - exp(eta) - lgamma(y+1)."""
  from scipy.special import gammaln
  eta = np.asarray(eta, dtype)
  yv = np.asarray(y).astype(dtype)
  return (yv * eta - np.exp(eta) - gammaln(yv + dtype(1.0)).astype(dtype)).astype(dtype)


def poisson_log_log_prob_grad(eta, y, dtype):
  eta = np.asarray(eta, dtype)
  yv = np.asarray(y).astype(dtype)
  return yv - np.exp(eta)


# --------------------------------------------------------------------------------------------
# HMC._log_joint — edward/inferences/hmc.py:161-192
# --------------------------------------------------------------------------------------------
def _split(theta, spec: GLMSpec):
  D = spec.n_features
  w = theta[:D]
  b = theta[D] if spec.has_bias else None
  return w, b


def linear_predictor(X, theta, spec: GLMSpec, dtype, check=True):
  """`ed.dot(X, w) + b` as in examples/bayesian_logistic_regression.py:44."""
  w, b = _split(np.asarray(theta, dtype), spec)
  Xd = np.asarray(X, dtype)
  eta = dot(Xd, w) if check else np.matmul(Xd, w[:, None]).reshape(-1)
  if b is not None:
    eta = eta + b
  return eta


def log_lik_terms(eta, y, spec: GLMSpec, dtype):
  if spec.family == BERNOULLI_LOGIT:
    return bernoulli_logit_log_prob(eta, y, dtype)
  if spec.family == NORMAL_IDENTITY:
    return normal_log_prob(np.asarray(y).astype(dtype), eta, dtype(spec.lik_scale), dtype)
  if spec.family == POISSON_LOG:
    return poisson_log_log_prob(eta, y, dtype)
  raise ValueError("unknown family")


def log_lik_grad_eta(eta, y, spec: GLMSpec, dtype):
  if spec.family == BERNOULLI_LOGIT:
    return bernoulli_logit_log_prob_grad(eta, y, dtype)
  if spec.family == NORMAL_IDENTITY:
    s = dtype(spec.lik_scale)
    yv = np.asarray(y).astype(dtype)
    # d/dloc of -0.5*((y-loc)/s)^2 = ((y-loc)/s)/s
    return ((yv - eta) / s) / s
  if spec.family == POISSON_LOG:
    return poisson_log_log_prob_grad(eta, y, dtype)
  raise ValueError("unknown family")


def log_joint(X, y, theta, spec: GLMSpec, dtype=np.float64):
  """HMC._log_joint (hmc.py:161-192): `log_joint = 0.0`, then `+= reduce_sum(z.log_prob(z_sample))`
  for each latent in latent_vars order (:183-185; w before b), then `+= reduce_sum(x.log_prob(data))`
  for each observed RV (:187-190). `self.scale` is ignored by HMC (it never appears in :182-190).
  _log_joint_unconstrained (:132-159) is the identity for Normal latents (real support)."""
  theta = np.asarray(theta, dtype)
  w, b = _split(theta, spec)
  D = spec.n_features
  lj = dtype(0.0)
  pr = prior_log_prob(theta, spec, dtype)
  lj = lj + np.sum(pr[:D], dtype=dtype)
  if spec.has_bias:
    lj = lj + np.sum(pr[D:], dtype=dtype)
  eta = linear_predictor(X, theta, spec, dtype)
  lj = lj + np.sum(log_lik_terms(eta, y, spec, dtype), dtype=dtype)
  return dtype(lj)


def grad_log_joint(X, y, theta, spec: GLMSpec, dtype=np.float64):
  """`tf.gradients(log_joint(z), z)` of hmc.py:199,206: prior term + X^T (dlogp/deta) via the backward
  MatMul of util/tensorflow.py:45 (+ reduce_sum for the bias)."""
  theta = np.asarray(theta, dtype)
  D = spec.n_features
  eta = linear_predictor(X, theta, spec, dtype)
  r = log_lik_grad_eta(eta, y, spec, dtype)
  Xd = np.asarray(X, dtype)
  g = np.empty(spec.n_params, dtype)
  g[:D] = np.matmul(Xd.T, r[:, None]).reshape(-1)
  if spec.has_bias:
    g[D] = np.sum(r, dtype=dtype)
  g = g + prior_log_prob_grad(theta, spec, dtype)
  return g.astype(dtype)


# --------------------------------------------------------------------------------------------
# leapfrog — edward/inferences/hmc.py:195-210
# --------------------------------------------------------------------------------------------
def leapfrog(X, y, z_old, r_old, step_size, n_steps, spec: GLMSpec, dtype=np.float64, trace=None):
  """hmc.py:195-210. Gradient once before the loop (:199), then per step: r += 0.5*eps*g; z += eps*r
  (:201-204), g = grad(z) (:206), r += 0.5*eps*g (:207-208). `0.5 * step_size` is a Python float
  product folded into one constant before it meets the tensor, as in the reference expression."""
  z = np.array(z_old, dtype)
  r = np.array(r_old, dtype)
  half_eps = dtype(0.5 * step_size)
  eps = dtype(step_size)
  g = grad_log_joint(X, y, z, spec, dtype)
  for _ in range(n_steps):
    r = r + half_eps * g
    z = z + eps * r
    g = grad_log_joint(X, y, z, spec, dtype)
    r = r + half_eps * g
    if trace is not None:
      trace.append((z.copy(), r.copy()))
  return z, r


# --------------------------------------------------------------------------------------------
# HMC.build_update — edward/inferences/hmc.py:61-130
# --------------------------------------------------------------------------------------------
@dataclass
class TransitionInfo:
  logp_old: float
  logp_new: float
  k_old: float
  k_new: float
  ratio: float
  log_u: float
  accept: bool
  proposal: np.ndarray = field(repr=False, default=None)

  @property
  def margin(self) -> float:
    """|log u - ratio|: how far the accept decision is from a tie."""
    return abs(self.log_u - self.ratio)


def transition(X, y, params, t, r0, u, step_size, n_steps, spec: GLMSpec, dtype=np.float64):
  """One HMC transition, hmc.py:81-130, with injected momentum r0 (:88-91) and uniform u (:108).
  Reads params[max(t-1,0)] (:81-85), leapfrogs (:94-97), ratio = K(r0) - K(rL) + logp(zL) - logp(z0)
  accumulated in that order (:100-105), accept = log(u) < ratio, strict (:108-109), writes params[t]
  (:121-126). Returns TransitionInfo; params is updated in place."""
  old = np.array(params[max(t - 1, 0)], dtype).reshape(-1)
  r0 = np.asarray(r0, dtype).reshape(-1)
  new, r_new = leapfrog(X, y, old, r0, step_size, n_steps, spec, dtype)
  k_old = dtype(0.5) * np.sum(np.square(r0), dtype=dtype)
  k_new = dtype(0.5) * np.sum(np.square(r_new), dtype=dtype)
  logp_new = log_joint(X, y, new, spec, dtype)
  logp_old = log_joint(X, y, old, spec, dtype)
  ratio = dtype(k_old)
  ratio = dtype(ratio - k_new)
  ratio = dtype(ratio + logp_new)
  ratio = dtype(ratio - logp_old)
  log_u = np.log(dtype(u))
  accept = bool(log_u < ratio)
  sample = new if accept else old
  params[t] = sample.astype(params.dtype)
  return TransitionInfo(float(logp_old), float(logp_new), float(k_old), float(k_new), float(ratio),
                        float(log_u), accept, new.copy())


def run(X, y, params, r0_all, u_all, step_size, n_steps, spec: GLMSpec, dtype=np.float64, t0=0,
        n_iter=None):
  """Inference.run's loop (inference.py:145-147) over MonteCarlo.update (monte_carlo.py:111-150):
  n_iter = number of Empirical rows (monte_carlo.py:96-97). Returns (infos, n_accept)."""
  T = params.shape[0]
  if n_iter is None:
    n_iter = T - t0
  if t0 + n_iter > T:
    raise IndexError("scatter_update index out of range")  # hmc.py:125
  infos = []
  n_accept = 0
  for i in range(n_iter):
    info = transition(X, y, params, t0 + i, r0_all[i], u_all[i], step_size, n_steps, spec, dtype)
    n_accept += int(info.accept)
    infos.append(info)
  return infos, n_accept


# --------------------------------------------------------------------------------------------
# Empirical — edward/models/empirical.py:87-110
# --------------------------------------------------------------------------------------------
def empirical_mean(params):
  """empirical.py:87-88: reduce_mean(params, 0)."""
  return np.mean(params, axis=0)


def empirical_stddev(params):
  """empirical.py:90-93: sqrt(reduce_mean(square(params - mean), 0)) — population std."""
  r = params - np.mean(params, axis=0)
  return np.sqrt(np.mean(np.square(r), axis=0))


# --------------------------------------------------------------------------------------------
# Synthetic inputs fixed by SURVEY.md §8(d)
# --------------------------------------------------------------------------------------------
GEN_BLOCK_ROWS = 65536


def synth_block(block_index, n_rows, D, w_true, base_seed=42):
  """Rows [block_index*65536, +n_rows) of the synthetic design: X ~ N(0,1) fp32, y ~ Bernoulli(sigmoid(X w_true)),
  seeded per block so that any row sharding on block boundaries yields identical data."""
  rng = np.random.Generator(np.random.Philox(key=base_seed + block_index))
  X = rng.standard_normal((n_rows, D), dtype=np.float32)
  p = 1.0 / (1.0 + np.exp(-(X.astype(np.float64) @ w_true.astype(np.float64))))
  y = (rng.random(n_rows) < p).astype(np.int32)
  return X, y


def synth_w_true(D, base_seed=42):
  rng = np.random.Generator(np.random.Philox(key=base_seed + (1 << 40)))
  return (rng.standard_normal(D) / math.sqrt(D)).astype(np.float32)


def synth_data(N, D, base_seed=42, row_start=0):
  """Rows [row_start, row_start+N); row_start must be a multiple of GEN_BLOCK_ROWS."""
  assert row_start % GEN_BLOCK_ROWS == 0
  w_true = synth_w_true(D, base_seed)
  Xs, ys = [], []
  done = 0
  b = row_start // GEN_BLOCK_ROWS
  while done < N:
    n = min(GEN_BLOCK_ROWS, N - done)
    Xb, yb = synth_block(b, n, D, w_true, base_seed)
    Xs.append(Xb)
    ys.append(yb)
    done += n
    b += 1
  return np.concatenate(Xs), np.concatenate(ys), w_true


def synth_draws(n_iter, P, seed=1234):
  """Momentum r0[n_iter,P] ~ N(0,1) and accept uniforms u[n_iter] in (0,1), injected into both the
  oracle and the device for parity runs."""
  rng = np.random.Generator(np.random.Philox(key=seed))
  r0 = rng.standard_normal((n_iter, P), dtype=np.float32)
  u = rng.random(n_iter, dtype=np.float32)
  u = np.clip(u, np.float32(1e-7), np.float32(1.0 - 1e-7)).astype(np.float32)
  return r0, u


def toy_dataset_cfg1(N=40, noise_std=0.1):
  """examples/bayesian_logistic_regression.py:23-31 under ed.set_seed(42) (:35 → np.random.seed,
  util/graphs.py:72)."""
  np.random.seed(42)
  X = np.linspace(-6, 6, num=N)
  y = np.tanh(X) + np.random.normal(0, noise_std, size=N)
  y[y < 0.5] = 0
  y[y >= 0.5] = 1
  X = (X - 4.0) / 4.0
  X = X.reshape((N, 1))
  return X, y


# --------------------------------------------------------------------------------------------
# SGLD / SGHMC — edward/inferences/sgld.py:52-121, sghmc.py:58-130
# --------------------------------------------------------------------------------------------
def scaled_grad_log_joint(X, y, theta, spec: GLMSpec, lik_factor=1.0, prior_factor=None, dtype=np.float64):
  """Gradient of SGLD._log_joint (sgld.py:89-121): every log_prob term is multiplied by `scale.get(rv, 1.0)`
  before it is summed, so the likelihood gradient carries lik_factor and each prior gradient its own factor."""
  theta = np.asarray(theta, dtype)
  D = spec.n_features
  eta = linear_predictor(X, theta, spec, dtype)
  r = log_lik_grad_eta(eta, y, spec, dtype)
  g = np.empty(spec.n_params, dtype)
  g[:D] = np.matmul(np.asarray(X, dtype).T, r[:, None]).reshape(-1)
  if spec.has_bias:
    g[D] = np.sum(r, dtype=dtype)
  pf = np.ones(spec.n_params, dtype) if prior_factor is None else np.asarray(prior_factor, dtype)
  return (dtype(lik_factor) * g + pf * normal_log_prob_grad(theta, spec.prior_loc, spec.prior_scale, dtype)).astype(dtype)


def sgld_run(X, y, params, noise, step_size, spec: GLMSpec, dtype=np.float64, t0=0, n_iter=None, lik_factor=1.0,
             prior_factor=None, batch_rows=0):
  """SGLD.build_update (sgld.py:52-87): lr = step_size / (t+1)^0.55; sample = old + 0.5*lr*grad + sqrt(lr)*normal,
  old = params[max(t-1,0)], written to params[t]. batch_rows > 0: mini-batch (t mod floor(N/B)) of the rows."""
  T = params.shape[0]
  n_iter = T - t0 if n_iter is None else n_iter
  N = X.shape[0]
  for i in range(n_iter):
    t = t0 + i
    old = np.array(params[max(t - 1, 0)], dtype)
    if batch_rows:
      lo = (t % (N // batch_rows)) * batch_rows
      Xb, yb = X[lo:lo + batch_rows], y[lo:lo + batch_rows]
    else:
      Xb, yb = X, y
    lr = dtype(step_size) / np.power(dtype(t + 1), dtype(0.55))
    g = scaled_grad_log_joint(Xb, yb, old, spec, lik_factor, prior_factor, dtype)
    params[t] = old + dtype(0.5) * lr * g + np.sqrt(lr) * np.asarray(noise[i], dtype)
  return params


def sghmc_run(X, y, params, noise, step_size, friction, spec: GLMSpec, dtype=np.float64, t0=0, n_iter=None,
              lik_factor=1.0, prior_factor=None, v0=None, batch_rows=0):
  """SGHMC.build_update (sghmc.py:58-96): lr = 0.01*step_size; sample = old + v; v = (1-0.5*friction)*v + lr*grad(old)
  + sqrt(lr*friction)*normal. Returns the final velocity."""
  T = params.shape[0]
  n_iter = T - t0 if n_iter is None else n_iter
  v = np.zeros(spec.n_params, dtype) if v0 is None else np.array(v0, dtype)
  N = X.shape[0]
  lr = dtype(step_size) * dtype(0.01)
  sd = np.sqrt(lr * dtype(friction))
  for i in range(n_iter):
    t = t0 + i
    old = np.array(params[max(t - 1, 0)], dtype)
    if batch_rows:
      lo = (t % (N // batch_rows)) * batch_rows
      Xb, yb = X[lo:lo + batch_rows], y[lo:lo + batch_rows]
    else:
      Xb, yb = X, y
    g = scaled_grad_log_joint(Xb, yb, old, spec, lik_factor, prior_factor, dtype)
    params[t] = old + v
    v = (dtype(1.0) - dtype(0.5) * dtype(friction)) * v + lr * g + sd * np.asarray(noise[i], dtype)
  return v


# ---------------------------------------------------------------------------------------------------------------------
# Posterior-predictive criticism: what edward/criticisms/evaluate.py computes for a GLM output under
# S posterior draws. evaluate.py:132-143 averages the Bernoulli probabilities over the draws; :158-162 averages draws of
# a continuous output (here: their conditional means); :222-227 ('log_lik') averages log p(y | draw) over rows and draws.
# ---------------------------------------------------------------------------------------------------------------------
def predictive(X, y, W, B, family, lik_scale=1.0):
  """X [N, D], W [S, D], B [S] or None -> (mean_s E[y_n | eta_ns] [N], sum_s log p(y_n | eta_ns) [N]), float64."""
  X = np.asarray(X, np.float64)
  W = np.asarray(W, np.float64)
  y = np.asarray(y, np.float64)
  eta = X @ W.T
  if B is not None:
    eta = eta + np.asarray(B, np.float64)[None, :]
  yc = y[:, None]
  if family == BERNOULLI_LOGIT:
    mean = (1.0 / (1.0 + np.exp(-eta))).mean(axis=1)
    ll = -(np.maximum(eta, 0.0) - eta * yc + np.log1p(np.exp(-np.abs(eta))))  # tf sigmoid_cross_entropy_with_logits
  elif family == NORMAL_IDENTITY:
    mean = eta.mean(axis=1)
    ll = -0.5 * ((yc - eta) / lik_scale) ** 2 - (0.5 * np.log(2.0 * np.pi) + np.log(lik_scale))
  else:
    from scipy.special import gammaln
    mean = np.exp(eta).mean(axis=1)
    ll = yc * eta - np.exp(eta) - gammaln(yc + 1.0)
  return mean, ll.sum(axis=1)
