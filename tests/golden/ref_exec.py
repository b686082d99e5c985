"""Executes the reference's OWN Python for the HMC transition with torch standing in for TensorFlow.

TEST INFRASTRUCTURE (checker side). Used by tests/golden/make_reference_golden.py (fixture generation) and by
tests/test_reference_exec.py (CPU, skipped where /root/reference does not exist — e.g. on the GPU box).

What runs from the reference tree, unmodified: the source text of `leapfrog` (edward/inferences/hmc.py:195-210) and of
`HMC.build_update` (hmc.py:61-130), extracted with `ast` from /root/reference at call time and compiled here. Nothing
is copied into this repository. What does NOT come from the reference tree, because it lives in TensorFlow (absent
here): the `tf.*` ops those functions call and the densities. They are provided by
  * `TorchTF`   — the handful of eager `tf` symbols used by the two functions (gather, maximum, random_normal /
                  random_uniform with INJECTED draws, gradients = torch.autograd.grad, reduce_sum, square, log, cond,
                  scatter_update, where, group, convert_to_tensor);
  * `tf_log_joint` — HMC._log_joint for the GLM models (hmc.py:161-192: `0.0`, `+= reduce_sum(z.log_prob(z_sample))`
                  per latent in latent_vars order, `+= reduce_sum(x.log_prob(data))`), with the [TF 1.5] density
                  expressions written op by op in torch: Normal `-0.5*square((x-loc)/scale) - (0.5*log(2*pi) +
                  log(scale))`, Bernoulli-logit `-(where(l>=0,l,0) - l*y + log1p(exp(where(l>=0,-l,l))))`, Poisson
                  `y*log_rate - exp(log_rate) - lgamma(y+1)`; `ed.dot` = matmul(X, w[:,None]) reshaped
                  (util/tensorflow.py:38-45). Gradients are AUTODIFF (the stand-in for tf.gradients, hmc.py:199,206),
                  not hand-derived formulas.
"""
from __future__ import annotations

import ast
import math
import os
from collections import OrderedDict

import numpy as np
import six
import torch

REF_HMC = "/root/reference/edward/inferences/hmc.py"


def reference_available() -> bool:
  return os.path.exists(REF_HMC)


class TorchTF(object):
  """Eager stand-in for the `tf` symbols used by hmc.py:61-130 and :195-210."""

  def __init__(self):
    self.normal_queue = []   # injected momentum draws, consumed in call order (hmc.py:88-91)
    self.uniform_queue = []  # injected accept uniforms (hmc.py:108)

  # --- state access -------------------------------------------------------------------------
  @staticmethod
  def maximum(a, b):
    return max(int(a), int(b))

  @staticmethod
  def gather(params, index):
    return params[int(index)].detach().clone().requires_grad_(True)

  @staticmethod
  def scatter_update(variable, index, value):
    with torch.no_grad():
      variable[int(index)] = value.detach().to(variable.dtype).reshape(variable[int(index)].shape)
    return variable

  # --- draws (injected so that every implementation sees the same numbers) -------------------
  def random_normal(self, shape, dtype=None):
    return self.normal_queue.pop(0)

  def random_uniform(self, shape, dtype=None):
    return self.uniform_queue.pop(0)

  # --- math ----------------------------------------------------------------------------------
  @staticmethod
  def gradients(y, xs):
    return list(torch.autograd.grad(y, xs))

  @staticmethod
  def convert_to_tensor(x):
    return x

  @staticmethod
  def reduce_sum(x):
    if isinstance(x, (list, tuple)):
      out = x[0]
      for v in x[1:]:
        out = out + v
      return out
    return torch.sum(x)

  @staticmethod
  def square(x):
    return x * x

  @staticmethod
  def log(x):
    return torch.log(x)

  @staticmethod
  def cond(pred, true_fn, false_fn):
    out = true_fn() if bool(pred) else false_fn()
    return out[0] if isinstance(out, list) and len(out) == 1 else out  # the reference handles this quirk (hmc.py:112-114)

  @staticmethod
  def where(cond, a, b):
    return a if bool(cond) else b

  @staticmethod
  def group(*ops):
    return ops


def _extract(names):
  """Source of the named top-level functions / HMC methods of the reference's hmc.py, as compiled code objects."""
  import warnings
  with open(REF_HMC) as f:
    src = f.read()
  with warnings.catch_warnings():
    warnings.simplefilter("ignore", SyntaxWarning)  # the reference's docstrings hold LaTeX escapes
    tree = ast.parse(src)
  found = {}
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in names:
      found[node.name] = node
    if isinstance(node, ast.ClassDef) and node.name == "HMC":
      for sub in node.body:
        if isinstance(sub, ast.FunctionDef) and sub.name in names:
          found[sub.name] = sub
  missing = set(names) - set(found)
  if missing:
    raise RuntimeError("reference hmc.py has no %s" % sorted(missing))
  mod = ast.Module(body=[found[n] for n in names], type_ignores=[])
  return compile(ast.fix_missing_locations(mod), REF_HMC, "exec")


def load_reference_functions(tf):
  ns = {"tf": tf, "six": six, "OrderedDict": OrderedDict}
  exec(_extract(["leapfrog", "build_update"]), ns)
  return ns["leapfrog"], ns["build_update"]


# ------------------------------------------------------------------------------------------------
# [TF 1.5] densities, op by op in torch (the part of the path that lives in TensorFlow)
# ------------------------------------------------------------------------------------------------
def tf_normal_log_prob(x, loc, scale):
  z = (x - loc) / scale
  return -0.5 * (z * z) - (0.5 * math.log(2.0 * math.pi) + torch.log(scale))


def tf_bernoulli_logits_log_prob(logits, y):
  zeros = torch.zeros_like(logits)
  cond = logits >= zeros
  relu_logits = torch.where(cond, logits, zeros)
  neg_abs_logits = torch.where(cond, -logits, logits)
  return -((relu_logits - logits * y) + torch.log1p(torch.exp(neg_abs_logits)))


def tf_poisson_log_rate_log_prob(log_rate, y):
  return y * log_rate - torch.exp(log_rate) - torch.lgamma(y + 1.0)


class Qz(object):
  """What hmc.py needs of an Empirical posterior: params (the tf.Variable), event_shape, dtype, get_variables()."""

  def __init__(self, params):
    self.params = params
    self.event_shape = tuple(params.shape[1:])
    self.dtype = params.dtype

  def get_variables(self):
    return [self.params]


class Counter(object):
  def __init__(self):
    self.value = 0

  def assign_add(self, v):
    self.value += int(v)
    return self.value


class FakeHMC(object):
  """The attributes HMC.build_update reads from `self` (hmc.py:61-130)."""

  def __init__(self, latent_order, qz, log_joint, step_size, n_steps):
    self.latent_vars_unconstrained = OrderedDict((k, qz[k]) for k in latent_order)
    self._log_joint_unconstrained = log_joint  # identity transformation for real-valued latents (hmc.py:132-159)
    self.step_size = step_size
    self.n_steps = n_steps
    self.t = 0
    self.n_accept = Counter()


def run_reference(X, y, has_bias, family, prior_loc, prior_scale, lik_scale, r0, u, step_size, n_steps, T, dtype,
                  z0=None):
  """T transitions of the reference's build_update on a GLM. Returns dict(params [T,P], n_accept, ratio/accept trace is
  recomputed by the caller from params if needed). Latent order: w, then b (the order the example passes them)."""
  td = torch.float64 if dtype == np.float64 else torch.float32
  Xt = torch.as_tensor(np.asarray(X), dtype=td)
  yt = torch.as_tensor(np.asarray(y).astype(np.float64), dtype=td)  # cast(event, float) inside log_prob
  D = Xt.shape[1]
  loc = torch.as_tensor(np.asarray(prior_loc), dtype=td)
  sc = torch.as_tensor(np.asarray(prior_scale), dtype=td)
  lik_s = torch.tensor(float(lik_scale), dtype=td)

  def log_joint(z_sample):  # hmc.py:161-192 for this model
    lj = 0.0
    for key in order:  # `for z in six.iterkeys(self.latent_vars)` (:183-185)
      if key == "w":
        lj = lj + torch.sum(tf_normal_log_prob(z_sample["w"], loc[:D], sc[:D]))
      else:
        lj = lj + torch.sum(tf_normal_log_prob(z_sample["b"], loc[D], sc[D]))
    eta = torch.matmul(Xt, z_sample["w"][:, None]).reshape(-1)  # ed.dot, util/tensorflow.py:38-45
    if has_bias:
      eta = eta + z_sample["b"]
    if family == 0:
      ll = tf_bernoulli_logits_log_prob(eta, yt)
    elif family == 1:
      ll = tf_normal_log_prob(yt, eta, lik_s)
    else:
      ll = tf_poisson_log_rate_log_prob(eta, yt)
    return lj + torch.sum(ll)  # :187-190

  order = ["w", "b"] if has_bias else ["w"]
  P = D + int(has_bias)
  z0 = np.zeros(P) if z0 is None else np.asarray(z0)
  qz = {"w": Qz(torch.zeros(T, D, dtype=td))}
  qz["w"].params[0] = torch.as_tensor(z0[:D], dtype=td)
  if has_bias:
    qz["b"] = Qz(torch.zeros(T, dtype=td))
    qz["b"].params[0] = float(z0[D])
  tf = TorchTF()
  leapfrog, build_update = load_reference_functions(tf)
  hmc = FakeHMC(order, qz, log_joint, step_size, n_steps)
  accepts = []
  for t in range(T):
    hmc.t = t
    tf.normal_queue = [torch.as_tensor(r0[t][:D], dtype=td)]
    if has_bias:
      tf.normal_queue.append(torch.tensor(float(r0[t][D]), dtype=td))
    tf.uniform_queue = [torch.tensor(float(u[t]), dtype=td)]
    before = hmc.n_accept.value
    build_update(hmc)
    accepts.append(hmc.n_accept.value - before)
  params = np.zeros((T, P), np.float64)
  params[:, :D] = qz["w"].params.detach().numpy()
  if has_bias:
    params[:, D] = qz["b"].params.detach().numpy()
  return {"params": params, "n_accept": hmc.n_accept.value, "accepts": np.array(accepts, np.int32)}


def reference_logp_grad(X, y, has_bias, family, prior_loc, prior_scale, lik_scale, theta, dtype):
  """log joint and its AUTODIFF gradient at theta (the quantity tf.gradients(log_joint(z), z) yields, hmc.py:199)."""
  td = torch.float64 if dtype == np.float64 else torch.float32
  Xt = torch.as_tensor(np.asarray(X), dtype=td)
  yt = torch.as_tensor(np.asarray(y).astype(np.float64), dtype=td)
  D = Xt.shape[1]
  loc = torch.as_tensor(np.asarray(prior_loc), dtype=td)
  sc = torch.as_tensor(np.asarray(prior_scale), dtype=td)
  th = torch.as_tensor(np.asarray(theta), dtype=td).clone().requires_grad_(True)
  w = th[:D]
  lj = 0.0 + torch.sum(tf_normal_log_prob(w, loc[:D], sc[:D]))
  eta = torch.matmul(Xt, w[:, None]).reshape(-1)
  if has_bias:
    lj = lj + torch.sum(tf_normal_log_prob(th[D], loc[D], sc[D]))
    eta = eta + th[D]
  if family == 0:
    ll = tf_bernoulli_logits_log_prob(eta, yt)
  elif family == 1:
    ll = tf_normal_log_prob(yt, eta, torch.tensor(float(lik_scale), dtype=td))
  else:
    ll = tf_poisson_log_rate_log_prob(eta, yt)
  lj = lj + torch.sum(ll)
  (g,) = torch.autograd.grad(lj, th)
  return float(lj.detach()), g.detach().numpy().astype(np.float64)
