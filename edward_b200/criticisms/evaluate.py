"""ed.evaluate (edward/criticisms/evaluate.py:20-235) for posterior predictives of the GLMs of this path, computed
on the device over the Empirical sample stores that ed.HMC / ed.SGLD left there.

The reference evaluates the output variable n_samples times, each run drawing a fresh posterior sample of every
latent (independently per latent, empirical.py:98-110), and averages: probabilities for Bernoulli outputs
(evaluate.py:132-143), draws for continuous outputs (:158-162), mean log-density for 'log_lik' (:222-227).
Here the n_samples draws are gathered from the device stores and pushed through ONE fused kernel (edhmc_predictive,
csrc/criticism.cuh): the [N, D]·[D, S] contraction with the family's link and the reductions over the draws in its epilogue.
"""
from __future__ import annotations

import numpy as np

from .. import graph as _g
from ..glm import _decompose
from ..models import Bernoulli, Empirical, Normal, Poisson
from ..models.random_variable import RandomVariable
from ..util.random_variables import _is_placeholder, check_data


def _device():
  import torch
  if not torch.cuda.is_available():
    raise RuntimeError("ed.evaluate runs on the GPU (no CPU fallback)")
  return torch.device("cuda", torch.cuda.current_device())


def _latent_draws(z, n_samples, dev):
  """[S, size] draws of latent z on the device: rows of an Empirical store, or prior draws of a Normal."""
  import torch
  if isinstance(z, Empirical):
    t = z._device_params()
    if t is None:
      t = torch.as_tensor(np.asarray(_g.evaluate(z.params), np.float32), device=dev)
    idx = torch.as_tensor(np.random.randint(0, z.n, size=n_samples), device=t.device)
    return t[idx].reshape(n_samples, -1).to(dev, torch.float32)
  if isinstance(z, Normal):
    loc = np.broadcast_to(_g.evaluate(z.loc), tuple(z.shape) or ()).reshape(-1)
    scale = np.broadcast_to(_g.evaluate(z.scale), tuple(z.shape) or ()).reshape(-1)
    draws = loc + scale * np.random.standard_normal((n_samples, loc.size))
    return torch.as_tensor(draws.astype(np.float32), device=dev)
  raise NotImplementedError("latent %s in the posterior predictive" % type(z).__name__)


def predictive(X, y, W, B, family, lik_scale=1.0, want_loglik=True):
  """One fused device pass (edhmc_predictive, csrc/criticism.cuh) over rows X [N, D] and S posterior draws W [S, D],
  B [S] or None: returns (mean_s E[y_n | eta_ns] as float32 [N], sum_s log p(y_n | eta_ns) as float64 [N] or None).
  eta [N, S] is never materialised."""
  import ctypes as C
  import torch
  from .. import _C
  from ..engine import _stream_ptr, _Y_DTYPES
  dev = X.device
  X = X.to(torch.float32)
  if X.stride(1) != 1:
    X = X.contiguous()
  S, D = int(W.shape[0]), int(W.shape[1])
  cols = [W.to(dev, torch.float32)]
  if B is not None:
    cols.append(B.to(dev, torch.float32).reshape(S, 1))
  packed = torch.cat(cols, dim=1).contiguous()  # [S, D (+1)]: the draws, gathered from the sample stores
  idx = torch.arange(S, dtype=torch.int32, device=dev)
  yt = None
  if y is not None:
    yt = y.to(dev)
    if yt.dtype not in _Y_DTYPES:
      yt = yt.to(torch.float32)
    yt = yt.contiguous()
  N = int(X.shape[0])
  mean = torch.empty(N, dtype=torch.float32, device=dev)
  ll = torch.empty(N, dtype=torch.float64, device=dev) if (want_loglik and yt is not None) else None
  with torch.cuda.device(dev):
    _C.check(_C.lib().edhmc_predictive(
        X.data_ptr(), N, int(X.stride(0)) if N > 1 else max(int(X.stride(0)), D), D,
        yt.data_ptr() if yt is not None else None, _Y_DTYPES[yt.dtype] if yt is not None else 0, int(family), float(lik_scale),
        packed.data_ptr(), int(packed.stride(0)), idx.data_ptr(), idx.data_ptr() if B is not None else None, D, S,
        mean.data_ptr(), ll.data_ptr() if ll is not None else None, dev.index, _stream_ptr(dev)))
  return mean, ll


def _predictive_inputs(output_key, data, n_samples, dev):
  """(X [N, D] on the device, W [S, D], B [S] | None, family, lik_scale) of the output variable under n_samples draws."""
  import torch
  from .. import _C
  lik_scale = 1.0
  if isinstance(output_key, Bernoulli):
    eta_node, family = output_key.logits, _C.BERNOULLI_LOGIT
  elif isinstance(output_key, Normal):
    eta_node, family = output_key.loc, _C.NORMAL_IDENTITY
    lik_scale = float(np.unique(_g.evaluate(output_key.scale))[0])
  elif isinstance(output_key, Poisson):
    eta_node, family = output_key.log_rate, _C.POISSON_LOG
  else:
    raise NotImplementedError("ed.evaluate: output %s" % type(output_key).__name__)
  if eta_node is None:
    raise NotImplementedError("ed.evaluate: the output must be parameterised by its linear predictor")
  latents = []

  def collect(n):
    if isinstance(n, RandomVariable):
      latents.append(n)
    elif isinstance(n, _g._Binary):
      collect(n.a)
      collect(n.b)
    elif isinstance(n, _g.Dot):
      collect(n.x)
      collect(n.y)
  collect(eta_node)
  x_node, w, b = _decompose(eta_node, latents)
  N = int(output_key.shape[0]) if len(output_key.shape) else 1
  if x_node is None:
    X = torch.ones(N, 1, device=dev)
  else:
    xv = data[x_node] if x_node in data else _g.evaluate(x_node)
    X = xv.to(dev, torch.float32) if isinstance(xv, torch.Tensor) else torch.as_tensor(np.asarray(xv, np.float32), device=dev)
  W = _latent_draws(w, n_samples, dev)          # [S, D]
  B = _latent_draws(b, n_samples, dev).reshape(n_samples) if b is not None else None
  return X, W, B, family, lik_scale


def evaluate(metrics, data, n_samples=500, output_key=None, seed=None):
  """evaluate.py:20-235, metrics: 'binary_accuracy', 'log_loss'/'binary_crossentropy', 'mse', 'mae',
  'log_lik'/'log_likelihood', 'accuracy' (binary), or callables f(y_true, y_pred) on numpy arrays."""
  import torch
  if isinstance(metrics, str) or callable(metrics):
    metrics = [metrics]
  elif not isinstance(metrics, list):
    raise TypeError("metrics must have type str or list, or be callable.")
  check_data(data)
  if not isinstance(n_samples, int):
    raise TypeError("n_samples must have type int.")
  if output_key is None:
    keys = [k for k in data.keys() if not _is_placeholder(k)]
    if len(keys) == 1:
      output_key = keys[0]
    else:
      raise KeyError("User must specify output_key.")
  elif not isinstance(output_key, RandomVariable):
    raise TypeError("output_key must have type RandomVariable.")
  if seed is not None:
    np.random.seed(seed)
  dev = _device()
  yv = data[output_key]
  y_true = yv.to(dev, torch.float32) if isinstance(yv, torch.Tensor) else torch.as_tensor(np.asarray(yv, np.float32), device=dev)
  X, W, B, family, lik_scale = _predictive_inputs(output_key, data, n_samples, dev)
  want_ll = any((m[0] if isinstance(m, tuple) else m) in ('log_lik', 'log_likelihood') for m in metrics)
  mean, ll = predictive(X, y_true, W, B, family, lik_scale, want_loglik=want_ll)

  probs = y_pred = None
  if isinstance(output_key, Bernoulli):
    probs = mean  # evaluate.py:132-143: the mean of the Bernoulli probabilities over the draws
    rnd = torch.rand_like(probs)
    y_pred = torch.round(torch.where(probs == 0.5, rnd, probs))
  elif isinstance(output_key, Normal):
    # evaluate.py:158-162 averages n_samples draws y ~ Normal(eta_s, scale): their mean is mean_s(eta_s) plus
    # Normal(0, scale^2 / n_samples) noise
    y_pred = mean + (lik_scale / np.sqrt(n_samples)) * torch.randn_like(mean)
  elif isinstance(output_key, Poisson):
    # the sum of independent Poisson(rate_s) draws is Poisson(sum_s rate_s)
    y_pred = torch.poisson(mean * n_samples) / n_samples

  out = []
  for metric in metrics:
    if isinstance(metric, tuple):
      metric = metric[0]
    if metric == 'accuracy':
      metric = 'binary_accuracy'
    if metric == 'binary_accuracy':
      out.append(float((y_true == y_pred).float().mean()))
    elif metric in ('log_loss', 'binary_crossentropy'):
      logit_pred = torch.log(y_pred.clamp(1e-8, 1 - 1e-8)) - torch.log1p(-y_pred.clamp(1e-8, 1 - 1e-8))
      ce = torch.clamp(logit_pred, min=0) - logit_pred * y_true + torch.log1p(torch.exp(-logit_pred.abs()))
      out.append(float(ce.mean()))
    elif metric in ('mse', 'MSE', 'mean_squared_error'):
      out.append(float(((y_pred - y_true) ** 2).mean()))
    elif metric in ('mae', 'MAE', 'mean_absolute_error'):
      out.append(float((y_pred - y_true).abs().mean()))
    elif metric in ('log_lik', 'log_likelihood'):
      out.append(float(ll.sum() / (ll.numel() * n_samples)))  # evaluate.py:222-227: mean over rows and draws
    elif callable(metric):
      out.append(metric(y_true.cpu().numpy(), y_pred.cpu().numpy()))
    else:
      raise NotImplementedError("Metric is not implemented: {}".format(metric))
  return out[0] if len(out) == 1 else out
