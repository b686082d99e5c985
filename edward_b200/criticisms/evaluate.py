"""ed.evaluate (edward/criticisms/evaluate.py:20-235) for posterior predictives of the GLMs of this path, computed
on the device over the Empirical sample stores that ed.HMC / ed.SGLD left there.

The reference evaluates the output variable n_samples times, each run drawing a fresh posterior sample of every
latent (independently per latent, empirical.py:98-110), and averages: probabilities for Bernoulli outputs
(evaluate.py:132-143), draws for continuous outputs (:158-162), mean log-density for 'log_lik' (:222-227).
Here the n_samples draws are gathered from the device stores and pushed through one [N, D]·[D, S] product.
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


def _predictive_eta(output_key, data, n_samples, dev):
  """eta [N, S] of the output variable's linear predictor under n_samples posterior draws."""
  import torch
  if isinstance(output_key, Bernoulli):
    eta_node = output_key.logits
  elif isinstance(output_key, Normal):
    eta_node = output_key.loc
  elif isinstance(output_key, Poisson):
    eta_node = output_key.log_rate
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
  eta = X @ W.t()                                # [N, S]
  if b is not None:
    eta = eta + _latent_draws(b, n_samples, dev).reshape(1, n_samples)
  return eta


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
  eta = _predictive_eta(output_key, data, n_samples, dev)
  y_col = y_true.reshape(-1, 1)

  probs = y_pred = None
  if isinstance(output_key, Bernoulli):
    probs = torch.sigmoid(eta).mean(dim=1)
    rnd = torch.rand_like(probs)
    y_pred = torch.round(torch.where(probs == 0.5, rnd, probs))
  elif isinstance(output_key, Normal):
    scale = float(np.unique(_g.evaluate(output_key.scale))[0])
    y_pred = (eta + scale * torch.randn_like(eta)).mean(dim=1)
  elif isinstance(output_key, Poisson):
    y_pred = torch.poisson(torch.exp(eta)).mean(dim=1)

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
      if isinstance(output_key, Bernoulli):
        lp = -(torch.clamp(eta, min=0) - eta * y_col + torch.log1p(torch.exp(-eta.abs())))
      elif isinstance(output_key, Normal):
        scale = float(np.unique(_g.evaluate(output_key.scale))[0])
        lp = -0.5 * ((y_col - eta) / scale) ** 2 - (0.5 * np.log(2 * np.pi) + np.log(scale))
      else:
        lp = y_col * eta - torch.exp(eta) - torch.lgamma(y_col + 1.0)
      out.append(float(lp.mean(dim=0).mean()))
    elif callable(metric):
      out.append(metric(y_true.cpu().numpy(), y_pred.cpu().numpy()))
    else:
      raise NotImplementedError("Metric is not implemented: {}".format(metric))
  return out[0] if len(out) == 1 else out
