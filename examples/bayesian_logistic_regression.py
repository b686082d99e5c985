#!/usr/bin/env python
"""Bayesian logistic regression with HMC on the B200 path (BASELINE config 1).

The model, the toy data and the sampler settings are those of the reference's example of the same name
(one feature + intercept, Normal(0, 3) priors, 5,000 Empirical samples, step_size 0.6, n_steps 2); the script
itself is written for this package: argparse options, an update()/print_progress() loop like a user would write
against Edward, and a text summary of the posterior (add --png FILE for the posterior-predictive curves).

    python examples/bayesian_logistic_regression.py [--T 5000] [--N 40] [--png fit.png]
"""
import argparse

import numpy as np

import edward_b200 as ed
from edward_b200 import tfshim as tf
from edward_b200.models import Bernoulli, Empirical, Normal


def toy_problem(n_points, noise=0.1):
  """Noisy tanh on [-6, 6] thresholded at 0.5, inputs rescaled to roughly [-2.5, 0.5]."""
  grid = np.linspace(-6, 6, num=n_points)
  target = np.tanh(grid) + np.random.normal(0, noise, size=n_points)
  labels = np.where(target < 0.5, 0.0, 1.0)
  features = ((grid - 4.0) / 4.0).reshape((n_points, 1))
  return features, labels


def build_model(n_points, n_features, n_samples):
  """Returns the placeholder, the latent variables, the likelihood and the Empirical posteriors."""
  x = tf.placeholder(tf.float32, [n_points, n_features])
  weights = Normal(loc=tf.zeros(n_features), scale=3.0 * tf.ones(n_features))
  intercept = Normal(loc=tf.zeros([]), scale=3.0 * tf.ones([]))
  labels = Bernoulli(logits=ed.dot(x, weights) + intercept)
  q_weights = Empirical(params=tf.get_variable("qw/params", [n_samples, n_features]))
  q_intercept = Empirical(params=tf.get_variable("qb/params", [n_samples]))
  return x, weights, intercept, labels, q_weights, q_intercept


def posterior_predictive(q_weights, q_intercept, inputs, n_curves):
  """sigmoid(inputs·w + b) for n_curves joint draws from the Empirical posteriors (a lazy tensor: evaluate it
  whenever a snapshot of the current samples is wanted)."""
  curves = [tf.sigmoid(ed.dot(inputs, q_weights.sample()) + q_intercept.sample()) for _ in range(n_curves)]
  return tf.stack(curves)


def save_png(path, x_train, y_train, inputs, curves):
  import matplotlib
  matplotlib.use("Agg")
  import matplotlib.pyplot as plt
  fig, ax = plt.subplots(figsize=(6, 6))
  ax.plot(x_train[:, 0], y_train, "bx")
  for row in curves:
    ax.plot(inputs[:, 0], row, alpha=0.2)
  ax.set_xlim(-5, 3)
  ax.set_ylim(-0.5, 1.5)
  fig.savefig(path)


def main():
  ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
  ap.add_argument("--N", type=int, default=40, help="data points")
  ap.add_argument("--T", type=int, default=5000, help="posterior samples")
  ap.add_argument("--png", default=None, help="write the posterior-predictive curves to this file")
  args = ap.parse_args()

  ed.set_seed(42)
  x_train, y_train = toy_problem(args.N)
  x, weights, intercept, labels, q_weights, q_intercept = build_model(args.N, 1, args.T)

  inference = ed.HMC({weights: q_weights, intercept: q_intercept}, data={x: x_train, labels: y_train})
  inference.initialize(n_print=10, step_size=0.6)
  tf.global_variables_initializer().run()

  inputs = np.linspace(-5, 3, num=400, dtype=np.float32).reshape((400, 1))
  curves = posterior_predictive(q_weights, q_intercept, inputs, n_curves=50)

  info = {}
  while inference.t.eval() < inference.n_iter:
    info = inference.update()
    inference.print_progress(info)
  inference.finalize()

  snapshot = curves.eval()
  print("posterior mean of w: %s, b: %s, accept rate %.3f" % (
      q_weights.mean().eval(), q_intercept.mean().eval(), info["accept_rate"]))
  print("posterior predictive at x=-2, 0, 2: %s" % np.round(snapshot.mean(axis=0)[[150, 250, 350]], 3))
  if args.png:
    save_png(args.png, x_train, y_train, inputs, snapshot)
  return 0


if __name__ == "__main__":
  raise SystemExit(main())
