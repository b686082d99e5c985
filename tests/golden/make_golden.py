"""Generates the golden vectors under tests/golden/ from the CPU oracle (oracle/hmc_oracle.py).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference cannot be executed here (no TensorFlow), and its own tests hold no golden
values for this path, so these fixtures pin the ORACLE's outputs (a regression anchor for the oracle, the
C port and the CUDA path), not TensorFlow's. Each .npz stores the inputs as well, so no RNG stream needs to
be reproducible across numpy versions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import hmc_oracle as o  # noqa: E402

CASES = {
    # name: (N, D, has_bias, family, T, L, eps, prior_scale, lik_scale)
    "cfg1_example": (40, 1, True, o.BERNOULLI_LOGIT, 40, 2, 0.6, 3.0, 1.0),
    "logit_1024x54": (1024, 54, False, o.BERNOULLI_LOGIT, 12, 10, 0.5 / 1024, 1.0, 1.0),
    "logit_1024x54_big_step": (1024, 54, False, o.BERNOULLI_LOGIT, 16, 5, 0.03, 1.0, 1.0),
    "logit_256x1000": (256, 1000, False, o.BERNOULLI_LOGIT, 6, 4, 0.01, 1.0, 1.0),
    "normal_200x10_bias": (200, 10, True, o.NORMAL_IDENTITY, 12, 3, 0.01, 1.0, 0.5),
    "poisson_300x6_bias": (300, 6, True, o.POISSON_LOG, 10, 3, 0.01, 1.0, 1.0),
}


def make_case(name, N, D, has_bias, family, T, L, eps, prior_scale, lik_scale):
  if name == "cfg1_example":
    X, y = o.toy_dataset_cfg1(N)
    X = X.astype(np.float32)
    y = y.astype(np.int32)
  else:
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    X = rng.standard_normal((N, D)).astype(np.float32)
    wt = (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32)
    eta = X.astype(np.float64) @ wt
    if family == o.BERNOULLI_LOGIT:
      y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.int32)
    elif family == o.NORMAL_IDENTITY:
      y = (eta + lik_scale * rng.standard_normal(N)).astype(np.float32)
    else:
      y = rng.poisson(np.exp(np.clip(eta, -3, 3))).astype(np.int32)
  P = D + int(has_bias)
  spec = o.GLMSpec(D, has_bias, family, np.zeros(P, np.float32), np.full(P, prior_scale, np.float32), lik_scale)
  r0, u = o.synth_draws(T, P, seed=1234)
  theta = (0.2 * np.random.default_rng(7).standard_normal(P) / np.sqrt(D)).astype(np.float32)
  out = dict(X=X, y=y, has_bias=has_bias, family=family, T=T, L=L, eps=eps, prior_loc=spec.prior_loc,
             prior_scale=spec.prior_scale, lik_scale=lik_scale, r0=r0, u=u, theta=theta)
  for tag, dt in (("f64", np.float64), ("f32", np.float32)):
    out["logp_theta_" + tag] = o.log_joint(X, y, theta, spec, dt)
    out["grad_theta_" + tag] = o.grad_log_joint(X, y, theta, spec, dt)
    out["logp_zero_" + tag] = o.log_joint(X, y, np.zeros(P, np.float32), spec, dt)
    out["grad_zero_" + tag] = o.grad_log_joint(X, y, np.zeros(P, np.float32), spec, dt)
    params = np.zeros((T, P), dt)
    infos, nacc = o.run(X, y, params, r0, u, eps, L, spec, dt)
    out["params_" + tag] = params
    out["n_accept_" + tag] = nacc
    out["trace_" + tag] = np.array([[i.logp_old, i.logp_new, i.k_old, i.k_new, i.ratio, i.log_u, float(i.accept)]
                                    for i in infos])
    out["proposals_" + tag] = np.array([i.proposal for i in infos])
  # leapfrog trajectory of the first transition, float64 (z and r after every step, hmc.py:200-208)
  tr = []
  o.leapfrog(X, y, np.zeros(P), r0[0], eps, L, spec, np.float64, trace=tr)
  out["leapfrog_z_f64"] = np.array([z for z, _ in tr])
  out["leapfrog_r_f64"] = np.array([r for _, r in tr])
  return out


if __name__ == "__main__":
  for name, cfg in CASES.items():
    d = make_case(name, *cfg)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(name, os.path.getsize(path) // 1024, "KiB", "n_accept", d["n_accept_f64"], d["n_accept_f32"])
