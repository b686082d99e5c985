"""Generates tests/golden/ref_<case>.npz by EXECUTING the reference's own `leapfrog` and `HMC.build_update`
(/root/reference/edward/inferences/hmc.py, read at generation time, see ref_exec.py) on the inputs stored in the
existing fixtures <case>.npz, with torch standing in for TensorFlow (eager ops, autodiff gradients, [TF 1.5] density
expressions). Runs only where /root/reference exists (this container):

    python tests/golden/make_reference_golden.py

The outputs are the pin for oracle/hmc_oracle.py (tests/test_reference_exec.py) and for the CUDA path
(tests/test_gpu_golden.py): stored Empirical rows, accept decisions, log joint and autodiff gradient, float64 and float32.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_exec as rx  # noqa: E402


def make(path):
  d = np.load(path)
  X, y = d["X"], d["y"]
  has_bias, family = bool(d["has_bias"]), int(d["family"])
  T, L, eps = int(d["T"]), int(d["L"]), float(d["eps"])
  out = {}
  for tag, dt in (("f64", np.float64), ("f32", np.float32)):
    r = rx.run_reference(X, y, has_bias, family, d["prior_loc"], d["prior_scale"], float(d["lik_scale"]), d["r0"], d["u"],
                         eps, L, T, dt)
    out["params_" + tag] = r["params"]
    out["n_accept_" + tag] = r["n_accept"]
    out["accepts_" + tag] = r["accepts"]
    for name, th in (("theta", d["theta"]), ("zero", np.zeros_like(d["theta"]))):
      lp, g = rx.reference_logp_grad(X, y, has_bias, family, d["prior_loc"], d["prior_scale"], float(d["lik_scale"]), th, dt)
      out["logp_%s_%s" % (name, tag)] = lp
      out["grad_%s_%s" % (name, tag)] = g
  return out


if __name__ == "__main__":
  if not rx.reference_available():
    raise SystemExit("/root/reference is not present: the reference fixtures can only be regenerated where it is")
  for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
    name = os.path.basename(path)
    if name.startswith("ref_"):
      continue
    out = make(path)
    dst = os.path.join(HERE, "ref_" + name)
    np.savez_compressed(dst, **out)
    print(name, "->", os.path.basename(dst), os.path.getsize(dst) // 1024, "KiB  n_accept", out["n_accept_f64"], out["n_accept_f32"])
