"""Pins the CPU oracle against the reference's OWN Python executed with torch standing in for TensorFlow.

tests/golden/ref_<case>.npz were produced by tests/golden/make_reference_golden.py: the unmodified source of `leapfrog`
(edward/inferences/hmc.py:195-210) and `HMC.build_update` (hmc.py:61-130) is read from /root/reference and executed
with eager torch ops for the `tf.*` calls, torch.autograd for `tf.gradients`, and the [TF 1.5] density expressions
written op by op (tests/golden/ref_exec.py). The oracle (oracle/hmc_oracle.py) restates the same algorithm with
hand-derived gradients; here the two must agree. CPU only; the fixtures travel, /root/reference does not."""
import glob
import os
import sys

import numpy as np
import pytest

import hmc_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_exec as rx  # noqa: E402

CASES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


def _load(path):
  d = np.load(path)
  r = np.load(os.path.join(os.path.dirname(path), "ref_" + os.path.basename(path)))
  spec = o.GLMSpec(d["X"].shape[1], bool(d["has_bias"]), int(d["family"]), d["prior_loc"], d["prior_scale"], float(d["lik_scale"]))
  return d, r, spec


def _rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_gradient_equals_reference_autodiff(path):
  """log joint and gradient: hand-derived oracle vs autodiff through the literal TF expressions (hmc.py:183-190,199)."""
  d, r, spec = _load(path)
  for name, th in (("theta", d["theta"]), ("zero", np.zeros_like(d["theta"]))):
    lp64 = float(o.log_joint(d["X"], d["y"], th, spec, np.float64))
    g64 = o.grad_log_joint(d["X"], d["y"], th, spec, np.float64)
    assert abs(lp64 - float(r["logp_%s_f64" % name])) <= 1e-12 * abs(lp64)
    assert _rel(g64, r["grad_%s_f64" % name]) <= 1e-11
    # the float32 paths (numpy op order vs torch op order) agree to float32 round-off
    lp32 = float(o.log_joint(d["X"], d["y"], th, spec, np.float32))
    g32 = o.grad_log_joint(d["X"], d["y"], th, spec, np.float32)
    assert abs(lp32 - float(r["logp_%s_f32" % name])) <= 1e-5 * abs(lp64)
    assert _rel(g32, r["grad_%s_f32" % name]) <= 1e-5
    # and the float32 reference path is within the north-star tolerance of float64
    assert _rel(r["grad_%s_f32" % name], r["grad_%s_f64" % name]) <= 1e-5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_chain_equals_reference_build_update(path):
  """T transitions: stored Empirical rows and accept decisions of the oracle vs the reference's build_update."""
  d, r, spec = _load(path)
  T, L, eps = int(d["T"]), int(d["L"]), float(d["eps"])
  params = np.zeros((T, spec.n_params), np.float64)
  infos, nacc = o.run(d["X"], d["y"], params, d["r0"], d["u"], eps, L, spec, np.float64)
  assert nacc == int(r["n_accept_f64"])
  np.testing.assert_array_equal(np.array([int(i.accept) for i in infos]), r["accepts_f64"])
  assert _rel(params, r["params_f64"]) <= 1e-10
  # the committed oracle fixtures are the same numbers
  assert _rel(d["params_f64"], r["params_f64"]) <= 1e-10
  # float32 chain: decisions identical unless a transition is a near-tie, positions to 1e-4 (north star)
  p32 = np.zeros((T, spec.n_params), np.float32)
  infos32, nacc32 = o.run(d["X"], d["y"], p32, d["r0"], d["u"], eps, L, spec, np.float32)
  ties = [i for i in infos if i.margin < 1e-3]
  if not ties:
    np.testing.assert_array_equal(np.array([int(i.accept) for i in infos32]), r["accepts_f32"])
    assert _rel(p32, r["params_f32"]) <= 1e-4


@pytest.mark.skipif(not rx.reference_available(), reason="/root/reference is not present on this machine")
def test_fixtures_are_reproducible_from_the_reference_tree():
  """Re-executes the reference source for two cases and compares with the committed fixtures (bitwise for float64
  decisions, 1e-12 for values): the fixtures are what the reference's code yields, not hand-edited numbers."""
  for case in ("cfg1_example", "poisson_300x6_bias"):
    path = os.path.join(HERE, "golden", case + ".npz")
    d, r, spec = _load(path)
    got = rx.run_reference(d["X"], d["y"], bool(d["has_bias"]), int(d["family"]), d["prior_loc"], d["prior_scale"],
                           float(d["lik_scale"]), d["r0"], d["u"], float(d["eps"]), int(d["L"]), int(d["T"]), np.float64)
    np.testing.assert_array_equal(got["accepts"], r["accepts_f64"])
    assert _rel(got["params"], r["params_f64"]) <= 1e-12


@pytest.mark.skipif(not rx.reference_available(), reason="/root/reference is not present on this machine")
def test_reference_functions_are_extracted_not_restated():
  """The executed code objects come from the reference file itself."""
  tf = rx.TorchTF()
  leapfrog, build_update = rx.load_reference_functions(tf)
  assert leapfrog.__code__.co_filename == rx.REF_HMC and build_update.__code__.co_filename == rx.REF_HMC
  assert leapfrog.__code__.co_firstlineno >= 190 and build_update.__code__.co_firstlineno >= 55
