"""The CUDA path against the committed golden fixtures (tests/golden/*.npz, written by make_golden.py from the
float64 oracle): log joint, gradient, every proposal, every accept decision and the final Empirical store,
through the C ABI with the fixtures' injected momentum / uniform draws. Both execution plans."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("ref_"))


def _sampler(d, plan):
  from edward_b200 import engine
  X = d["X"]
  spec = engine.GLMSpec(X.shape[1], bool(d["has_bias"]), int(d["family"]), d["prior_loc"], d["prior_scale"], float(d["lik_scale"]))
  return engine.GLMSampler(spec, X, d["y"], plan=plan)


@pytest.mark.parametrize("plan", [1, 2])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_reproduces_golden(path, plan):
  import torch
  d = np.load(path)
  s = _sampler(d, plan)
  P = d["theta"].shape[0]
  for name in ("theta", "zero"):
    th = d["theta"] if name == "theta" else np.zeros(P, np.float32)
    lp, g = s.logp_grad(th)
    want_lp, want_g = float(d["logp_%s_f64" % name]), d["grad_%s_f64" % name]
    assert abs(float(lp[0]) - want_lp) <= 1e-5 * abs(want_lp), (name, float(lp[0]), want_lp)
    assert np.max(np.abs(g.cpu().numpy() - want_g)) <= 1e-5 * np.max(np.abs(want_g)), name
  T, L, eps = int(d["T"]), int(d["L"]), float(d["eps"])
  params = torch.zeros(T, P, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=torch.tensor(d["r0"]), u=torch.tensor(d["u"]))
  sc, pos, got = sc.cpu().numpy(), pos.cpu().numpy(), params.cpu().numpy()
  tr = d["trace_f64"]
  margin = np.abs(tr[:, 5] - tr[:, 4])
  for i in range(T):
    same = (sc[i, 6] > 0.5) == (tr[i, 6] > 0.5)
    if not same:
      assert margin[i] < 1e-3, (i, sc[i], tr[i])  # only a genuine near-tie may flip
      break
    assert np.max(np.abs(pos[i] - d["proposals_f64"][i])) <= 1e-4 * max(np.max(np.abs(d["proposals_f64"][i])), 1e-6), i
    assert abs(sc[i, 4] - tr[i, 4]) <= 1e-5 * max(abs(tr[i, 1]), 1.0) + 1e-4, (i, sc[i, 4], tr[i, 4])
    assert np.max(np.abs(got[i] - d["params_f64"][i])) <= 1e-4 * max(np.max(np.abs(d["params_f64"][i])), 1e-6), i
  else:
    assert s.read_state()[0] == int(d["n_accept_f64"])
  s.close()


def test_cuda_leapfrog_trajectory_matches_golden():
  """z after every leapfrog step of the first transition (hmc.py:200-208): run L' = 1..L steps from the same
  start with the same momentum and compare the proposal with the stored trajectory."""
  import torch
  d = np.load([p for p in GOLDEN if p.endswith("logit_1024x54_big_step.npz")][0])
  s = _sampler(d, 1)
  P = d["theta"].shape[0]
  L, eps = int(d["L"]), float(d["eps"])
  for steps in range(1, L + 1):
    params = torch.zeros(1, P, device="cuda")
    sc, pos = s.set_trace(1)
    s.run(params, 0, 1, eps, steps, r0=torch.tensor(d["r0"][:1]), u=torch.tensor(d["u"][:1]))
    want = d["leapfrog_z_f64"][steps - 1]
    assert np.max(np.abs(pos.cpu().numpy()[0] - want)) <= 1e-4 * np.max(np.abs(want)), steps
  s.close()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_matches_reference_execution(path):
  """The CUDA path against tests/golden/ref_<case>.npz: the outputs of the reference's OWN `leapfrog` and
  `HMC.build_update` source (hmc.py:61-130,195-210) executed with torch standing in for TensorFlow
  (tests/golden/ref_exec.py, autodiff gradients). North-star tolerances: log joint and gradient 1e-5 relative, stored
  positions 1e-4 relative, accept decisions identical."""
  import torch
  d = np.load(path)
  r = np.load(os.path.join(os.path.dirname(path), "ref_" + os.path.basename(path)))
  s = _sampler(d, 0)
  P = d["theta"].shape[0]
  for name in ("theta", "zero"):
    th = d["theta"] if name == "theta" else np.zeros(P, np.float32)
    lp, g = s.logp_grad(th)
    for tag, tol in (("f64", 1e-5), ("f32", 2e-5)):  # vs the float32 reference path: both sides carry float32 round-off
      want_lp, want_g = float(r["logp_%s_%s" % (name, tag)]), r["grad_%s_%s" % (name, tag)]
      assert abs(float(lp[0]) - want_lp) <= tol * abs(want_lp), (name, tag, float(lp[0]), want_lp)
      assert np.max(np.abs(g.cpu().numpy() - want_g)) <= tol * np.max(np.abs(want_g)), (name, tag)
  T, L, eps = int(d["T"]), int(d["L"]), float(d["eps"])
  params = torch.zeros(T, P, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=torch.tensor(d["r0"]), u=torch.tensor(d["u"]))
  sc, got = sc.cpu().numpy(), params.cpu().numpy()
  margin = np.abs(d["trace_f64"][:, 5] - d["trace_f64"][:, 4])
  if np.all(margin >= 1e-3):  # no near-tie in this fixture: every decision must be the reference's
    np.testing.assert_array_equal((sc[:, 6] > 0.5).astype(np.int32), r["accepts_f64"])
    assert s.read_state()[0] == int(r["n_accept_f64"])
    want = r["params_f64"]
    assert np.max(np.abs(got - want)) <= 1e-4 * max(np.max(np.abs(want)), 1e-6)
  s.close()
