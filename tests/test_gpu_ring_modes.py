"""GPU parity of the three data-pass layouts of the narrow-model kernel on the same inputs:
ring mode 0 (one TMA ring per warp), 1 (one TMA ring per CTA) and 2 (bind-time re-lay into column-pair-major 32-row
tiles read with LDG, csrc/stream_ldg.cuh), each against the float64 oracle, and the plan rule that picks between them.

The oracle restates /root/reference/edward/inferences/hmc.py:61-210 (see oracle/hmc_oracle.py); tolerances are the
ones of tests/test_gpu_engine_parity.py (north_star: 1e-5 log joint / gradient, 1e-4 positions).
"""
import os

import numpy as np
import pytest

import hmc_oracle as o
from test_gpu_engine_parity import REL_GRAD, REL_LOGP, REL_POS, _mk, _rel, _sampler

pytestmark = pytest.mark.gpu


class _ring:
  """EDHMC_RING is read when a handle plans its passes (edhmc_create / edhmc_bind_data)."""

  def __init__(self, mode):
    self.mode = mode

  def __enter__(self):
    self.old = os.environ.get("EDHMC_RING")
    if self.mode is None:
      os.environ.pop("EDHMC_RING", None)
    else:
      os.environ["EDHMC_RING"] = str(self.mode)

  def __exit__(self, *exc):
    if self.old is None:
      os.environ.pop("EDHMC_RING", None)
    else:
      os.environ["EDHMC_RING"] = self.old


# rows chosen so that the last 32-row tile is ragged, a CTA gets no tile at all, or a warp's last batch of tiles is short
LDG_SHAPES = [
    (1, 1, False, o.BERNOULLI_LOGIT),
    (31, 2, True, o.BERNOULLI_LOGIT),
    (33, 3, False, o.BERNOULLI_LOGIT),
    (1000, 8, False, o.BERNOULLI_LOGIT),
    (4097, 16, True, o.BERNOULLI_LOGIT),
    (70001, 20, False, o.BERNOULLI_LOGIT),
    (9999, 33, False, o.BERNOULLI_LOGIT),
    (50017, 48, True, o.BERNOULLI_LOGIT),
    (70001, 54, False, o.BERNOULLI_LOGIT),
    (20011, 63, True, o.BERNOULLI_LOGIT),
    (300000, 64, False, o.BERNOULLI_LOGIT),
    (1500, 10, True, o.NORMAL_IDENTITY),
    (2000, 12, True, o.POISSON_LOG),
]


@pytest.mark.parametrize("N,D,bias,fam", LDG_SHAPES)
def test_ldg_tiles_logp_grad_matches_oracle_and_rings(N, D, bias, fam):
  X, y, spec = _mk(N, D, bias, fam, seed=3 * N + D, lik_scale=0.7)
  rng = np.random.default_rng(5)
  theta = (0.3 * rng.standard_normal(spec.n_params) / np.sqrt(D)).astype(np.float32)
  lp64 = float(o.log_joint(X, y, theta, spec, np.float64))
  g64 = o.grad_log_joint(X, y, theta, spec, np.float64)
  got = {}
  for mode in (1, 2):
    with _ring(mode):
      s = _sampler(X, y, spec)
    if mode == 2:
      assert s.plan_info()["ring_mode"] == 2, s.plan_info()
    for _ in range(2):  # odd and even passes walk the tiles in opposite directions
      lp, g = s.logp_grad(theta)
    got[mode] = (float(lp.cpu()[0]), g.cpu().numpy())
    s.close()
    assert abs(got[mode][0] - lp64) <= REL_LOGP * abs(lp64), (mode, got[mode][0], lp64)
    assert _rel(got[mode][1], g64) <= REL_GRAD, (mode, _rel(got[mode][1], g64))


@pytest.mark.parametrize("plan", [1, 2])
@pytest.mark.parametrize("N,D,bias,fam,T,L,eps", [
    (9000, 16, False, o.BERNOULLI_LOGIT, 12, 10, 0.004),
    (5003, 40, True, o.BERNOULLI_LOGIT, 10, 5, 0.006),
    (9000, 54, False, o.BERNOULLI_LOGIT, 12, 10, 0.004),
    (3000, 12, True, o.POISSON_LOG, 10, 3, 0.003),
])
def test_ldg_tiles_run_matches_oracle(N, D, bias, fam, T, L, eps, plan):
  import torch
  X, y, spec = _mk(N, D, bias, fam, seed=11 * N + D, lik_scale=0.5)
  P = spec.n_params
  r0, u = o.synth_draws(T, P, seed=4)
  p64 = np.zeros((T, P))
  infos, nacc = o.run(X, y, p64, r0, u, eps, L, spec, np.float64)
  with _ring(2):
    s = _sampler(X, y, spec, plan=plan)
  assert s.plan_info()["ring_mode"] == 2
  params = torch.zeros(T, P, device="cuda")
  sc, pos = s.set_trace(T)
  s.run(params, 0, T, eps, L, r0=torch.tensor(r0), u=torch.tensor(u))
  sc, pos = sc.cpu().numpy(), pos.cpu().numpy()
  for i, info in enumerate(infos):
    assert abs(sc[i, 1] - info.logp_new) <= REL_LOGP * abs(info.logp_new) + 1e-6, (i, sc[i], info)
    assert _rel(pos[i], info.proposal) <= REL_POS, (i, _rel(pos[i], info.proposal))
    if bool(sc[i, 6] > 0.5) != info.accept:
      assert abs(info.log_u - info.ratio) < 1e-3, (i, info)
      break
  else:
    assert s.read_state()[0] == nacc
    assert _rel(params.cpu().numpy(), p64) <= REL_POS
  s.close()


def test_persistent_and_stepwise_plans_bit_identical_on_ldg_tiles():
  import torch
  X, y, spec = _mk(40000, 32, False, o.BERNOULLI_LOGIT, seed=9)
  out = []
  for plan in (1, 2):
    with _ring(2):
      s = _sampler(X, y, spec, plan=plan)
    s.seed(77)
    params = torch.zeros(6, 32, device="cuda")
    s.run(params, 0, 6, 0.002, 7)
    out.append(params.cpu().numpy().copy())
    s.close()
  assert np.array_equal(out[0], out[1])


def test_plan_rule_picks_ldg_tiles_for_the_measured_shapes():
  """make_plan (csrc/edhmc.cu): ring mode 2 for D <= 48 and D = 63, 64; for 49 <= D <= 62 ring mode 2 when at least 30 % of
  a CTA's tiles stay resident in shared / tensor memory (cfg 2), else the CTA-wide ring."""
  with _ring(None):
    for D, N, want in ((8, 5000, 2), (32, 5000, 2), (48, 5000, 2), (54, 5000, 2), (54, 581012, 2), (54, 2000000, 1),
                       (60, 1000000, 1), (64, 5000, 2), (100, 5000, None)):
      X, y, spec = _mk(N, D, False, o.BERNOULLI_LOGIT, seed=D)
      s = _sampler(X, y, spec)
      rm = s.plan_info()["ring_mode"]
      s.close()
      if want is None:
        assert rm != 2
      else:
        assert rm == want, (D, rm)


# Mini-batch SGLD / SGHMC (row windows of the caller's X, which take the row-major plan even when the full-data passes
# read the re-laid copy) is covered by tests/test_gpu_sgmcmc.py::test_sgmcmc_matches_oracle (D = 20: ring mode 2 by the rule).
