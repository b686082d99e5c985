"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement of the reference's schedule) runs
here and prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
  e = dict(os.environ, PYTHONPATH=ROOT)
  if env:
    e.update(env)
  res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, env=e,
                       timeout=900)
  assert res.returncode == 0, res.stderr[-2000:]
  return res.stdout


def test_reference_arm_prints_one_contract_line():
  out = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
  lines = [ln for ln in out.splitlines() if ln.startswith("{")]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
    assert key in d, key
  assert d["impl"] == "reference" and d["metric"] == "hmc_leapfrog_steps_per_s" and d["value"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "cfg2" in d["config"]["workload"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
  out = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
  assert out.strip() == ""


def test_help_lists_the_contract_flags():
  out = _run("--help")
  for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--collective"):
    assert flag in out
