"""The C-ABI library builds, loads, and exports every symbol include/edhmc.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, "include", "edhmc.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(edhmc_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
  from edward_b200 import _C
  assert _declared() == sorted(_C.EXPORTS)


def test_library_builds_and_exports_all_symbols():
  from edward_b200 import _C, build
  lib_path = build.build()
  assert os.path.exists(lib_path)
  L = _C.lib()
  for name in _declared():
    assert hasattr(L, name), name
  assert L.edhmc_version() == 1
  assert L.edhmc_last_error() == b""


def test_cfg_struct_layout_matches_header():
  """Field order/types of the ctypes mirror follow the header's edhmc_cfg."""
  from edward_b200 import _C
  src = open(os.path.join(ROOT, "include", "edhmc.h")).read()
  body = re.search(r"typedef struct edhmc_cfg \{(.*?)\} edhmc_cfg;", src, flags=re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  names = [re.search(r"(\w+)(\[\d+\])?\s*;", line).group(1) for line in body.split("\n") if ";" in line]
  assert names == [f[0] for f in _C.Cfg._fields_]
  assert ctypes.sizeof(_C.Cfg) == 96


def test_invalid_arguments_fail_without_a_gpu():
  """Argument validation precedes any CUDA call, with a readable message."""
  from edward_b200 import _C
  L = _C.lib()
  h = ctypes.c_void_p()
  cfg = _C.Cfg()
  cfg.n_rows = 10
  cfg.n_rows_global = 10
  cfg.n_features = 0
  cfg.ldx = 0
  assert L.edhmc_create(ctypes.byref(h), ctypes.byref(cfg)) == _C.ERR_INVALID
  assert b"n_features" in L.edhmc_last_error()
  assert L.edhmc_create(None, None) == _C.ERR_INVALID
  assert L.edhmc_run(None, None, 0, 0, 0, 0, 0.1, 1, None, None, None) == _C.ERR_INVALID
  assert L.edhmc_destroy(None) == 0


def test_product_does_not_import_oracle():
  """The product path must never route through the oracle: no Python import of it, no native reference
  to the CPU port, and the only library the C ABI dlopens is NCCL."""
  pkg = os.path.join(ROOT, "edward_b200")
  imp = re.compile(r"^\s*(import|from)\s+(hmc_oracle|ref_c|oracle)\b", re.M)
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      path = os.path.join(dirpath, f)
      if f.endswith(".py"):
        text = open(path).read()
        assert not imp.search(text), path
        assert "oracle/" not in text.replace("oracle/hmc_oracle.py", ""), path
      elif f.endswith((".cu", ".cuh", ".h")):
        text = open(path).read()
        assert "hmc_ref" not in text, path
        for m in re.findall(r'dlopen\(([^,]+),', text):
          assert m.strip() == "n", (path, m)  # the loop variable over {"libnccl.so.2", "libnccl.so"}


def test_no_gpu_means_loud_failure():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  import numpy as np
  from edward_b200 import _C, engine
  with pytest.raises(_C.EdhmcError):
    engine.GLMSampler(engine.GLMSpec(2), np.zeros((4, 2), np.float32), np.zeros(4, np.int32))


def test_generated_kernel_instances_are_in_sync(tmp_path, monkeypatch):
  """csrc/inst_*.cu and inst_table.inc are what csrc/gen_inst.py generates (nobody edited them by hand, nobody forgot to
  re-run the generator after changing the tiers)."""
  import importlib.util
  import shutil
  csrc = os.path.join(ROOT, "edward_b200", "csrc")
  spec = importlib.util.spec_from_file_location("gen_inst", os.path.join(csrc, "gen_inst.py"))
  gen = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(gen)
  monkeypatch.setattr(gen, "HERE", str(tmp_path))
  gen.main()
  produced = sorted(os.listdir(tmp_path))
  committed = sorted(f for f in os.listdir(csrc) if f.startswith("inst_"))
  assert produced == committed
  for f in produced:
    assert open(os.path.join(tmp_path, f)).read() == open(os.path.join(csrc, f)).read(), f
  # every feature count the C ABI accepts has a kernel: the planner's rule restated in gen_inst.select
  for V in (1, 2, 4):
    for D in range(1, gen.MAXD + 1):
      r = gen.select(D, V)
      assert r is not None and r[1] is not None, (D, V)


def test_wide_chain_tiling_probe():
  """Host-side planner of the two-GEMM vectorised-chain path (chains_wide.cu::mcw_plan), no GPU needed."""
  from edward_b200 import _C
  L = _C.lib()
  out = (ctypes.c_int64 * 8)()

  def probe(n, d, c, sms=148):
    assert L.edhmc_chains_plan_probe(n, d, c, sms, out) == 0, L.edhmc_last_error()
    return dict(zip(["nct", "Kp1", "nrt", "nft", "NB2", "g1", "splits", "Dp2"], list(out)))

  p = probe(1_250_000, 1000, 1024)  # config 5, one GPU's shard
  assert p["nct"] == 8 and p["Kp1"] == 1008 and p["nrt"] == 9766  # 128-row tiles of GEMM 1
  assert (p["nft"], p["NB2"], p["splits"], p["Dp2"]) == (6, 176, 3, 1056)  # 144 of 148 SMs instead of 128
  assert p["g1"] == 18
  for n, d, c in [(300, 1000, 128), (2111, 1000, 256), (5000, 200, 128), (20000, 72, 128), (1000, 54, 128), (7, 2048, 256)]:
    p = probe(n, d, c)
    assert p["NB2"] % 16 == 0 and 16 <= p["NB2"] <= 256 and p["nft"] * p["NB2"] == p["Dp2"] >= d
    assert p["Kp1"] % 16 == 0 and p["Kp1"] >= d and p["nrt"] * 128 >= n
    assert 1 <= p["g1"] <= max(1, p["nrt"]) and p["nct"] * p["g1"] <= 148
    assert p["splits"] >= 1 and p["nct"] * p["nft"] * p["splits"] <= max(148, p["nct"] * p["nft"])
  assert probe(100, 8, 100)["nct"] == 1 and probe(1000, 200, 130)["nct"] == 2  # any chain count: whole 128-chain tiles
  assert L.edhmc_chains_plan_probe(100, 8, 1, 148, out) == _C.ERR_INVALID
