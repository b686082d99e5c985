"""Builds edward_b200/lib/libedhmc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libedhmc.so")
SOURCES = ["edhmc.cu"]
HEADERS = ["ptx.cuh", "common.cuh", "stream.cuh", "chain.cuh", os.path.join("..", "..", "include", "edhmc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found; libedhmc.so cannot be built")


def is_stale() -> bool:
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  for f in SOURCES + HEADERS:
    p = os.path.join(CSRC, f)
    if os.path.exists(p) and os.path.getmtime(p) > t:
      return True
  return False


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not is_stale():
    return LIB
  os.makedirs(LIBDIR, exist_ok=True)
  cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
      ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    raise RuntimeError("nvcc failed building libedhmc.so")
  if verbose:
    sys.stderr.write(res.stderr)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
