"""Builds edward_b200/lib/libedhmc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The kernel template is instantiated in one translation unit per (lanes-per-row, vector-width) pair
(csrc/inst_g*_v*.cu) so the units compile in parallel; objects land in a scratch directory under /tmp.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("EDHMC_OBJ_DIR", os.path.join("/tmp", "edhmc_obj_%08x" % zlib.crc32(CSRC.encode())))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libedhmc.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found; libedhmc.so cannot be built")


def sources():
  return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
  return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "edhmc.h")]


def _stale(target, srcs) -> bool:
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def is_stale() -> bool:
  return _stale(LIB, sources() + _deps())


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
  if not force and not is_stale():
    return LIB
  nvcc = _nvcc()
  os.makedirs(OBJ, exist_ok=True)
  os.makedirs(LIBDIR, exist_ok=True)
  deps = _deps()

  def compile_one(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if not force and not _stale(obj, [src] + deps):
      return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
      raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
    return obj, res.stderr

  with ThreadPoolExecutor(max_workers=jobs or min(10, os.cpu_count() or 4)) as ex:
    results = list(ex.map(compile_one, sources()))
  if verbose:
    for _, log in results:
      sys.stderr.write(log)
  objs = [o for o, _ in results]
  cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", LIB] + objs + ["-ldl"]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("link failed:\n%s%s" % (res.stdout, res.stderr))
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
