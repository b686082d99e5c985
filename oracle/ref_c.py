"""ctypes wrapper of oracle/hmc_ref.c (C restatement of the reference's CPU schedule).

TEST / BASELINE INFRASTRUCTURE ONLY — see the header of hmc_ref.c. PARITY UNPINNED (no TensorFlow here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libhmc_ref.so")
_lib = None


def build(force=False):
  src = os.path.join(HERE, "hmc_ref.c")
  if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
    subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
  return LIB


def lib():
  global _lib
  if _lib is None:
    build()
    L = C.CDLL(LIB)
    fp, vp, i64, i32, f32 = C.POINTER(C.c_float), C.c_void_p, C.c_int64, C.c_int, C.c_float
    L.hmc_ref_run.argtypes = [fp, vp, i64, i32, i32, i32, f32, fp, fp, fp, i64, i64, i64, f32, i32, fp, fp, i32,
                              C.POINTER(i64), C.POINTER(C.c_double)]
    L.hmc_ref_run.restype = C.c_int
    L.hmc_ref_logp_grad.argtypes = [fp, vp, i64, i32, i32, i32, f32, fp, fp, fp, fp, fp]
    L.hmc_ref_logp_grad.restype = C.c_int
    L.hmc_ref_num_threads.restype = C.c_int
    _lib = L
  return _lib


def _fp(a):
  return a.ctypes.data_as(C.POINTER(C.c_float))


def _prep(X, y, spec):
  X = np.ascontiguousarray(X, np.float32)
  y = np.ascontiguousarray(y, np.float32 if spec.family == 1 else np.int32)
  loc = np.ascontiguousarray(spec.prior_loc, np.float32)
  scale = np.ascontiguousarray(spec.prior_scale, np.float32)
  return X, y, loc, scale


def num_threads():
  return int(lib().hmc_ref_num_threads())


def logp_grad(X, y, theta, spec):
  X, y, loc, scale = _prep(X, y, spec)
  theta = np.ascontiguousarray(theta, np.float32)
  lp = C.c_float(0)
  g = np.zeros(spec.n_params, np.float32)
  rc = lib().hmc_ref_logp_grad(_fp(X), y.ctypes.data_as(C.c_void_p), X.shape[0], spec.n_features, int(spec.has_bias),
                               spec.family, float(spec.lik_scale), _fp(loc), _fp(scale), _fp(theta), C.byref(lp), _fp(g))
  if rc == -3:
    raise ValueError("InvalidArgumentError: Tensor had NaN or Inf values")
  assert rc == 0, rc
  return float(lp.value), g


def run(X, y, params, r0, u, step_size, n_steps, spec, t0=0, n_iter=None, check_numerics=True, trace=True):
  """In place on params [T,P] float32. Returns (n_accept, trace[n_iter,8] or None)."""
  X, y, loc, scale = _prep(X, y, spec)
  assert params.dtype == np.float32 and params.flags.c_contiguous
  T = params.shape[0]
  if n_iter is None:
    n_iter = T - t0
  r0 = np.ascontiguousarray(r0, np.float32)
  u = np.ascontiguousarray(u, np.float32)
  tr = np.zeros((n_iter, 8), np.float64) if trace else None
  nacc = C.c_int64(0)
  rc = lib().hmc_ref_run(_fp(X), y.ctypes.data_as(C.c_void_p), X.shape[0], spec.n_features, int(spec.has_bias),
                         spec.family, float(spec.lik_scale), _fp(loc), _fp(scale), _fp(params), T, t0, n_iter,
                         float(step_size), int(n_steps), _fp(r0), _fp(u), int(check_numerics), C.byref(nacc),
                         tr.ctypes.data_as(C.POINTER(C.c_double)) if trace else None)
  if rc == -3:
    raise ValueError("InvalidArgumentError: Tensor had NaN or Inf values")
  if rc == -4:
    raise IndexError("scatter_update index out of range")
  assert rc == 0, rc
  return int(nacc.value), tr
