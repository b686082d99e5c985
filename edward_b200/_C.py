"""ctypes binding of libedhmc.so (include/edhmc.h). No torch types cross this boundary: device buffers
are passed as raw addresses (`tensor.data_ptr()`), the stream as a cudaStream_t handle.

There is no CPU path: importing works anywhere (so the CPU test-suite can check the exported symbols),
but every compute entry point needs a CUDA device and raises `EdhmcError` otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EDHMC_LIB_PATH") or os.path.join(HERE, "lib", "libedhmc.so")  # override: A/B of two builds

EDHMC_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NONFINITE, ERR_RANGE, ERR_STATE, ERR_COMM, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7
BERNOULLI_LOGIT, NORMAL_IDENTITY, POISSON_LOG = 0, 1, 2
PRIOR_NORMAL, PRIOR_BETA_LOGIT = 0, 1
Y_I32, Y_F32, Y_U8, Y_F64 = 0, 1, 2, 3
F32, F64 = 0, 1
PLAN_AUTO, PLAN_PERSISTENT, PLAN_STEPWISE = 0, 1, 2

EXPORTS = [
    "edhmc_version", "edhmc_last_error", "edhmc_create", "edhmc_destroy", "edhmc_bind_data",
    "edhmc_logp_grad", "edhmc_run", "edhmc_set_trace", "edhmc_read_state", "edhmc_reset", "edhmc_seed",
    "edhmc_comm_unique_id", "edhmc_comm_init", "edhmc_peer_export", "edhmc_peer_attach", "edhmc_peer_detach", "edhmc_plan_info",
    "edhmc_sgmcmc_run", "edhmc_run_chains", "edhmc_logp_grad_chains", "edhmc_read_chain_state", "edhmc_set_chain_trace", "edhmc_set_chain_debug", "edhmc_chains_plan_probe", "edhmc_set_timeline", "edhmc_probe_read", "edhmc_comm_cached", "edhmc_comm_release", "edhmc_set_prior_kinds",
    "edhmc_bind_data_f64", "edhmc_logp_grad_f64", "edhmc_run_f64", "edhmc_set_trace_f64", "edhmc_predictive",
]


class EdhmcError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__("libedhmc error %d: %s" % (code, msg))
    self.code = code


class NonFiniteError(EdhmcError, ValueError):
  """NaN/Inf in the data (the reference's ed.dot raises InvalidArgumentError, util/tensorflow.py:27-36)."""


class RangeError(EdhmcError, IndexError):
  """update() past the last Empirical row (the reference's scatter_update fails, hmc.py:125)."""


class Cfg(C.Structure):
  _fields_ = [
      ("n_rows", C.c_int64),
      ("n_rows_global", C.c_int64),
      ("n_features", C.c_int32),
      ("ldx", C.c_int64),
      ("has_bias", C.c_int32),
      ("family", C.c_int32),
      ("y_dtype", C.c_int32),
      ("lik_scale", C.c_float),
      ("prior_loc_host", C.POINTER(C.c_float)),
      ("prior_scale_host", C.POINTER(C.c_float)),
      ("device", C.c_int32),
      ("plan", C.c_int32),
      ("debug", C.c_int32),
      ("n_chains", C.c_int32),
      ("dtype", C.c_int32),
      ("reserved", C.c_int32 * 3),
  ]


_lib = None


def lib():
  """Loads libedhmc.so; fails loudly if it has not been built (python -m edward_b200.build)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libedhmc.so is missing at %s — build it with `python -m edward_b200.build` (needs nvcc). "
        "edward_b200 has no CPU fallback." % LIB_PATH)
  L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
  vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
  L.edhmc_version.restype = C.c_int
  L.edhmc_last_error.restype = C.c_char_p
  L.edhmc_create.argtypes = [C.POINTER(vp), C.POINTER(Cfg)]
  L.edhmc_destroy.argtypes = [vp]
  L.edhmc_bind_data.argtypes = [vp, vp, vp, C.c_int, vp]
  L.edhmc_logp_grad.argtypes = [vp, vp, vp, vp, vp]
  L.edhmc_run.argtypes = [vp, vp, i64, i64, i64, i64, f32, i32, vp, vp, vp]
  L.edhmc_set_trace.argtypes = [vp, vp, vp]
  L.edhmc_read_state.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_double), vp]
  L.edhmc_reset.argtypes = [vp, vp]
  L.edhmc_seed.argtypes = [vp, C.c_uint64]
  L.edhmc_comm_unique_id.argtypes = [vp]
  L.edhmc_comm_init.argtypes = [vp, vp, i32, i32]
  L.edhmc_peer_export.argtypes = [vp, vp]
  L.edhmc_peer_attach.argtypes = [vp, vp, i32, i32]
  L.edhmc_peer_detach.argtypes = [vp]
  L.edhmc_plan_info.argtypes = [vp, C.POINTER(i64), i32]
  L.edhmc_sgmcmc_run.argtypes = [vp, i32, vp, i64, i64, i64, i64, f32, f32, f32, vp, vp, vp, i64, vp]
  L.edhmc_run_chains.argtypes = [vp, vp, i64, i64, i64, f32, i32, vp, vp, vp]
  L.edhmc_logp_grad_chains.argtypes = [vp, vp, vp, vp, vp]
  L.edhmc_read_chain_state.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_double), vp]
  L.edhmc_set_chain_trace.argtypes = [vp, vp]
  L.edhmc_set_chain_debug.argtypes = [vp, vp]
  L.edhmc_chains_plan_probe.argtypes = [i64, i32, i32, i32, C.POINTER(i64)]
  L.edhmc_set_timeline.argtypes = [vp, vp, i32]
  L.edhmc_probe_read.argtypes = [vp, i64, i32, i32, vp, vp]
  L.edhmc_comm_cached.argtypes = [i32, i32, i32]
  L.edhmc_comm_release.argtypes = [i32]
  L.edhmc_set_prior_kinds.argtypes = [vp, C.POINTER(i32)]
  L.edhmc_bind_data_f64.argtypes = [vp, vp, vp, C.c_int, vp]
  L.edhmc_logp_grad_f64.argtypes = [vp, vp, vp, vp, vp]
  L.edhmc_run_f64.argtypes = [vp, vp, i64, i64, i64, i64, C.c_double, i32, vp, vp, vp]
  L.edhmc_set_trace_f64.argtypes = [vp, vp, vp]
  L.edhmc_predictive.argtypes = [vp, i64, i64, i32, vp, i32, i32, f32, vp, i64, vp, vp, i32, i32, vp, vp, i32, vp]
  for name in EXPORTS:
    if name not in ("edhmc_last_error",):
      getattr(L, name).restype = C.c_int
  _lib = L
  return L


def check(rc):
  if rc >= 0:
    return rc
  msg = lib().edhmc_last_error().decode("utf-8", "replace")
  if rc == ERR_NONFINITE:
    raise NonFiniteError(rc, msg)
  if rc == ERR_RANGE:
    raise RangeError(rc, msg)
  raise EdhmcError(rc, msg)
