"""GLMSampler — owns one libedhmc handle: the device-resident data, the chain state and the HMC loop.

This is the object `ed.HMC` builds in `initialize()` in place of the reference's TensorFlow graph
(edward/inferences/hmc.py:61-130). torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _C
from . import graph as _graph


@dataclass
class GLMSpec:
  """theta = [w(D), b?] ~ Normal(prior_loc, prior_scale);  y ~ family(X w + b)."""
  n_features: int
  has_bias: bool = False
  family: int = _C.BERNOULLI_LOGIT
  prior_loc: Optional[np.ndarray] = None
  prior_scale: Optional[np.ndarray] = None
  lik_scale: float = 1.0

  @property
  def n_params(self) -> int:
    return self.n_features + (1 if self.has_bias else 0)


def _require_cuda(device) -> torch.device:
  if not torch.cuda.is_available():
    raise _C.EdhmcError(_C.ERR_CUDA, "no CUDA device: edward_b200 runs its HMC path only on the GPU "
                        "(there is no CPU fallback)")
  dev = torch.device(device if device is not None else "cuda")
  if dev.type != "cuda":
    raise _C.EdhmcError(_C.ERR_INVALID, "device must be a CUDA device, got %s" % dev)
  if dev.index is None:
    dev = torch.device("cuda", torch.cuda.current_device())
  return dev


def _stream_ptr(dev) -> int:
  return torch.cuda.current_stream(dev).cuda_stream


_Y_DTYPES = {torch.int32: _C.Y_I32, torch.float32: _C.Y_F32, torch.uint8: _C.Y_U8}
_Y_DTYPES_F64 = {torch.int32: _C.Y_I32, torch.float64: _C.Y_F64}


class GLMSampler:
  """One chain of HMC over a GLM. Row-sharded when `comm` (a torch.distributed group spec) is given."""

  def __init__(self, spec: GLMSpec, X, y, device=None, plan: int = _C.PLAN_AUTO, debug: bool = False,
               check_finite: bool = True, n_rows_global: Optional[int] = None, n_chains: int = 1, dtype=torch.float32):
    """`dtype` torch.float64 selects the compact float64 path (edhmc_dtype, hmc_test.py:93-97): X, params, draws and
    gradients are then double; one chain, HMC only."""
    self.lib = _C.lib()
    self.dev = _require_cuda(device)
    self.spec = spec
    if dtype not in (torch.float32, torch.float64):
      raise TypeError("dtype must be torch.float32 or torch.float64, got %r" % (dtype,))
    self.dtype = dtype
    f64 = dtype == torch.float64
    P = spec.n_params
    loc = np.zeros(P, np.float32) if spec.prior_loc is None else np.ascontiguousarray(spec.prior_loc, np.float32).reshape(P)
    scale = np.ones(P, np.float32) if spec.prior_scale is None else np.ascontiguousarray(spec.prior_scale, np.float32).reshape(P)
    # the uploads are asynchronous when the caller's arrays are pinned: handle creation and planning below overlap with them
    # (bind_data's finite check, or the explicit synchronize at the end, completes them before __init__ returns)
    self.X = self._to_device(X, dtype, non_blocking=True)
    if self.X.dim() != 2 or self.X.shape[1] != spec.n_features:
      raise TypeError("X must have shape [N, %d], got %s" % (spec.n_features, tuple(self.X.shape)))
    if self.X.stride(1) != 1 or self.X.data_ptr() % 16 != 0:
      self.X = self.X.contiguous()
    yt = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
    ymap = _Y_DTYPES_F64 if f64 else _Y_DTYPES
    if yt.dtype not in ymap:
      # the reference casts observed data to the random variable's dtype (inference.py:88-95)
      yt = yt.to(torch.int32 if spec.family != _C.NORMAL_IDENTITY else dtype)
    self.y = yt.to(self.dev, non_blocking=True).contiguous()
    if self.y.dim() != 1 or self.y.shape[0] != self.X.shape[0]:
      raise TypeError("y must have shape [%d], got %s" % (self.X.shape[0], tuple(self.y.shape)))
    self.n_rows = int(self.X.shape[0])
    self.P = P

    cfg = _C.Cfg()
    cfg.n_rows = self.n_rows
    cfg.n_rows_global = int(n_rows_global if n_rows_global is not None else self.n_rows)
    cfg.n_features = spec.n_features
    cfg.ldx = int(self.X.stride(0)) if self.n_rows > 1 else max(int(self.X.stride(0)), spec.n_features)
    cfg.has_bias = 1 if spec.has_bias else 0
    cfg.family = int(spec.family)
    cfg.y_dtype = ymap[self.y.dtype]
    cfg.dtype = _C.F64 if f64 else _C.F32
    cfg.lik_scale = float(spec.lik_scale)
    cfg.prior_loc_host = loc.ctypes.data_as(C.POINTER(C.c_float))
    cfg.prior_scale_host = scale.ctypes.data_as(C.POINTER(C.c_float))
    cfg.device = self.dev.index
    cfg.plan = int(plan)
    cfg.debug = 1 if debug else 0
    cfg.n_chains = int(n_chains) if n_chains and n_chains > 1 else 0
    self.n_chains = max(1, int(n_chains or 1))
    self._h = C.c_void_p()
    _C.check(self.lib.edhmc_create(C.byref(self._h), C.byref(cfg)))
    with torch.cuda.device(self.dev):
      bind = self.lib.edhmc_bind_data_f64 if f64 else self.lib.edhmc_bind_data
      _C.check(bind(self._h, self.X.data_ptr(), self.y.data_ptr(), 1 if check_finite else 0,
                                        _stream_ptr(self.dev)))
    if not check_finite:
      torch.cuda.current_stream(self.dev).synchronize()  # the caller may reuse its host arrays once we return
    self._trace = None
    self.nranks = 1

  def _to_device(self, a, dtype, non_blocking=False):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    return t.to(device=self.dev, dtype=dtype, non_blocking=non_blocking)

  # ---- row shards (extension) -----------------------------------------------------------------
  def init_comm(self, nranks: int, rank: int, group=None, peers=None):
    """Attaches this handle to the process-wide NCCL communicator of its device (created on first use; the unique id
    travels over torch.distributed) and, unless `peers` is False, to the peer-mapped inboxes of the in-kernel
    all-reduce. Later samplers of the same process reuse both without any collective call."""
    from .sharding import allgather_bytes, broadcast_bytes
    if peers is None:
      peers = os.environ.get("EDHMC_COLLECTIVE", "peer") != "nccl"
    want_peers = bool(peers) and nranks <= 8
    cached = int(self.lib.edhmc_comm_cached(self.dev.index, int(nranks), int(rank)))  # 0 none, 1 comm, 2 comm + peers
    # every rank must take the same branch: one small all-gather settles on the lowest level any rank holds
    level = min(allgather_bytes(bytes([cached]), group=group))
    with torch.cuda.device(self.dev):
      if level >= 1:
        _C.check(self.lib.edhmc_comm_init(self._h, None, int(nranks), int(rank)))
      else:
        buf = (C.c_char * 128)()
        if rank == 0:
          _C.check(self.lib.edhmc_comm_unique_id(C.cast(buf, C.c_void_p)))
        uid = broadcast_bytes(bytes(buf), src=0, group=group)
        idbuf = C.create_string_buffer(uid, 128)
        _C.check(self.lib.edhmc_comm_init(self._h, C.cast(idbuf, C.c_void_p), int(nranks), int(rank)))
    self.nranks = int(nranks)
    self.peer_exchange = False
    if want_peers:
      if level >= 2:
        with torch.cuda.device(self.dev):
          _C.check(self.lib.edhmc_peer_attach(self._h, None, int(nranks), int(rank)))
        self.peer_exchange = True
      else:
        self.peer_exchange = self._attach_peers(nranks, rank, group)

  def _attach_peers(self, nranks, rank, group):
    """Maps every rank's inbox into this process (cudaIpc) so that edhmc_run can all-reduce the shard totals
    inside its persistent kernel. If the mapping is impossible on this machine (no peer access between the
    GPUs), every rank falls back together to the per-pass launch + ncclAllReduce plan — still all on the GPUs."""
    import sys
    from .sharding import allgather_bytes
    hbuf = (C.c_char * 64)()
    with torch.cuda.device(self.dev):
      rc = self.lib.edhmc_peer_export(self._h, C.cast(hbuf, C.c_void_p))
    table = allgather_bytes(bytes(hbuf) if rc == 0 else b"\0" * 64, group=group)
    ok = rc == 0
    if ok:
      tbuf = C.create_string_buffer(table, 64 * nranks)
      with torch.cuda.device(self.dev):
        ok = self.lib.edhmc_peer_attach(self._h, C.cast(tbuf, C.c_void_p), int(nranks), int(rank)) == 0
    msg = None if ok else self.lib.edhmc_last_error().decode("utf-8", "replace")
    votes = allgather_bytes(b"\x01" if ok else b"\x00", group=group)
    all_ok = all(v == 1 for v in votes)
    if not all_ok:
      if msg:
        sys.stderr.write("edhmc: peer exchange unavailable (%s); using ncclAllReduce per pass\n" % msg)
      with torch.cuda.device(self.dev):
        _C.check(self.lib.edhmc_peer_detach(self._h))
    return all_ok

  # ---- evaluation ------------------------------------------------------------------------------
  def logp_grad(self, theta):
    """log p(y, theta) (float64 scalar tensor) and its gradient ([P], the model's dtype) — hmc.py:161-192,199."""
    th = self._to_device(theta, self.dtype).contiguous().reshape(self.P)
    logp = torch.empty(1, dtype=torch.float64, device=self.dev)
    grad = torch.empty(self.P, dtype=self.dtype, device=self.dev)
    with torch.cuda.device(self.dev):
      fn = self.lib.edhmc_logp_grad_f64 if self.dtype == torch.float64 else self.lib.edhmc_logp_grad
      _C.check(fn(self._h, th.data_ptr(), logp.data_ptr(), grad.data_ptr(), _stream_ptr(self.dev)))
    return logp, grad

  def run(self, params: torch.Tensor, t0: int, n_iter: int, step_size: float, n_steps: int,
          r0: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None):
    """n_iter transitions in place on `params` [T, >=P] (device, the model's dtype, row-major)."""
    if params.device != self.dev or params.dtype != self.dtype:
      raise TypeError("params must be a %s tensor on %s" % (str(self.dtype).replace("torch.", ""), self.dev))
    if params.dim() == 1:
      params = params.view(-1, 1)
    if params.stride(1) != 1:
      raise TypeError("params rows must be contiguous")
    r0p = up = None
    if r0 is not None:
      r0 = self._to_device(r0, self.dtype).contiguous()
      if r0.numel() < n_iter * self.P:
        raise ValueError("r0 must hold n_iter*P momentum draws")
      r0p = r0.data_ptr()
    if u is not None:
      u = self._to_device(u, self.dtype).contiguous()
      if u.numel() < n_iter:
        raise ValueError("u must hold n_iter uniforms")
      up = u.data_ptr()
    with torch.cuda.device(self.dev):
      fn = self.lib.edhmc_run_f64 if self.dtype == torch.float64 else self.lib.edhmc_run
      _C.check(fn(self._h, params.data_ptr(), int(params.stride(0)), int(params.shape[0]), int(t0),
                  int(n_iter), float(step_size), int(n_steps), r0p, up, _stream_ptr(self.dev)))
    _graph.bump_device_epoch()  # the sample store changed: host mirrors of its variables are stale

  # ---- SGLD / SGHMC on the same gradient kernel (sgld.py:52-87, sghmc.py:58-96) -----------------
  def sgmcmc_run(self, kind: str, params: torch.Tensor, t0: int, n_iter: int, step_size: float, friction: float = 0.1,
                 lik_factor: float = 1.0, prior_factor=None, velocity: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, batch_rows: int = 0):
    if params.device != self.dev or params.dtype != torch.float32:
      raise TypeError("params must be a float32 tensor on %s" % self.dev)
    if params.dim() == 1:
      params = params.view(-1, 1)
    k = {"sgld": 0, "sghmc": 1}[kind]
    pf = None
    if prior_factor is not None:
      self._pf = self._to_device(prior_factor, torch.float32).contiguous().reshape(self.P)
      pf = self._pf.data_ptr()
    if k == 1 and velocity is None:
      raise ValueError("SGHMC needs a velocity tensor")
    nz = None
    if noise is not None:
      noise = self._to_device(noise, torch.float32).contiguous()
      if noise.numel() < n_iter * self.P:
        raise ValueError("noise must hold n_iter*P draws")
      nz = noise.data_ptr()
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_sgmcmc_run(self._h, k, params.data_ptr(), int(params.stride(0)), int(params.shape[0]), int(t0),
                                         int(n_iter), float(step_size), float(friction), float(lik_factor), pf,
                                         velocity.data_ptr() if velocity is not None else None, nz, int(batch_rows),
                                         _stream_ptr(self.dev)))
    _graph.bump_device_epoch()

  # ---- C vectorised chains (extension) ---------------------------------------------------------
  def logp_grad_chains(self, theta):
    """theta [C, P] → (logp [C] float64, grad [C, P] float32), one dense contraction on the tensor cores."""
    th = self._to_device(theta, torch.float32).contiguous().reshape(self.n_chains, self.P)
    logp = torch.empty(self.n_chains, dtype=torch.float64, device=self.dev)
    grad = torch.empty(self.n_chains, self.P, dtype=torch.float32, device=self.dev)
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_logp_grad_chains(self._h, th.data_ptr(), logp.data_ptr(), grad.data_ptr(), _stream_ptr(self.dev)))
    return logp, grad

  def run_chains(self, params: torch.Tensor, t0: int, n_iter: int, step_size: float, n_steps: int,
                 r0: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None):
    """n_iter transitions of all C chains in place on `params` [T, C, P] (device, float32, contiguous)."""
    if params.device != self.dev or params.dtype != torch.float32 or not params.is_contiguous():
      raise TypeError("params must be a contiguous float32 tensor on %s" % self.dev)
    if params.dim() != 3 or params.shape[1] != self.n_chains or params.shape[2] != self.P:
      raise TypeError("params must have shape [T, %d, %d]" % (self.n_chains, self.P))
    r0p = up = None
    if r0 is not None:
      r0 = self._to_device(r0, torch.float32).contiguous()
      if r0.numel() < n_iter * self.n_chains * self.P:
        raise ValueError("r0 must hold n_iter*C*P momentum draws")
      r0p = r0.data_ptr()
    if u is not None:
      u = self._to_device(u, torch.float32).contiguous()
      if u.numel() < n_iter * self.n_chains:
        raise ValueError("u must hold n_iter*C uniforms")
      up = u.data_ptr()
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_run_chains(self._h, params.data_ptr(), int(params.shape[0]), int(t0), int(n_iter),
                                         float(step_size), int(n_steps), r0p, up, _stream_ptr(self.dev)))
    _graph.bump_device_epoch()

  def read_chain_state(self):
    n = (C.c_int64 * self.n_chains)()
    lp = (C.c_double * self.n_chains)()
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_read_chain_state(self._h, n, lp, _stream_ptr(self.dev)))
    return np.array(n[:], np.int64), np.array(lp[:], np.float64)

  def set_chain_trace(self, n_iter: int):
    tr = torch.zeros(n_iter, self.n_chains, 8, dtype=torch.float64, device=self.dev)
    self._chain_trace = tr
    _C.check(self.lib.edhmc_set_chain_trace(self._h, tr.data_ptr()))
    return tr

  def set_trace(self, n_iter: int):
    sc = torch.zeros(n_iter, 8, dtype=torch.float64, device=self.dev)
    pos = torch.zeros(n_iter, self.P, dtype=self.dtype, device=self.dev)
    self._trace = (sc, pos)
    fn = self.lib.edhmc_set_trace_f64 if self.dtype == torch.float64 else self.lib.edhmc_set_trace
    _C.check(fn(self._h, sc.data_ptr(), pos.data_ptr()))
    return sc, pos

  def set_timeline(self, n_passes: int):
    """Development aid: per-CTA clock64 stamps of the first n_passes passes of the next persistent run()
    ([n_passes, grid, 32] int64; see edhmc_set_timeline). n_passes = 0 switches it off."""
    if not n_passes:
      self._timeline = None
      _C.check(self.lib.edhmc_set_timeline(self._h, None, 0))
      return None
    grid = self.plan_info()["grid_ctas"]
    self._timeline = torch.zeros(n_passes, grid, 32, dtype=torch.int64, device=self.dev)
    _C.check(self.lib.edhmc_set_timeline(self._h, self._timeline.data_ptr(), int(n_passes)))
    return self._timeline

  def clear_trace(self):
    self._trace = None
    fn = self.lib.edhmc_set_trace_f64 if self.dtype == torch.float64 else self.lib.edhmc_set_trace
    _C.check(fn(self._h, None, None))

  def read_state(self):
    n = C.c_int64(0)
    lp = C.c_double(0.0)
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_read_state(self._h, C.byref(n), C.byref(lp), _stream_ptr(self.dev)))
    return int(n.value), float(lp.value)

  def reset(self):
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_reset(self._h, _stream_ptr(self.dev)))

  def set_prior_kinds(self, kinds):
    """Per-latent prior family (edhmc_set_prior_kinds): 0 Normal(loc, scale), 1 Beta(a = loc, b = scale) through a sigmoid."""
    k = np.ascontiguousarray(kinds, np.int32).reshape(self.P)
    with torch.cuda.device(self.dev):
      _C.check(self.lib.edhmc_set_prior_kinds(self._h, k.ctypes.data_as(C.POINTER(C.c_int32))))

  def seed(self, seed: int):
    _C.check(self.lib.edhmc_seed(self._h, C.c_uint64(int(seed) & (2**64 - 1))))

  def plan_info(self) -> dict:
    if self.dtype == torch.float64:
      return {"plan_in_use": _C.PLAN_STEPWISE, "dtype": "f64"}
    out = (C.c_int64 * 13)()
    n = _C.check(self.lib.edhmc_plan_info(self._h, out, 13))
    keys = ["grid_ctas", "warps_per_cta", "ring_stages", "tile_rows", "lanes_per_row", "vec_width", "smem_bytes",
            "plan_in_use", "passes_last_run", "launches_last_run", "ring_mode", "smem_resident_tiles_per_cta",
            "tmem_resident_tiles_per_cta"]
    return {k: int(out[i]) for i, k in enumerate(keys[:n])}

  def close(self):
    if getattr(self, "_h", None) is not None and self._h.value:
      self.lib.edhmc_destroy(self._h)
      self._h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass
