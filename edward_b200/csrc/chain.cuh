// chain.cuh — the HMC transition (hmc.py:61-130) and leapfrog integrator (hmc.py:195-210) on the device.
//
// Persistent plan: k_hmc_persistent runs ALL transitions of an edhmc_run in one cooperative launch.
// Every CTA carries a redundant copy of the O(P) chain state in shared memory and performs the same
// integrator / accept arithmetic on the same all-CTA totals, so the only grid-wide communication per
// leapfrog step is one barrier + one read of the per-CTA partial sums.
//
// Stepwise plan: k_pass (one data pass, last-arriving CTA folds the partials) + single-CTA chain
// kernels, with an NCCL all-reduce of the [grad, logp] sums between them when rows are sharded.
//
// Schedule: one gradient evaluation per leapfrog step. The reference evaluates L+1 gradients and 2
// extra forward passes per transition (hmc.py:199,206,104-105); the values it recomputes — gradient
// and log joint of the transition's start state — are cached here from the previous transition.
#pragma once
#include "stream.cuh"

namespace edhmc {

// leapfrog pieces, float32 with separate multiply and add exactly as the TF graph has them
// (hmc.py:203-204,208: r + 0.5*step_size*grad ; z + step_size*r).
__device__ __forceinline__ float kick(float r, float half_eps, float g) { return __fadd_rn(r, __fmul_rn(half_eps, g)); }
__device__ __forceinline__ float drift(float z, float eps, float r) { return __fadd_rn(z, __fmul_rn(eps, r)); }

struct AcceptResult {
  double ratio, log_u;
  bool accept;
};
// hmc.py:100-109: ratio = K(r0) - K(rL) + logp(zL) - logp(z0) accumulated in that order; accept = log(u) < ratio.
__device__ __forceinline__ AcceptResult mh_accept(double k_old, double k_new, double logp_new, double logp_old, float u) {
  AcceptResult a;
  a.ratio = ((k_old - k_new) + logp_new) - logp_old;
  a.log_u = static_cast<double>(logf(u));
  a.accept = a.log_u < a.ratio;
  return a;
}

__device__ __forceinline__ void write_trace(const KArgs& a, long long it, double logp_old, double logp_new, double k_old,
                                            double k_new, const AcceptResult& ar) {
  if (a.trace_scalars) {
    double* ts = a.trace_scalars + it * 8;
    ts[0] = logp_old;
    ts[1] = logp_new;
    ts[2] = k_old;
    ts[3] = k_new;
    ts[4] = ar.ratio;
    ts[5] = ar.log_u;
    ts[6] = ar.accept ? 1.0 : 0.0;
    ts[7] = 0.0;
  }
}

// =============================================================================================
// Persistent plan
// =============================================================================================
template <int G, int V, int KMAX>
__global__ void __launch_bounds__(kThreads, 1) k_hmc_persistent(const KArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemLayout sm = carve_smem(smem_raw, a, true);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, D = a.D;
  const int ncta = gridDim.x;
  const size_t ppad = align_up(static_cast<size_t>(P) * 4, 16) / 4;
  float* z = sm.state;
  float* r = z + ppad;
  float* g = r + ppad;
  float* zc = g + ppad;
  float* gc = zc + ppad;

  smem_setup(sm, a);
  const uint64_t policy = a.l2_hint ? l2_policy_evict_last() : 0ull;
  const WarpRows wr = warp_rows(a, blockIdx.x * kWarpsPerCta + warp, ncta * kWarpsPerCta);
  Ring ring;
  ring_init(ring, sm.ring + static_cast<size_t>(warp) * a.S * a.stage_floats, sm.bars + warp * kMaxStages);

  // ---- current state: row max(t0-1,0) of the Empirical store (hmc.py:81-85) + cached (grad, logp) ----
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  bool mismatch = false;
  for (int c = tid; c < P; c += kThreads) {
    const float v = a.params[t_prev * a.ldp + c];
    zc[c] = v;
    gc[c] = a.gcur[c];
    if (__float_as_uint(v) != __float_as_uint(a.zcur[c])) mismatch = true;
  }
  const int valid = a.sc->valid;
  double logp_cur = a.sc->logp_cur;
  const bool need_init = __syncthreads_or((mismatch || !valid) ? 1 : 0) != 0;

  const long long n_passes = (need_init ? 1 : 0) + a.n_iter * a.L;
  ring_prologue(a, wr, ring, n_passes, lane, policy);

  int bufsel = 0;
  unsigned long long epoch = 0;
  double logp_new = logp_cur;

  // Evaluates log joint and gradient at theta = `pos` (shared memory, [P]); writes the gradient to `gout`.
  auto evaluate = [&](const float* pos, float* gout) {
    for (int c = tid; c < D; c += kThreads) sm.theta_s[c] = pos[c];
    __syncthreads();
    const float bias = a.has_bias ? pos[D] : 0.0f;
    stream_pass<G, V, KMAX>(a, wr, ring, sm.theta_s, bias, policy, sm.cta_acc);
    if (ncta > 1) {
      double* mine = a.partials + (static_cast<size_t>(bufsel) * ncta + blockIdx.x) * (P + 1);
      for (int c = tid; c <= P; c += kThreads) mine[c] = sm.cta_acc[c];
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        red_release_add_u64(a.bar, 1ull);
        const unsigned long long target = (epoch + 1) * static_cast<unsigned long long>(ncta);
        while (ld_acquire_u64(a.bar) < target) {
        }
      }
      __syncthreads();
      reduce_partials(a.partials + static_cast<size_t>(bufsel) * ncta * (P + 1), ncta, P, sm.cta_acc, sm.comb);
      bufsel ^= 1;
      ++epoch;
    }
    double pl = 0.0;
    for (int c = tid; c < P; c += kThreads) {
      const float loc = a.prior_loc[c], sc = a.prior_scale[c];
      gout[c] = static_cast<float>(sm.cta_acc[c] + prior_grad(pos[c], loc, sc));
      pl += prior_quad(pos[c], loc, sc);
    }
    const double lik = sm.cta_acc[P];
    logp_new = (block_sum_f64(pl, sm.red) - a.prior_const) + lik;
  };

  if (need_init) {
    evaluate(zc, gc);
    logp_cur = logp_new;
  }

  long long n_acc = 0;
  int nonfinite = 0;
  for (long long it = 0; it < a.n_iter; ++it) {
    const long long t = a.t0 + it;
    // momentum r ~ N(0, I) (hmc.py:88-91) or injected
    double ks = 0.0;
    for (int c = tid; c < P; c += kThreads) {
      const float rv = a.r0 ? a.r0[it * P + c] : philox_normal(a.seed, t, c);
      r[c] = rv;
      z[c] = zc[c];
      g[c] = gc[c];
      ks += static_cast<double>(__fmul_rn(rv, rv));
    }
    const double k_old = 0.5 * block_sum_f64(ks, sm.red);
    const float u = a.u ? a.u[it] : philox_uniform(a.seed, t);
    logp_new = logp_cur;
    for (int s = 0; s < a.L; ++s) {
      for (int c = tid; c < P; c += kThreads) {
        const float rn = kick(r[c], a.half_eps, g[c]);
        r[c] = rn;
        z[c] = drift(z[c], a.eps, rn);
      }
      __syncthreads();
      evaluate(z, g);
      for (int c = tid; c < P; c += kThreads) r[c] = kick(r[c], a.half_eps, g[c]);
    }
    ks = 0.0;
    for (int c = tid; c < P; c += kThreads) ks += static_cast<double>(__fmul_rn(r[c], r[c]));
    const double k_new = 0.5 * block_sum_f64(ks, sm.red);
    const AcceptResult ar = mh_accept(k_old, k_new, logp_new, logp_cur, u);
    if (!isfinite(logp_new)) nonfinite = 1;
    if (blockIdx.x == 0) {
      if (tid == 0) write_trace(a, it, logp_cur, logp_new, k_old, k_new, ar);
      if (a.trace_pos)
        for (int c = tid; c < P; c += kThreads) a.trace_pos[it * P + c] = z[c];
    }
    if (ar.accept) {
      for (int c = tid; c < P; c += kThreads) {
        zc[c] = z[c];
        gc[c] = g[c];
      }
      logp_cur = logp_new;
      ++n_acc;
    }
    // Empirical write, on device (hmc.py:121-126)
    if (blockIdx.x == 0)
      for (int c = tid; c < P; c += kThreads) a.params[t * a.ldp + c] = zc[c];
  }

  if (blockIdx.x == 0) {
    for (int c = tid; c < P; c += kThreads) {
      a.zcur[c] = zc[c];
      a.gcur[c] = gc[c];
    }
    if (tid == 0) {
      a.sc->logp_cur = logp_cur;
      a.sc->valid = 1;
      a.sc->n_accept += n_acc;
      if (nonfinite) a.sc->nonfinite = 1;
    }
  }
}

// =============================================================================================
// Stepwise plan
// =============================================================================================
// One data pass at theta (global, [P]). Shard sums land in a.sums[0..P] (float64), written by the
// last CTA to finish (fixed summation order → reproducible). gate != 0: skip unless sc->need_init.
template <int G, int V, int KMAX>
__global__ void __launch_bounds__(kThreads, 1) k_pass(const KArgs a, const float* theta, int gate) {
  if (gate && !a.sc->need_init) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemLayout sm = carve_smem(smem_raw, a, false);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, D = a.D, ncta = gridDim.x;
  __shared__ int s_last;

  smem_setup(sm, a);
  const uint64_t policy = a.l2_hint ? l2_policy_evict_last() : 0ull;
  const WarpRows wr = warp_rows(a, blockIdx.x * kWarpsPerCta + warp, ncta * kWarpsPerCta);
  Ring ring;
  ring_init(ring, sm.ring + static_cast<size_t>(warp) * a.S * a.stage_floats, sm.bars + warp * kMaxStages);
  ring_prologue(a, wr, ring, 1, lane, policy);

  for (int c = tid; c < D; c += kThreads) sm.theta_s[c] = theta[c];
  __syncthreads();
  const float bias = a.has_bias ? theta[D] : 0.0f;
  stream_pass<G, V, KMAX>(a, wr, ring, sm.theta_s, bias, policy, sm.cta_acc);

  double* mine = a.partials + static_cast<size_t>(blockIdx.x) * (P + 1);
  for (int c = tid; c <= P; c += kThreads) mine[c] = sm.cta_acc[c];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(a.ticket, 1u);
    s_last = (ticket == static_cast<unsigned int>(ncta - 1));
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    reduce_partials(a.partials, ncta, P, sm.cta_acc, sm.comb);
    for (int c = tid; c <= P; c += kThreads) a.sums[c] = sm.cta_acc[c];
    if (tid == 0) *a.ticket = 0u;
  }
}

// ---- single-CTA chain kernels (<<<1, kThreads>>>) operating on the global chain state ----

__device__ __forceinline__ double finish_gradient(const KArgs& a, const float* pos, float* gout, double* red) {
  double pl = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kThreads) {
    const float loc = a.prior_loc[c], sc = a.prior_scale[c];
    gout[c] = static_cast<float>(a.sums[c] + prior_grad(pos[c], loc, sc));
    pl += prior_quad(pos[c], loc, sc);
  }
  return (block_sum_f64(pl, red) - a.prior_const) + a.sums[a.P];
}

// Decides whether the cached (gcur, logp_cur) still describe params[max(t0-1,0)].
__global__ void __launch_bounds__(kThreads, 1) k_chain_check(const KArgs a) {
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  bool mismatch = false;
  for (int c = threadIdx.x; c < a.P; c += kThreads) {
    const float v = a.params[t_prev * a.ldp + c];
    if (__float_as_uint(v) != __float_as_uint(a.zcur[c])) mismatch = true;
  }
  const int need = __syncthreads_or((mismatch || !a.sc->valid) ? 1 : 0);
  if (need)
    for (int c = threadIdx.x; c < a.P; c += kThreads) a.zcur[c] = a.params[t_prev * a.ldp + c];
  if (threadIdx.x == 0) a.sc->need_init = need ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads, 1) k_chain_init_finish(const KArgs a) {
  __shared__ double red[64];
  if (!a.sc->need_init) return;
  const double lp = finish_gradient(a, a.zcur, a.gcur, red);
  if (threadIdx.x == 0) {
    a.sc->logp_cur = lp;
    a.sc->valid = 1;
    a.sc->need_init = 0;
  }
}

__device__ __forceinline__ void chain_finish(const KArgs& a, long long it, float* gnew, double logp_new, double* red) {
  const long long t = a.t0 + it;
  double ks = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kThreads) ks += static_cast<double>(__fmul_rn(a.r[c], a.r[c]));
  const double k_new = 0.5 * block_sum_f64(ks, red);
  const double logp_cur = a.sc->logp_cur;
  const double k_old = a.sc->k_old;
  AcceptResult ar;
  ar.ratio = ((k_old - k_new) + logp_new) - logp_cur;
  ar.log_u = a.sc->log_u;
  ar.accept = ar.log_u < ar.ratio;
  __syncthreads();
  if (threadIdx.x == 0) {
    write_trace(a, it, logp_cur, logp_new, k_old, k_new, ar);
    if (!isfinite(logp_new)) a.sc->nonfinite = 1;
    if (ar.accept) {
      a.sc->logp_cur = logp_new;
      a.sc->n_accept += 1;
    }
  }
  for (int c = threadIdx.x; c < a.P; c += kThreads) {
    if (a.trace_pos) a.trace_pos[it * a.P + c] = a.z[c];
    if (ar.accept) {
      a.zcur[c] = a.z[c];
      a.gcur[c] = gnew[c];
    }
    a.params[t * a.ldp + c] = ar.accept ? a.z[c] : a.zcur[c];
  }
}

// Start of transition `it`: draw momentum and uniform, kinetic energy, first half kick + drift.
__global__ void __launch_bounds__(kThreads, 1) k_chain_begin(const KArgs a, long long it, float* gwork) {
  __shared__ double red[64];
  const long long t = a.t0 + it;
  double ks = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kThreads) {
    const float rv = a.r0 ? a.r0[it * a.P + c] : philox_normal(a.seed, t, c);
    ks += static_cast<double>(__fmul_rn(rv, rv));
    float zz = a.zcur[c];
    float rr = rv;
    if (a.L > 0) {
      rr = kick(rv, a.half_eps, a.gcur[c]);
      zz = drift(zz, a.eps, rr);
    }
    a.r[c] = rr;
    a.z[c] = zz;
    gwork[c] = a.gcur[c];
  }
  const double k_old = 0.5 * block_sum_f64(ks, red);
  if (threadIdx.x == 0) {
    const float u = a.u ? a.u[it] : philox_uniform(a.seed, t);
    a.sc->k_old = k_old;
    a.sc->log_u = static_cast<double>(logf(u));
  }
  if (a.L == 0) {
    __syncthreads();
    chain_finish(a, it, gwork, a.sc->logp_cur, red);
  }
}

// After the pass (and all-reduce) of leapfrog step s: second half kick; then either the next step's
// first half kick + drift, or the Metropolis–Hastings accept and the Empirical write.
__global__ void __launch_bounds__(kThreads, 1) k_chain_leap(const KArgs a, long long it, int s, float* gwork) {
  __shared__ double red[64];
  const double logp_new = finish_gradient(a, a.z, gwork, red);
  const bool last = (s == a.L - 1);
  for (int c = threadIdx.x; c < a.P; c += kThreads) {
    float rr = kick(a.r[c], a.half_eps, gwork[c]);
    if (!last) {
      rr = kick(rr, a.half_eps, gwork[c]);
      a.z[c] = drift(a.z[c], a.eps, rr);
    }
    a.r[c] = rr;
  }
  if (last) {
    __syncthreads();
    chain_finish(a, it, gwork, logp_new, red);
  }
}

// edhmc_logp_grad epilogue: prior + all-reduced sums → caller's buffers.
__global__ void __launch_bounds__(kThreads, 1) k_logp_grad_finish(const KArgs a, const float* theta, double* logp_out,
                                                                 float* grad_out) {
  __shared__ double red[64];
  const double lp = finish_gradient(a, theta, grad_out, red);
  if (threadIdx.x == 0) *logp_out = lp;
}

// ---- bind-time scan: counts NaN/Inf in X (D valid columns per row) and y ----
__global__ void k_check_finite(const float* X, long long n_rows, long long ldx, int D, const void* y, int y_dtype,
                               unsigned long long* bad) {
  unsigned long long local = 0;
  const long long total = n_rows * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / D;
    const int col = static_cast<int>(i - row * D);
    if (!isfinite(X[row * ldx + col])) ++local;
  }
  if (y_dtype == 1) {
    const float* yf = reinterpret_cast<const float*>(y);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_rows;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      if (!isfinite(yf[i])) ++local;
  }
  if (local) atomicAdd(bad, local);
}

__global__ void k_u8_to_i32(const unsigned char* src, int* dst, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

}  // namespace edhmc
