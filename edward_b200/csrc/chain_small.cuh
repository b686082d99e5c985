// chain_small.cuh — single-CTA kernels of the stepwise plan and bind-time utilities. Included by edhmc.cu only.
#pragma once
#include "chain.cuh"

namespace edhmc {

// ---- single-CTA chain kernels (<<<1, kChainThreads>>>) of the stepwise plan, on the global chain state ----

__device__ __forceinline__ double finish_gradient(const KArgs& a, const float* pos, float* gout, double* red) {
  double pl = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    const float loc = a.prior_loc[c], sc = a.prior_scale[c];
    const int kind = a.prior_kind ? a.prior_kind[c] : 0;
    gout[c] = static_cast<float>(a.sums[c] + prior_grad_kind(kind, pos[c], loc, sc, prior_inv_var(sc)));
    pl += prior_logp_kind(kind, pos[c], loc, sc);
  }
  return (block_sum_f64(pl, red) - a.prior_const) + a.sums[a.P];
}

// Decides whether the cached (gcur, logp_cur) still describe params[max(t0-1,0)].
__global__ void __launch_bounds__(kChainThreads, 1) k_chain_check(const KArgs a) {
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  bool mismatch = false;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    const float v = a.params[t_prev * a.ldp + c];
    if (__float_as_uint(v) != __float_as_uint(a.zcur[c])) mismatch = true;
  }
  const int need = __syncthreads_or((mismatch || !a.sc->valid) ? 1 : 0);
  if (need)
    for (int c = threadIdx.x; c < a.P; c += kChainThreads) a.zcur[c] = a.params[t_prev * a.ldp + c];
  if (threadIdx.x == 0) a.sc->need_init = need ? 1 : 0;
}

__global__ void __launch_bounds__(kChainThreads, 1) k_chain_init_finish(const KArgs a) {
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
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) ks += static_cast<double>(__fmul_rn(a.r[c], a.r[c]));
  const double k_new = 0.5 * block_sum_f64(ks, red);
  const double logp_cur = a.sc->logp_cur;
  const double k_old = a.sc->k_old;
  const AcceptResult ar = mh_accept(k_old, k_new, logp_new, logp_cur, a.sc->log_u);
  __syncthreads();
  if (threadIdx.x == 0) {
    write_trace(a, it, logp_cur, logp_new, k_old, k_new, ar);
    if (!isfinite(logp_new)) a.sc->nonfinite = 1;
    if (ar.accept) {
      a.sc->logp_cur = logp_new;
      a.sc->n_accept += 1;
    }
  }
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    if (a.trace_pos) a.trace_pos[it * a.P + c] = a.z[c];
    if (ar.accept) {
      a.zcur[c] = a.z[c];
      a.gcur[c] = gnew[c];
    }
    a.params[t * a.ldp + c] = ar.accept ? a.z[c] : a.zcur[c];
  }
}

// Start of transition `it`: draw momentum and uniform, kinetic energy, first half kick + drift.
__global__ void __launch_bounds__(kChainThreads, 1) k_chain_begin(const KArgs a, long long it, float* gwork) {
  __shared__ double red[64];
  const long long t = a.t0 + it;
  double ks = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
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
__global__ void __launch_bounds__(kChainThreads, 1) k_chain_leap(const KArgs a, long long it, int s, float* gwork) {
  __shared__ double red[64];
  const double logp_new = finish_gradient(a, a.z, gwork, red);
  const bool last = (s == a.L - 1);
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
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

// One SGLD / SGHMC update after the pass at params[max(t-1,0)] (sgld.py:52-87, sghmc.py:58-96); float32 in the
// op order of the reference graph.
struct SgArgs {
  int kind;
  float step_size, friction, lik_factor;
  const float* prior_factor;
  float* velocity;
  const float* noise;
};
__global__ void __launch_bounds__(kChainThreads, 1) k_sg_update(const KArgs a, const SgArgs g, long long it) {
  const long long t = a.t0 + it;
  const long long t_prev = t > 0 ? t - 1 : 0;
  float lr, sd;
  if (g.kind == 0) {
    lr = __fdiv_rn(g.step_size, powf(static_cast<float>(t + 1), 0.55f));
    sd = sqrtf(lr);
  } else {
    lr = __fmul_rn(g.step_size, 0.01f);
    sd = sqrtf(__fmul_rn(lr, g.friction));
  }
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    const float old = a.params[t_prev * a.ldp + c];
    const double pf = g.prior_factor ? static_cast<double>(g.prior_factor[c]) : 1.0;
    const float grad = static_cast<float>(static_cast<double>(g.lik_factor) * a.sums[c] +
                                          pf * prior_grad_kind(a.prior_kind ? a.prior_kind[c] : 0, old, a.prior_loc[c], a.prior_scale[c],
                                                               prior_inv_var(a.prior_scale[c])));
    const float nz = g.noise ? g.noise[it * a.P + c] : philox_normal(a.seed, t, c);
    float out;
    if (g.kind == 0) {
      out = __fadd_rn(__fadd_rn(old, __fmul_rn(__fmul_rn(0.5f, lr), grad)), __fmul_rn(sd, nz));
    } else {
      const float v = g.velocity[c];
      out = __fadd_rn(old, v);
      g.velocity[c] = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(1.0f, __fmul_rn(0.5f, g.friction)), v), __fmul_rn(lr, grad)),
                                __fmul_rn(sd, nz));
    }
    a.params[t * a.ldp + c] = out;
  }
  if (threadIdx.x == 0) {
    a.sc->n_accept += 1;
    a.sc->valid = 0;  // the HMC cache does not describe this row
  }
}

// edhmc_logp_grad epilogue: prior + all-reduced sums → caller's buffers.
__global__ void __launch_bounds__(kChainThreads, 1) k_logp_grad_finish(const KArgs a, const float* theta, double* logp_out,
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
