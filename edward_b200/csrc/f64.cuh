// f64.cuh — float64 models (the reference runs its HMC tests in float32 AND float64, tests/inferences/hmc_test.py:93-97).
//
// The hot path of this library is float32 (the dtype of every BASELINE configuration). A float64 model takes this
// compact path instead: one generic data-pass kernel (a warp per row, lanes over the features, float64 arithmetic,
// fixed-order reductions: per-warp shared-memory slices, per-CTA partials, one reduce kernel) and single-CTA chain kernels
// that mirror chain_small.cuh in double. Same schedule as the stepwise float32 plan (one gradient evaluation per leapfrog step, the start state's
// gradient and log joint cached across transitions), same draws (Philox, widened to double), all on the device.
// It is correct and device-resident, not tuned: float64 models in the reference are its unit tests (N <= 50).
#pragma once
#include "common.cuh"

namespace edhmc {

struct K64Args {
  const double* X;
  const void* y;  // int32 or float64
  long long n_rows, ldx;
  int D, P, has_bias, family, y_is_f64;
  double lik_scale;
  const float* prior_loc;
  const float* prior_scale;
  const int* prior_kind;
  double prior_const;
  double* sums;  // [P+1]
  ChainScalars* sc;
  double* zcur;  // [P]
  double* gcur;
  double* z;
  double* r;
  double* g;
  double* params;
  long long ldp, t0, n_iter;
  double eps;
  int L;
  const double* r0;
  const double* u;
  unsigned long long seed;
  double* trace_scalars;
  double* trace_pos;
};

__device__ __forceinline__ void row_terms64(int family, double eta, double yv, double s, double& lp, double& r) {
  if (family == 0) {
    const double e = exp(-fabs(eta));
    lp = -((eta >= 0.0 ? eta : 0.0) - eta * yv + log1p(e));
    const double q = e / (1.0 + e);
    r = eta >= 0.0 ? (yv - 1.0) + q : yv - q;
  } else if (family == 1) {
    const double zz = (yv - eta) / s;
    lp = -0.5 * zz * zz - (0.9189385332046727 + log(s));
    r = zz / s;
  } else {
    const double mu = exp(eta);
    lp = yv * eta - mu - lgamma(yv + 1.0);
    r = yv - mu;
  }
}

// Data pass, deterministic (no atomics): warp w of a CTA owns rows n = (cta*nw + w) + k*grid*nw and accumulates
// r_n X[n,:] (lane l owns features l, l+32, ...), r_n (bias) and log p(y_n | eta_n) into its own shared-memory slice in
// row order; the CTA then sums its warps in warp order into partials[cta, 0..P], and k64_reduce sums the CTAs in CTA order
// into sums[0..P] (sums[P] = log-likelihood). A chain is therefore bitwise reproducible and independent of how run() calls
// are chunked, like the float32 path.
__global__ void __launch_bounds__(256) k64_pass(const K64Args a, const double* theta, double* partials, int gate) {
  if (gate && !a.sc->need_init) return;
  extern __shared__ double sacc[];  // [nw][P+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int W = a.P + 1;
  double* mine = sacc + static_cast<size_t>(warp) * W;
  for (int c = lane; c < W; c += 32) mine[c] = 0.0;
  __syncwarp();
  const double bias = a.has_bias ? theta[a.D] : 0.0;
  double lp_acc = 0.0, gb = 0.0;
  for (long long n = static_cast<long long>(blockIdx.x) * nw + warp; n < a.n_rows; n += static_cast<long long>(gridDim.x) * nw) {
    const double* row = a.X + n * a.ldx;
    double dot = 0.0;
    for (int d = lane; d < a.D; d += 32) dot += row[d] * theta[d];
    dot = warp_sum_f64(dot);
    const double yv = a.y_is_f64 ? reinterpret_cast<const double*>(a.y)[n] : static_cast<double>(reinterpret_cast<const int*>(a.y)[n]);
    double lp, r;
    row_terms64(a.family, dot + bias, yv, a.lik_scale, lp, r);
    for (int d = lane; d < a.D; d += 32) mine[d] += r * row[d];
    lp_acc += lp;
    gb += r;
  }
  if (lane == 0) {
    mine[a.P] = lp_acc;
    if (a.has_bias) mine[a.D] = gb;
  }
  __syncthreads();
  for (int c = tid; c < W; c += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += sacc[static_cast<size_t>(w) * W + c];
    partials[static_cast<size_t>(blockIdx.x) * W + c] = t;
  }
}

__global__ void __launch_bounds__(256) k64_reduce(const double* partials, int n_ctas, int W, double* sums, const ChainScalars* sc,
                                                  int gate) {
  if (gate && !sc->need_init) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= W) return;
  double t = 0.0;
  for (int b = 0; b < n_ctas; ++b) t += partials[static_cast<size_t>(b) * W + c];
  sums[c] = t;
}

__device__ __forceinline__ double prior_logp64(int kind, double z, double p0, double p1) {
  if (kind == 1) return p0 * log_sigmoid64(z) + p1 * log_sigmoid64(-z);
  const double t = (z - p0) / p1;
  return -0.5 * t * t;
}
__device__ __forceinline__ double prior_grad64(int kind, double z, double p0, double p1) {
  if (kind == 1) return p0 - (p0 + p1) / (1.0 + exp(-z));
  return -((z - p0) / p1) / p1;
}

__device__ __forceinline__ double finish_gradient64(const K64Args& a, const double* pos, double* gout, double* red) {
  double pl = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    const double p0 = a.prior_loc[c], p1 = a.prior_scale[c];
    const int kind = a.prior_kind ? a.prior_kind[c] : 0;
    gout[c] = a.sums[c] + prior_grad64(kind, pos[c], p0, p1);
    pl += prior_logp64(kind, pos[c], p0, p1);
  }
  return (block_sum_f64(pl, red) - a.prior_const) + a.sums[a.P];
}

__global__ void __launch_bounds__(kChainThreads, 1) k64_check(const K64Args a) {
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  bool mismatch = false;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads)
    if (a.params[t_prev * a.ldp + c] != a.zcur[c]) mismatch = true;
  const int need = __syncthreads_or((mismatch || !a.sc->valid) ? 1 : 0);
  if (need)
    for (int c = threadIdx.x; c < a.P; c += kChainThreads) a.zcur[c] = a.params[t_prev * a.ldp + c];
  if (threadIdx.x == 0) a.sc->need_init = need ? 1 : 0;
}

__global__ void __launch_bounds__(kChainThreads, 1) k64_init_finish(const K64Args a) {
  __shared__ double red[64];
  if (!a.sc->need_init) return;
  const double lp = finish_gradient64(a, a.zcur, a.gcur, red);
  if (threadIdx.x == 0) {
    a.sc->logp_cur = lp;
    a.sc->valid = 1;
    a.sc->need_init = 0;
  }
}

__device__ __forceinline__ void chain_finish64(const K64Args& a, long long it, double logp_new, double* red) {
  const long long t = a.t0 + it;
  double ks = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) ks += a.r[c] * a.r[c];
  const double k_new = 0.5 * block_sum_f64(ks, red);
  const double logp_cur = a.sc->logp_cur, k_old = a.sc->k_old, log_u = a.sc->log_u;
  const double ratio = ((k_old - k_new) + logp_new) - logp_cur;  // hmc.py:100-105
  const bool accept = log_u < ratio;                               // hmc.py:108-109, strict
  __syncthreads();
  if (threadIdx.x == 0) {
    if (a.trace_scalars) {
      double* ts = a.trace_scalars + it * 8;
      ts[0] = logp_cur; ts[1] = logp_new; ts[2] = k_old; ts[3] = k_new; ts[4] = ratio; ts[5] = log_u; ts[6] = accept ? 1.0 : 0.0; ts[7] = 0.0;
    }
    if (!isfinite(logp_new)) a.sc->nonfinite = 1;
    if (accept) {
      a.sc->logp_cur = logp_new;
      a.sc->n_accept += 1;
    }
  }
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    if (a.trace_pos) a.trace_pos[it * a.P + c] = a.z[c];
    if (accept) {
      a.zcur[c] = a.z[c];
      a.gcur[c] = a.g[c];
    }
    a.params[t * a.ldp + c] = accept ? a.z[c] : a.zcur[c];
  }
}

__global__ void __launch_bounds__(kChainThreads, 1) k64_begin(const K64Args a, long long it) {
  __shared__ double red[64];
  const long long t = a.t0 + it;
  double ks = 0.0;
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    const double rv = a.r0 ? a.r0[it * a.P + c] : static_cast<double>(philox_normal(a.seed, t, c));
    ks += rv * rv;
    double zz = a.zcur[c], rr = rv;
    if (a.L > 0) {
      rr = rv + 0.5 * a.eps * a.gcur[c];
      zz = zz + a.eps * rr;
    }
    a.r[c] = rr;
    a.z[c] = zz;
    a.g[c] = a.gcur[c];
  }
  const double k_old = 0.5 * block_sum_f64(ks, red);
  if (threadIdx.x == 0) {
    const double u = a.u ? a.u[it] : static_cast<double>(philox_uniform(a.seed, t));
    a.sc->k_old = k_old;
    a.sc->log_u = log(u);
  }
  if (a.L == 0) {
    __syncthreads();
    chain_finish64(a, it, a.sc->logp_cur, red);
  }
}

__global__ void __launch_bounds__(kChainThreads, 1) k64_leap(const K64Args a, long long it, int s) {
  __shared__ double red[64];
  const double logp_new = finish_gradient64(a, a.z, a.g, red);
  const bool last = (s == a.L - 1);
  for (int c = threadIdx.x; c < a.P; c += kChainThreads) {
    double rr = a.r[c] + 0.5 * a.eps * a.g[c];
    if (!last) {
      rr = rr + 0.5 * a.eps * a.g[c];
      a.z[c] = a.z[c] + a.eps * rr;
    }
    a.r[c] = rr;
  }
  if (last) {
    __syncthreads();
    chain_finish64(a, it, logp_new, red);
  }
}

__global__ void __launch_bounds__(kChainThreads, 1) k64_logp_grad_finish(const K64Args a, const double* theta, double* logp_out,
                                                                   double* grad_out) {
  __shared__ double red[64];
  const double lp = finish_gradient64(a, theta, grad_out, red);
  if (threadIdx.x == 0) *logp_out = lp;
}

__global__ void k64_check_finite(const double* X, long long n_rows, long long ldx, int D, const void* y, int y_is_f64,
                                 unsigned long long* bad) {
  unsigned long long local = 0;
  const long long total = n_rows * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / D;
    if (!isfinite(X[row * ldx + (i - row * D)])) ++local;
  }
  if (y_is_f64)
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_rows;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      if (!isfinite(reinterpret_cast<const double*>(y)[i])) ++local;
  if (local) atomicAdd(bad, local);
}

}  // namespace edhmc
