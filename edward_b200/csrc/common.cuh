// common.cuh — shared device-side definitions: kernel argument block, likelihood families,
// Philox4x32-10, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace edhmc {

constexpr int kMaxWarps = 16;   // consumer warps per CTA (each owns a private TMA ring): 8, 12 or 16
constexpr int kChainThreads = 256;  // block size of the single-CTA chain kernels (stepwise plan)
constexpr int kMaxStages = 8;
constexpr int kXwFloats = 2048;  // per-CTA scratch for the one-shot cross-warp reduction (8 KB)
constexpr int kMaxFeatures = 2048;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kTlRec = 32;  // int64 slots per (pass, CTA) record of the development timeline

// Inbox of the in-kernel all-reduce (row shards over the GPUs of one NVLink domain; also the second level of
// the wide single-GPU reduction). One cudaMalloc per rank, exported with cudaIpc: flags[2][kMaxRanks] (u64) at
// byte 0, slice counters[2][kMaxRanks] (u64) at byte kInboxCountOff, data[2][kMaxRanks][kInboxStride] (f64) at
// byte kInboxDataOff. Slot = parity of the pass sequence number; entry r is written by rank r.
constexpr int kMaxRanks = 8;
constexpr int kInboxStride = kMaxFeatures + 8;
constexpr int kInboxCountOff = 128;
constexpr int kInboxDataOff = 256;
constexpr int kLlGroup = 16;     // CTAs per first-level group of the narrow-model reduction (leader protocol and reduce_partials)
constexpr int kLlMaxGroups = 16;  // so grids of up to 256 CTAs
constexpr int kLlCopies = 8;      // replicas of the broadcast position: CTA i polls copy i % 8 (spreads the polling over L2 slices)
constexpr int kWideCols = 256;    // more totals than this: two-level reduction (column slices) instead of all-read-all
constexpr int kWideSlices = 128;  // column slices of the two-level reduction (independent of the grid size)
constexpr size_t kInboxBytes = kInboxDataOff + sizeof(double) * 2 * kMaxRanks * kInboxStride;

// Default warps per CTA as a function of the floats each lane keeps of a row (x, theta and gradient slices
// live in registers, ~3*KV + 40): 16 warps under 128 registers/thread, 12 under 168, else 8.
__host__ __device__ constexpr int warps_for(int kv) { return 3 * kv + 40 <= 128 ? 16 : (3 * kv + 40 <= 168 ? 12 : 8); }

// Scalars of the chain that live in global memory between launches.
struct ChainScalars {
  double logp_cur;   // log joint of the current state (cached across transitions and launches)
  double logp_new;   // stepwise plan: log joint at the proposal
  double k_old;      // stepwise plan: kinetic energy of the drawn momentum
  double log_u;      // stepwise plan
  long long n_accept;
  int valid;         // 1: (zcur, gcur, logp_cur) describe params[max(t-1,0)]
  int need_init;     // stepwise plan: the initial evaluation pass is live
  int nonfinite;     // debug: set if a log joint / gradient went NaN/Inf
  int pad;
};

struct KArgs {
  // ---- problem (edhmc_cfg + bound data) ----
  const float* X;
  const void* y;
  long long n_rows;
  long long ldx;
  int D;
  int P;
  int has_bias;
  int family;
  int y_dtype;
  float lik_scale;
  const float* prior_loc;
  const float* prior_scale;
  const int* prior_kind;  // [P] or nullptr (all Normal): 0 Normal(loc, scale), 1 Beta(a = loc, b = scale) through a sigmoid
  double prior_const;  // sum of the priors' additive constants (Normal: 0.5*log(2*pi) + log(scale); Beta: lbeta), host float64
  // ---- streaming plan ----
  int Kact;          // active vector chunks per lane
  int J;             // row groups per tile
  int RT;            // rows per tile = (32/G)*J
  int S;             // ring stages per warp
  int stage_floats;  // ring stage stride in floats (x tile + y slice + pad)
  int y_off;         // float offset of the y slice inside a stage
  int wpad;          // padded length of theta in shared memory (G*KMAX*V)
  int zigzag;        // 1: odd passes walk a warp's tiles backwards (L2 reuse)
  int l2_hint;       // 0 none, 1 evict_last on all X tiles, 2 evict_last on the fraction l2_frac of X, evict_first on the rest
  float l2_frac;
  int ldx_i;         // ldx as int
  int tl;            // floats per full tile = RT*ldx
  int tm;            // tl & 3: per-tile drift of the 16-byte alignment (non-zero only if ldx % 4 != 0)
  int wpg;           // ring mode 1: warps per group (each group owns a row range and a ring of S stages)
  int interleave;    // 1: tile t belongs to warp t % (grid*NW) (moving window); 0: contiguous row range per warp
  // ---- ring mode 2 (stream_ldg.cuh): X re-laid at bind time into column-pair-major 32-row tiles, read with LDG ----
  const float2* Xt;  // [n_tiles][Kact][32] float2
  const float* Yt;   // [n_tiles][32]
  long long n_tiles;
  int n_res;          // tiles per CTA kept resident in shared memory for the whole launch (persistent plan only)
  int n_tm;           // tiles per CTA that a persistent launch parks in tensor memory (they are the last of the CTA's range
                      // in either plan: the tile order must not depend on the plan)
  int tm_on;          // 1: this launch allocates tensor memory and reads those tiles from there
  // ---- launch mode ----
  int mode;               // 0: run n_iter transitions (cooperative launch); 1: one data pass at theta_in
  int gate;               // mode 1: return immediately unless sc->need_init
  int par0;               // mode 1: parity of this pass's global leapfrog-step index (zig-zag direction)
  int single_lp;          // mode 1: 1 = accumulate the log-likelihood too, 0 = gradient-only pass (as the persistent plan
                          // runs the leapfrog steps inside a trajectory, so that both plans stay bit-identical)
  const float* theta_in;  // mode 1: [P]
  // ---- row shards over several GPUs (persistent plan): one-shot all-reduce through peer memory ----
  int nranks;                        // 1: no exchange
  int rank;
  unsigned char* const* peer_inbox;  // device array [nranks]: inbox base of every rank (own included), peer-mapped
  unsigned long long* comm_seq;      // passes exchanged so far (identical on every rank, never reset)
  int* abort_flag;                   // set when a wait on a peer timed out; every spin loop of the kernel polls it
  long long spin_limit;              // clock64 ticks before a wait on a peer gives up
  // ---- scratch ----
  double* partials;            // [2][grid][P+1]
  unsigned long long* bar;     // grid barrier counter (persistent plan)
  unsigned int* ticket;        // last-arriver ticket (stepwise plan)
  double* sums;                // [P+1] shard sums → all-reduced in place (stepwise plan)
  // ---- leader protocol of the persistent plan (one GPU, P+1 <= kWideCols; chain.cuh) ----
  int leader;                  // 1: CTA 0 owns the chain; the other CTAs only run data passes
  unsigned int ll_seq0;        // sequence number of this launch's pass 0, minus 1 (never repeats within 2^32 passes)
  uint4* ll_part;              // [2][grid][P+1] {lo32, seq, hi32, seq}: per-CTA float64 sums with the flag in the data
  uint4* ll_group;             // [2][ceil(grid / kLlGroup)][P+1]: sums of groups of kLlGroup CTAs, same entry format
  uint2* ll_theta;             // [2][kLlCopies][P] {float bits, seq}: the position of the next pass
  // ---- chain state ----
  ChainScalars* sc;
  float* zcur;  // [P]
  float* gcur;  // [P]
  float* z;     // [P] stepwise plan working position (also theta of edhmc_logp_grad)
  float* r;     // [P] stepwise plan working momentum
  // ---- run ----
  float* params;
  long long ldp;
  long long t0;
  long long n_iter;
  float eps;
  float half_eps;
  int L;
  const float* r0;
  const float* u;
  unsigned long long seed;
  double* trace_scalars;
  float* trace_pos;
  long long* timeline;  // development: [tl_cap][grid][kTlRec] int64 per-pass stamps of the persistent plan, or nullptr
  int tl_cap;
};

// ---------------------------------------------------------------------------------------------
// Likelihood families. Returns log p(y|eta) and d/deta log p(y|eta), float32, in the op order of
// the TensorFlow path the reference executes (see oracle/hmc_oracle.py for the citations).
// ---------------------------------------------------------------------------------------------
// Gradient-only variant for the leapfrog steps inside a trajectory: the log joint enters the transition only at
// its two ends (hmc.py:104-105), so the log1p / lgamma work is skipped there. r is bit-identical to row_terms'.
__device__ __forceinline__ float row_resid(int family, float eta, float yv, float lik_scale) {
  if (family == 0) {
    const bool pos = eta >= 0.0f;
    const float e = expf(-fabsf(eta));
    const float u = __fadd_rn(1.0f, e);
    const float inv = __fdividef(1.0f, u);
    const float q = __fmul_rn(e, inv);
    return pos ? __fadd_rn(__fsub_rn(yv, 1.0f), q) : __fsub_rn(yv, q);
  } else if (family == 1) {
    const float zz = __fdiv_rn(__fsub_rn(yv, eta), lik_scale);
    return __fdiv_rn(zz, lik_scale);
  } else {
    return __fsub_rn(yv, expf(eta));
  }
}

// Bernoulli-logit residual y - sigmoid(eta) for the gradient-only passes inside a trajectory, on the special-function
// unit: e = 2^(-|eta| log2 e) (ex2.approx) and 1/(1+e) (rcp.approx) — 7 dependent instructions instead of the ~25 of
// expf + division. Absolute error <= 4e-7 per row (ex2.approx and rcp.approx are good to ~2 ulp on (0,1] and (1,2]),
// two orders of magnitude inside the 1e-5 gradient tolerance; the log joint that decides acceptance is always
// evaluated with row_terms below.
__device__ __forceinline__ float bernoulli_resid_fast(float eta, float yv) {
  float e, inv;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * fabsf(eta)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.0f + e));
  const float q = e * inv;                       // sigmoid(-|eta|)
  return (eta >= 0.0f) ? (yv - 1.0f) + q : yv - q;  // y - sigmoid(eta)
}

// Shortest form, for the tensor-core many-chain epilogue where the instruction count per element paces the pipeline:
// y - 1/(1 + 2^(-eta*log2(e))). ex2 overflows to +inf for eta < -88 (sigmoid -> 0) and underflows to 0 for eta > 88
// (sigmoid -> 1): no range handling needed. Absolute error of the residual <= 1.2e-7.
__device__ __forceinline__ float bernoulli_resid_direct(float eta, float yv) {
  float e, inv;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * eta));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.0f + e));
  return yv - inv;
}

// Bernoulli log-likelihood term and residual from ex2 / lg2 / rcp (many-chain epilogue, last step of a trajectory):
// lp = -(max(eta, 0) - eta*y + log(1 + e)), e = exp(-|eta|); r = y - sigmoid(eta). log(1 + e) through lg2.approx of the
// rounded 1 + e: absolute error <= 1.2e-7 per term (the terms are O(1) and there are N of them: <= 1e-7 relative on the sum).
__device__ __forceinline__ void bernoulli_terms_fast(float eta, float yv, float& lp, float& r) {
  float e, inv, lg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * fabsf(eta)));
  const float u = 1.0f + e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(u));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));
  const float q = e * inv;  // sigmoid(-|eta|)
  const bool pos = eta >= 0.0f;
  lp = -(((pos ? eta : 0.0f) - eta * yv) + 0.6931471805599453f * lg);
  r = pos ? (yv - 1.0f) + q : yv - q;
}

__device__ __forceinline__ void row_terms(int family, float eta, float yv, float lik_scale, float& lp, float& r) {
  if (family == 0) {
    // -(where(l>=0,l,0) - l*y + log1p(exp(-|l|)));  gradient y - sigmoid(l), piecewise as autodiff does.
    // log1p(e) is evaluated as log(u) - ((u-1)-e)/u with u = 1+e (exact-to-rounding compensation), and
    // e/(1+e) with one reciprocal: u is in (1,2], so no range checks are needed.
    const bool pos = eta >= 0.0f;
    const float e = expf(-fabsf(eta));
    const float u = __fadd_rn(1.0f, e);
    const float inv = __fdividef(1.0f, u);
    const float l1p = __fsub_rn(logf(u), __fmul_rn(__fsub_rn(__fsub_rn(u, 1.0f), e), inv));
    const float relu = pos ? eta : 0.0f;
    lp = -__fadd_rn(__fsub_rn(relu, __fmul_rn(eta, yv)), l1p);
    const float q = __fmul_rn(e, inv);
    r = pos ? __fadd_rn(__fsub_rn(yv, 1.0f), q) : __fsub_rn(yv, q);
  } else if (family == 1) {
    const float zz = __fdiv_rn(__fsub_rn(yv, eta), lik_scale);
    lp = __fsub_rn(__fmul_rn(-0.5f, __fmul_rn(zz, zz)), __fadd_rn(0.9189385332046727f, logf(lik_scale)));
    r = __fdiv_rn(zz, lik_scale);
  } else {
    const float mu = expf(eta);
    lp = __fsub_rn(__fsub_rn(__fmul_rn(yv, eta), mu), lgammaf(yv + 1.0f));
    r = __fsub_rn(yv, mu);
  }
}

__device__ __forceinline__ float load_y(const void* y, int y_dtype, long long i) {
  if (y_dtype == 0) return static_cast<float>(__ldg(reinterpret_cast<const int*>(y) + i));
  if (y_dtype == 1) return __ldg(reinterpret_cast<const float*>(y) + i);
  return static_cast<float>(__ldg(reinterpret_cast<const unsigned char*>(y) + i));
}
__device__ __forceinline__ float y_from_bits(uint32_t bits, int y_dtype) {
  return y_dtype == 0 ? static_cast<float>(static_cast<int>(bits)) : __uint_as_float(bits);
}

// ---------------------------------------------------------------------------------------------
// Priors of the (unconstrained) latents, float64. kind 0: Normal(loc = p0, scale = p1) — the hot path. kind 1: a latent
// with support (0, 1) and a Beta(a = p0, b = p1) prior, sampled in the unconstrained space u = logit(z) as the reference
// does under auto_transform (inference.py:223-264, util/random_variables.py:856-917, hmc.py:132-159): the density
// of u is Beta(sigmoid(u)) * |d sigmoid / du|, so  log p(u) = a log sigmoid(u) + b log sigmoid(-u) - lbeta(a, b)  (the
// log-det-Jacobian log z + log(1-z) is folded in) and  d/du = a - (a + b) sigmoid(u).
// The additive constants (0.5 log 2 pi + log scale; lbeta) are summed once on the host into KArgs::prior_const.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double log_sigmoid64(double u) {  // min(u, 0) - log1p(exp(-|u|))
  return fmin(u, 0.0) - log1p(exp(-fabs(u)));
}
__device__ __forceinline__ double prior_logp_kind(int kind, float zc, float p0, float p1) {
  if (kind == 1) {
    const double u = static_cast<double>(zc);
    return static_cast<double>(p0) * log_sigmoid64(u) + static_cast<double>(p1) * log_sigmoid64(-u);
  }
  const double t = (static_cast<double>(zc) - static_cast<double>(p0)) / static_cast<double>(p1);
  return -0.5 * t * t;
}
// `aux` is the precomputed reciprocal variance for kind 0 (prior_inv_var) and unused otherwise.
__device__ __forceinline__ double prior_grad_kind(int kind, float zc, float p0, float p1, double aux) {
  if (kind == 1) {
    const double u = static_cast<double>(zc);
    const double sg = 1.0 / (1.0 + exp(-u));
    return static_cast<double>(p0) - (static_cast<double>(p0) + static_cast<double>(p1)) * sg;
  }
  return -((static_cast<double>(zc) - static_cast<double>(p0)) * aux);
}

// Normal prior, float64: log density without the constant, and its gradient.
__device__ __forceinline__ double prior_quad(float zc, float loc, float scale) {
  const double t = (static_cast<double>(zc) - static_cast<double>(loc)) / static_cast<double>(scale);
  return -0.5 * t * t;
}
// Same gradient with the reciprocal variance precomputed (one multiply on the critical path between two data passes
// instead of two float64 divisions). The persistent and the stepwise plan both use this form, so they stay bit-identical.
__device__ __forceinline__ double prior_inv_var(float scale) {
  const double s = static_cast<double>(scale);
  return 1.0 / (s * s);
}
__device__ __forceinline__ double prior_grad_iv(float zc, float loc, double inv_var) {
  return -((static_cast<double>(zc) - static_cast<double>(loc)) * inv_var);
}
__device__ __forceinline__ double prior_grad(float zc, float loc, float scale) {
  const double s = static_cast<double>(scale);
  return -((static_cast<double>(zc) - static_cast<double>(loc)) / s) / s;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: draws are a pure function of
// (seed, transition index, element index), so every CTA and every rank generates the same numbers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]);
    const uint32_t lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]);
    const uint32_t lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0;
    const uint32_t n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01_open(uint32_t x) {  // (0,1)
  return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
// Standard normal for element `idx` of transition `t` (Box–Muller on two Philox words).
__device__ __forceinline__ float philox_normal(unsigned long long seed, long long t, int idx) {
  uint32_t c[4] = {static_cast<uint32_t>(idx), static_cast<uint32_t>(t), static_cast<uint32_t>(t >> 32), 0x4e6f726du};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float u1 = u01_open(c[0]);
  const float u2 = u01_open(c[1]);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, long long t) {
  uint32_t c[4] = {0u, static_cast<uint32_t>(t), static_cast<uint32_t>(t >> 32), 0x556e6966u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return u01_open(c[0]);
}

// ---------------------------------------------------------------------------------------------
// Deterministic reductions (fixed tree, float64).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}
// Block-wide sum; every thread gets the result. `scratch` holds >= 32 doubles. Contains barriers.
__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum_f64(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += scratch[w];
  return t;
}

}  // namespace edhmc
