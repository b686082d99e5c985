// chains.cuh — C vectorised HMC chains over one shared design matrix (extension: the reference runs one
// chain per ed.HMC object, hmc.py:14-130). Shared declarations between chains.cu (kernels) and edhmc.cu.
//
// Per leapfrog step the C chains need  S = X·W  (N×D · D×C),  R = y − σ(S)  and  G = Xᵀ·R  (D×C): a dense
// contraction, executed on the tcgen05 tensor cores in 3xTF32 (k_mc_pass_tc3) with TMEM accumulators, or
// on the CUDA cores (k_mc_pass_simple, the bring-up / cross-check path). The O(C·P) integrator, prior,
// kinetic energy and Metropolis–Hastings step of every chain run in small per-chain kernels between passes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edhmc {

constexpr int kMcChainsPerCta = 128;  // chains per CTA: one TMEM lane / one thread per chain
constexpr int kMcTileRows = 128;      // rows of X per tensor-core tile
constexpr int kMcMaxD = 64;           // latents (features + bias) supported by the many-chain pass kernels

struct McArgs {
  // problem
  const float* X;
  const void* y;
  long long n_rows;
  long long ldx;
  int D;   // latents per chain: the columns of X plus, with a bias latent, one column of ones appended by the pre-tiling
  int Dx;  // columns physically present in X (D - 1 with a bias latent, else D)
  int Dp;  // D rounded up to a multiple of 8 (MMA K granularity for tf32)
  int family;
  int y_dtype;
  float lik_scale;
  const float* prior_loc;    // [D]
  const float* prior_scale;  // [D]
  double prior_const;
  int C;           // chains the pass kernels run: the caller's count rounded up to a multiple of 128 (padding chains sit at
                   // theta = 0 in the handle's own state arrays and are never read back)
  int Cu;          // the caller's chains: extent / stride of every caller-owned array (params, r0, u, trace, theta, logp, grad)
  int n_rowgroups;  // CTAs along the rows
  int seg_mode;     // development: 0 = float64 flush, 1 = synchronisation only, 2 = store only
  int seg_tiles;    // tensor-core pass: 64-row tiles accumulated in TMEM before a float64 flush (0 = default)
  int want_logp;    // 1: the pass also accumulates the log-likelihood (only needed at the end of a trajectory)
  // state, all [C][D] float32 unless noted
  float* z;
  float* r;
  float* g;
  float* zcur;
  float* gcur;
  double* logp_cur;  // [C]
  double* k_old;     // [C]
  double* log_u;     // [C]
  long long* n_accept;  // [C]
  int* valid;           // [1]
  int* need_init;       // [1]
  // pass output: per row group partial sums, [n_rowgroups][Dp][C] float64 gradient + [n_rowgroups][C] float64 logp
  double* part_g;
  double* part_lp;
  // run
  float* params;  // [T][C][D]
  long long t0;
  long long n_iter;
  float eps, half_eps;
  int L;
  const float* r0;  // [n_iter][C][D] or null
  const float* u;   // [n_iter][C] or null
  unsigned long long seed;
  double* trace;  // [n_iter][C][8] or null
  // pre-tiled operand copy of X for the tensor-core pass (k_mc_pretile): per 64-row tile {hi, lo} x [Dp/4][64][4] f32
  const float* xt;
  const float* yt;  // [ntiles64][64] float
  long long* dbg;  // optional per-role clock64 timeline of CTA (0,0): [tile][16]
  int dbg_lp;      // which passes stamp it: 1 = the log-likelihood passes, 0 = the gradient-only ones
};

// host launchers (chains.cu)
cudaError_t mc_prepare_tc();
cudaError_t mc_launch_pass(const McArgs& a, const float* theta, int use_tc, int gate, cudaStream_t s);
cudaError_t mc_launch_pretile(const McArgs& a, float* xt, float* yt, cudaStream_t s);
size_t mc_pretile_bytes(long long n_rows, int Dp, size_t* yt_bytes);
cudaError_t mc_launch_check(const McArgs& a, cudaStream_t s);
cudaError_t mc_launch_init_finish(const McArgs& a, cudaStream_t s);
cudaError_t mc_launch_begin(const McArgs& a, long long it, cudaStream_t s);
cudaError_t mc_launch_leap(const McArgs& a, long long it, int step, cudaStream_t s);
cudaError_t mc_launch_logp_grad_finish(const McArgs& a, const float* theta, double* logp, float* grad, cudaStream_t s);

}  // namespace edhmc
