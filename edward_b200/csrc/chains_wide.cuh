// chains_wide.cuh — many vectorised HMC chains for wide models (n_features > 64) and for row shards over several
// GPUs: BASELINE config 5 (N = 10M, D = 1,000, 1,024 chains, 8 GPUs). Shared between chains_wide.cu and edhmc.cu.
//
// Per leapfrog step the C chains need S = X·W (N×D · D×C), R = y − σ(S) and G = Xᵀ·R (D×C). With D = 1,000 neither
// W (1 MB per 128 chains as hi/lo planes) nor the G accumulator (1,000 TMEM columns) fits next to the operand
// ring of one SM, so the step is two tcgen05 GEMMs in 3xTF32 with R (hi/lo planes) parked in HBM between them —
// 10 GB of extra traffic per step per GPU at config 5, about 3 ms against ~50 ms of tensor work:
//   GEMM 1  Sᵀ[chain, row] = Wᵀ·Xᵀ, K = features; epilogue R = y − σ(Sᵀ) → hi/lo planes in GEMM 2's A layout
//   GEMM 2  G'[chain, feature] = R·X, K = local rows, split over CTAs; float64 partials per split
//   fold    float64 sums per chain → [grad, logp] buffer → (row shards) one ncclAllReduce → per-chain integrator
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edhmc {

constexpr int kMwRowTile = 128;   // rows of X per GEMM-1 tile (MMA N): main + correction accumulator, double-buffered = 512 TMEM columns
constexpr int kMwKC = 16;         // K elements per pipeline stage (two k-steps of 8)
constexpr int kMwSegChunks = 32;  // GEMM 2: stages accumulated in TMEM (fp32, truncating) before a float64 flush: 512 rows =
                                  // 192 accumulations (8,192 rows left a systematic 2e-5 shrink of the gradient at 1.25M rows)

struct McwArgs {
  // problem
  const float* X;
  const void* y;
  long long n_rows;
  long long ldx;
  int D;   // latents per chain (columns of X + the bias latent's column of ones, appended by the pre-tiling)
  int Dx;  // columns physically present in X
  int family;
  int y_dtype;
  float lik_scale;
  const float* prior_loc;
  const float* prior_scale;
  double prior_const;
  int C;   // chains the GEMMs run (the caller's count rounded up to a multiple of 128)
  int Cu;  // the caller's chains: extent / stride of every caller-owned array
  // tiling
  int nct;          // chain tiles of 128
  int Kp1;          // D rounded up to kMwKC
  int nrt;          // row tiles of 256
  long long rowsP;  // nrt * 256
  int nft, NB2, Dp2;  // GEMM 2: feature tiles, their width (multiple of 16, <= 256), nft * NB2
  int g1;           // GEMM 1: CTAs per chain tile
  int splits;       // GEMM 2: K splits per (chain tile, feature tile)
  int want_logp;
  // operand planes (hi then lo), no-swizzle K-major core-matrix layout
  float* xk;   // [nrt][2][Kp1/4][256][4]
  float* xt;   // [nft][2][rowsP/4][NB2][4]
  float* yt;   // [rowsP]
  float* wt;   // [nct][2][Kp1/4][128][4]
  float* rp;   // [nct][2][rowsP/4][128][4]
  // reduction buffers
  double* part_g64;  // [splits][Dp2][C]
  double* part_lp;   // [g1][C]
  double* gsum;      // [C][D+1]: likelihood gradient sums and log-likelihood, all-reduced over row shards
  // chain state, [C][D] float32 unless noted
  float* z;
  float* r;
  float* g;
  float* zcur;
  float* gcur;
  double* logp_cur;
  double* k_old;
  double* log_u;
  long long* n_accept;
  int* valid;
  int* need_init;
  // run
  float* params;
  long long t0;
  long long n_iter;
  float eps, half_eps;
  int L;
  const float* r0;
  const float* u;
  unsigned long long seed;
  double* trace;
};

struct McwSizes {
  size_t xk, xt, yt, wt, rp, part_g64, part_lp, gsum;
};
void mcw_plan(McwArgs& a, int num_sms);  // fills the tiling fields from n_rows, D, C
McwSizes mcw_sizes(const McwArgs& a);
cudaError_t mcw_prepare();
cudaError_t mcw_launch_pretile(const McwArgs& a, cudaStream_t s);
// one evaluation of [grad, logp] sums at theta ([C][D]) into a.gsum (local rows only)
cudaError_t mcw_launch_pass(const McwArgs& a, const float* theta, int gate, cudaStream_t s);
cudaError_t mcw_launch_check(const McwArgs& a, cudaStream_t s);
cudaError_t mcw_launch_init_finish(const McwArgs& a, cudaStream_t s);
cudaError_t mcw_launch_begin(const McwArgs& a, long long it, cudaStream_t s);
cudaError_t mcw_launch_leap(const McwArgs& a, long long it, int step, cudaStream_t s);
cudaError_t mcw_launch_logp_grad_finish(const McwArgs& a, const float* theta, double* logp, float* grad, cudaStream_t s);

}  // namespace edhmc
