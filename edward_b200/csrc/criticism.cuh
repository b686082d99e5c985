// criticism.cuh — posterior-predictive evaluation over the device-resident sample store.
//
// ed.evaluate (edward/criticisms/evaluate.py:20-235) runs the output variable n_samples times, each run with a fresh
// posterior draw of every latent (empirical.py:98-110), and averages: probabilities for a Bernoulli output
// (evaluate.py:132-143), draws for a continuous one (:158-162), the mean log-density for 'log_lik' (:222-227). For the
// GLMs of this path that is one contraction eta[n, s] = x_n . w_s (+ b_s) over the rows of X and S stored draws, followed
// by the family's link, reduced over s on the fly — eta [N, S] is never written.
//
// k_predictive: a CTA owns 64 rows and walks ALL draws in chunks of 64 (so every per-row sum is accumulated by one CTA in
// a fixed order: deterministic), the classic shared-memory register-tiled product (16 x 16 threads, 4 x 4 outputs each,
// K-chunks of 16 columns of X). The draws are gathered straight from the packed [T, ldp] store through the index lists.
// HBM traffic: X once (4ND bytes) + the gathered draws per CTA (L2-resident).
#pragma once
#include "common.cuh"

namespace edhmc {

constexpr int kPredRows = 64, kPredDraws = 64, kPredK = 16;

__global__ void __launch_bounds__(256) k_predictive(const float* __restrict__ X, long long n_rows, long long ldx, int D,
                                                    const void* __restrict__ y, int y_dtype, int family, float lik_scale,
                                                    const float* __restrict__ params, long long ldp,
                                                    const int* __restrict__ idx_w, const int* __restrict__ idx_b, int bias_col,
                                                    int S, float* __restrict__ mean_out, double* __restrict__ loglik_out) {
  __shared__ float xs[kPredK][kPredRows + 1];   // X tile, transposed: [k][row]
  __shared__ float ws[kPredK][kPredDraws + 1];  // draws tile: [k][draw]
  __shared__ float bs[kPredDraws];
  __shared__ float red_m[16][kPredRows];
  __shared__ double red_l[16][kPredRows];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // tx: draws, ty: rows
  const long long row0 = static_cast<long long>(blockIdx.x) * kPredRows;
  float yv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long n = row0 + ty * 4 + i;
    float v = 0.0f;
    if (n < n_rows && y != nullptr)
      v = y_dtype == EDHMC_Y_F32 ? reinterpret_cast<const float*>(y)[n]
                                 : (y_dtype == EDHMC_Y_U8 ? static_cast<float>(reinterpret_cast<const unsigned char*>(y)[n])
                                                          : static_cast<float>(reinterpret_cast<const int*>(y)[n]));
    yv[i] = v;
  }
  float msum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  double lsum[4] = {0.0, 0.0, 0.0, 0.0};
  const bool want_ll = loglik_out != nullptr;
  const float log_norm = 0.9189385332046727f + logf(lik_scale);
  for (int s0 = 0; s0 < S; s0 += kPredDraws) {
    __syncthreads();  // the previous chunk's epilogue has read bs
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    if (tid < kPredDraws) {
      const int s = s0 + tid;
      bs[tid] = (idx_b != nullptr && s < S) ? params[static_cast<long long>(idx_b[s]) * ldp + bias_col] : 0.0f;
    }
    for (int k0 = 0; k0 < D; k0 += kPredK) {
      __syncthreads();
      // X tile: 64 rows x 16 columns, coalesced along the columns
      for (int e = tid; e < kPredRows * kPredK; e += 256) {
        const int r = e / kPredK, k = e % kPredK;
        const long long n = row0 + r;
        xs[k][r] = (n < n_rows && k0 + k < D) ? X[n * ldx + k0 + k] : 0.0f;
      }
      for (int e = tid; e < kPredDraws * kPredK; e += 256) {
        const int d = e / kPredK, k = e % kPredK;
        const int s = s0 + d;
        ws[k][d] = (s < S && k0 + k < D) ? params[static_cast<long long>(idx_w[s]) * ldp + k0 + k] : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kPredK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = xs[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = ws[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    // link + reduction over this thread's 4 draws of the chunk (draw order inside a row: fixed)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = tx + 16 * j;
      if (s0 + d < S) {
        const float bias = bs[d];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float eta = acc[i][j] + bias;
          if (family == EDHMC_BERNOULLI_LOGIT) {
            const float e = expf(-fabsf(eta));
            const float q = e / (1.0f + e);
            msum[i] += eta >= 0.0f ? 1.0f - q : q;  // sigmoid(eta)
            if (want_ll) lsum[i] += static_cast<double>(-(fmaxf(eta, 0.0f) - eta * yv[i] + log1pf(e)));
          } else if (family == EDHMC_NORMAL_IDENTITY) {
            msum[i] += eta;
            if (want_ll) {
              const float zz = (yv[i] - eta) / lik_scale;
              lsum[i] += static_cast<double>(-0.5f * zz * zz - log_norm);
            }
          } else {
            const float mu = expf(eta);
            msum[i] += mu;
            if (want_ll) lsum[i] += static_cast<double>(yv[i] * eta - mu - lgammaf(yv[i] + 1.0f));
          }
        }
      }
    }
  }
  // the 16 threads tx = 0..15 of a row: fixed-order sum through shared memory
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red_m[tx][ty * 4 + i] = msum[i];
    red_l[tx][ty * 4 + i] = lsum[i];
  }
  __syncthreads();
  if (tid < kPredRows) {
    const long long n = row0 + tid;
    if (n < n_rows) {
      float m = 0.0f;
      double l = 0.0;
      for (int t = 0; t < 16; ++t) {
        m += red_m[t][tid];
        l += red_l[t][tid];
      }
      mean_out[n] = m / static_cast<float>(S);
      if (want_ll) loglik_out[n] = l;
    }
  }
}

}  // namespace edhmc

