// chains.cu — many vectorised HMC chains over one design matrix (see chains.cuh).
//
//   k_mc_pass_tc      the dense contraction on the 5th-generation tensor cores, hand-written tcgen05:
//                       Sᵀ[chain, row] = Wᵀ·Xᵀ      tcgen05.mma kind::tf32, A = Wᵀ (smem), B = X tile (smem),
//                                                    accumulator in TMEM (chains on lanes, rows on columns)
//                       R = y − σ(Sᵀ)                epilogue: tcgen05.ld → CUDA cores → tcgen05.st (R stays in TMEM)
//                       G'[chain, d] += R·X          tcgen05.mma with A = R read straight from TMEM, B = X tile
//                     both contractions in 3xTF32 (hi·hi + hi·lo + lo·hi) so the results hold fp32 tolerance.
//   k_mc_pass_simple  the same pass on the CUDA cores (bring-up / cross-check; any ldx, D <= 64).
//   k_mc_*            per-chain leapfrog / prior / kinetic / Metropolis–Hastings kernels (hmc.py:81-130,195-210).
#include "chains.cuh"

#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace edhmc {

// ------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_range(long long n_rows, int nrg, int rg, long long& t0, long long& t1) {
  const long long ntiles = (n_rows + kMcTileRows - 1) / kMcTileRows;
  t0 = ntiles * rg / nrg;
  t1 = ntiles * (rg + 1) / nrg;
}

__device__ __forceinline__ float ld_y(const void* y, int y_dtype, long long i) {
  return y_dtype == 0 ? static_cast<float>(reinterpret_cast<const int*>(y)[i]) : reinterpret_cast<const float*>(y)[i];
}

// ------------------------------------------------------------------------------------------------
// CUDA-core pass: thread = chain, rows broadcast from shared memory.
// grid (n_rowgroups, C/128), block 128.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMcChainsPerCta, 1) k_mc_pass_simple(const McArgs a, const float* theta, int gate) {
  if (gate && !*a.need_init) return;
  __shared__ float xs[32][kMcMaxD];
  __shared__ float ys[32];
  const int tid = threadIdx.x;
  const int c = blockIdx.y * kMcChainsPerCta + tid;
  const int D = a.D;
  float w[kMcMaxD], g[kMcMaxD];
#pragma unroll
  for (int d = 0; d < kMcMaxD; ++d) {
    w[d] = d < D ? theta[static_cast<size_t>(c) * D + d] : 0.0f;
    g[d] = 0.0f;
  }
  double lp = 0.0;
  long long t0, t1;
  tile_range(a.n_rows, a.n_rowgroups, blockIdx.x, t0, t1);
  const long long row_lo = t0 * kMcTileRows;
  long long row_hi = t1 * kMcTileRows;
  if (row_hi > a.n_rows) row_hi = a.n_rows;
  for (long long base = row_lo; base < row_hi; base += 32) {
    const int nr = static_cast<int>(row_hi - base < 32 ? row_hi - base : 32);
    __syncthreads();
    for (int i = tid; i < 32 * D; i += kMcChainsPerCta) {
      const int m = i / D, d = i - m * D;
      xs[m][d] = m < nr ? a.X[(base + m) * a.ldx + d] : 0.0f;
    }
    if (tid < 32) ys[tid] = tid < nr ? ld_y(a.y, a.y_dtype, base + tid) : 0.0f;
    __syncthreads();
    for (int m = 0; m < nr; ++m) {
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int d = 0; d < kMcMaxD; d += 2) {
        if (d < D) s0 = fmaf(xs[m][d], w[d], s0);
        if (d + 1 < D) s1 = fmaf(xs[m][d + 1], w[d + 1], s1);
      }
      float lpv, rv;
      row_terms(a.family, s0 + s1, ys[m], a.lik_scale, lpv, rv);
      lp += static_cast<double>(lpv);
#pragma unroll
      for (int d = 0; d < kMcMaxD; ++d)
        if (d < D) g[d] = fmaf(rv, xs[m][d], g[d]);
    }
  }
  float* pg = a.part_g + (static_cast<size_t>(blockIdx.x) * a.C + c) * a.Dp;
#pragma unroll
  for (int d = 0; d < kMcMaxD; ++d)
    if (d < a.Dp) pg[d] = d < D ? g[d] : 0.0f;
  a.part_lp[static_cast<size_t>(blockIdx.x) * a.C + c] = lp;
}

// ------------------------------------------------------------------------------------------------
// tensor-core pass. grid (n_rowgroups, C/128), block 256 (8 warps), one CTA per SM.
// Shared memory (dynamic), all operand tiles in the no-swizzle K-major core-matrix layout:
//   raw   [128][ldx] f32      the X tile as it lies in HBM (one TMA bulk copy)
//   A1h/l [Dp/4][128][4]      Wᵀ block: off(c,d) = (d/4)*2048 + c*16 + (d%4)*4          (M=128 chains, K=Dp)
//   B1h/l [Dp/4][128][4]      X tile:   off(m,d) = (d/4)*2048 + m*16 + (d%4)*4          (N=128 rows,   K=Dp)
//   B2h/l [128/4][64][4]      X tileᵀ:  off(d,m) = (m/4)*1024 + d*16 + (m%4)*4          (N=64 feats,   K=128 rows)
// TMEM (512 columns allocated): [0,128) Sᵀ then R_hi in place, [128,256) R_lo, [256,320) G' accumulator.
// ------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 256;
constexpr int kTcN2 = 64;

__host__ __device__ inline int tc_smem_layout(int Dp, int ldx, int* off /*8*/) {
  int o = 0;
  off[0] = o;  // raw
  o += kMcTileRows * ldx * 4;
  o = (o + 127) / 128 * 128;
  off[1] = o;  // ys
  o += kMcTileRows * 4;
  off[2] = o;  // A1 hi, lo
  o += 2 * (Dp / 4) * 2048;
  off[3] = o;  // B1 hi, lo
  o += 2 * (Dp / 4) * 2048;
  off[4] = o;  // B2 hi, lo
  o += 2 * (kMcTileRows / 4) * 1024;
  off[5] = o;  // mbarriers (3) + tmem address
  o += 64;
  off[6] = o;  // logp combine [2][128] doubles
  o += 2 * 128 * 8;
  return o;
}

__global__ void __launch_bounds__(kTcThreads, 1) k_mc_pass_tc(const McArgs a, const float* theta, int gate) {
  if (gate && !*a.need_init) return;
  extern __shared__ __align__(128) unsigned char smem[];
  int off[8];
  const int ldx = static_cast<int>(a.ldx);
  const int Dp = a.Dp, D = a.D;
  tc_smem_layout(Dp, ldx, off);
  float* raw = reinterpret_cast<float*>(smem + off[0]);
  float* ys = reinterpret_cast<float*>(smem + off[1]);
  unsigned char* A1 = smem + off[2];
  unsigned char* B1 = smem + off[3];
  unsigned char* B2 = smem + off[4];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off[5]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off[5] + 32);
  double* lpc = reinterpret_cast<double*>(smem + off[6]);
  const int kc1 = Dp / 4;
  const int a1_half = kc1 * 2048, b2_half = (kMcTileRows / 4) * 1024;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.y * kMcChainsPerCta;  // first chain of this CTA
  const uint32_t bar_raw = smem_u32(&bars[0]), bar_m1 = smem_u32(&bars[1]), bar_m2 = smem_u32(&bars[2]);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  // A1 = Wᵀ block, split hi/lo: item (c, kc) → 16 bytes
  for (int i = tid; i < kMcChainsPerCta * kc1; i += kTcThreads) {
    const int c = i % kMcChainsPerCta, kc = i / kMcChainsPerCta;
    float4 h, l;
    float* hp = &h.x;
    float* lp_ = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kc * 4 + e;
      const float v = d < D ? theta[static_cast<size_t>(cb + c) * D + d] : 0.0f;
      split_tf32(v, hp[e], lp_[e]);
    }
    *reinterpret_cast<float4*>(A1 + kc * 2048 + c * 16) = h;
    *reinterpret_cast<float4*>(A1 + a1_half + kc * 2048 + c * 16) = l;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_s = tm, tm_lo = tm + 128, tm_g = tm + 256;

  long long t0, t1;
  tile_range(a.n_rows, a.n_rowgroups, blockIdx.x, t0, t1);
  const int ntiles = static_cast<int>(t1 - t0);
  const uint32_t id1 = idesc_tf32(128, 128), id2 = idesc_tf32(128, kTcN2);
  const int q = warp & 3, hf = warp >> 2;  // TMEM lane quadrant (chains 32q..), column half (tile rows 64hf..)
  double lp = 0.0;

  // raw tile loader: one TMA bulk copy (full tiles) or cooperative plain loads (the ragged last tile of X)
  auto load_raw = [&](int it) {
    const long long row0 = (t0 + it) * kMcTileRows;
    const long long left = a.n_rows - row0;
    const int rows = left < kMcTileRows ? static_cast<int>(left) : kMcTileRows;
    if (rows == kMcTileRows) {
      if (tid == 0) {
        const uint32_t bytes = kMcTileRows * ldx * 4;
        fence_proxy_async_smem();
        mbar_arrive_expect_tx_s(bar_raw, bytes);
        bulk_g2s_s(smem_u32(raw), a.X + row0 * a.ldx, bytes, bar_raw);
      }
    } else {
      for (int i = tid; i < rows * ldx; i += kTcThreads) {
        const int m = i / ldx, d = i - m * ldx;
        raw[i] = d < D ? a.X[(row0 + m) * a.ldx + d] : 0.0f;
      }
      __syncthreads();
      if (tid == 0) mbar_arrive(&bars[0]);
    }
  };

  if (ntiles > 0) load_raw(0);
  for (int it = 0; it < ntiles; ++it) {
    const uint32_t par = it & 1;
    const long long row0 = (t0 + it) * kMcTileRows;
    const long long left = a.n_rows - row0;
    const int rows = left < kMcTileRows ? static_cast<int>(left) : kMcTileRows;
    // MMA2 of the previous tile must be done before B1/B2/S are overwritten
    if (it > 0) mbar_wait_s(bar_m2, (it - 1) & 1);
    mbar_wait_s(bar_raw, par);
    if (tid < kMcTileRows) ys[tid] = tid < rows ? ld_y(a.y, a.y_dtype, row0 + tid) : 0.0f;
    // ---- build B1 (rows x K=Dp) and B2 (feats x K=rows), hi/lo ----
    for (int i = tid; i < kMcTileRows * kc1; i += kTcThreads) {
      const int m = i % kMcTileRows, kc = i / kMcTileRows;
      float4 h, l;
      float* hp = &h.x;
      float* lq = &l.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = kc * 4 + e;
        const float v = (m < rows && d < D) ? raw[m * ldx + d] : 0.0f;
        split_tf32(v, hp[e], lq[e]);
      }
      *reinterpret_cast<float4*>(B1 + kc * 2048 + m * 16) = h;
      *reinterpret_cast<float4*>(B1 + a1_half + kc * 2048 + m * 16) = l;
    }
    for (int i = tid; i < kTcN2 * (kMcTileRows / 4); i += kTcThreads) {
      const int d = i % kTcN2, mc = i / kTcN2;
      float4 h, l;
      float* hp = &h.x;
      float* lq = &l.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = mc * 4 + e;
        const float v = (m < rows && d < D) ? raw[m * ldx + d] : 0.0f;
        split_tf32(v, hp[e], lq[e]);
      }
      *reinterpret_cast<float4*>(B2 + mc * 1024 + d * 16) = h;
      *reinterpret_cast<float4*>(B2 + b2_half + mc * 1024 + d * 16) = l;
    }
    fence_proxy_async_smem();  // generic-proxy writes of the operand tiles → visible to the tensor core (async proxy)
    __syncthreads();
    // raw is consumed: prefetch the next tile while the tensor core and the epilogue work
    if (it + 1 < ntiles) load_raw(it + 1);

    // ---- MMA1: Sᵀ[128 chains x 128 rows] = Wᵀ·Xᵀ, K = Dp, 3xTF32 ----
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_h = smem_u32(A1), a_l = a_h + a1_half, b_h = smem_u32(B1), b_l = b_h + a1_half;
      for (int ks = 0; ks < Dp / 8; ++ks) {
        const uint32_t o = ks * 4096;  // 2 core matrices along K per MMA (K = 8 tf32)
        tc_mma_ss(tm_s, smem_desc(a_h + o, 2048, 128), smem_desc(b_h + o, 2048, 128), id1, ks > 0);
        tc_mma_ss(tm_s, smem_desc(a_h + o, 2048, 128), smem_desc(b_l + o, 2048, 128), id1, 1);
        tc_mma_ss(tm_s, smem_desc(a_l + o, 2048, 128), smem_desc(b_h + o, 2048, 128), id1, 1);
      }
      tc_commit(bar_m1);
    }
    mbar_wait_s(bar_m1, par);
    tc_fence_after();

    // ---- epilogue: R = dlogp/deta, written back to TMEM as the A operand of MMA2 (hi in place, lo beside) ----
    {
      const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col = hf * 64 + cc * 16;
        uint32_t v[16], vh[16], vl[16];
        tmem_ld16(tm_s + lane_base + col, v);
        tmem_wait_ld();
        if (a.want_logp || a.family != 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = col + j;
            float lpv, rv;
            row_terms(a.family, __uint_as_float(v[j]), ys[m], a.lik_scale, lpv, rv);
            if (m >= rows) {
              lpv = 0.0f;
              rv = 0.0f;
            }
            lp += static_cast<double>(lpv);
            float h, l;
            split_tf32(rv, h, l);
            vh[j] = __float_as_uint(h);
            vl[j] = __float_as_uint(l);
          }
        } else {
          // inside a trajectory only the residual y - sigmoid(eta) is needed (the log joint enters the
          // Metropolis–Hastings ratio at the trajectory's end only): one exp and one reciprocal per element
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = col + j;
            const float eta = __uint_as_float(v[j]);
            const float e = expf(-fabsf(eta));
            const float qv = __fdividef(e, 1.0f + e);
            const float yv = ys[m];
            float rv = eta >= 0.0f ? (yv - 1.0f) + qv : yv - qv;
            if (m >= rows) rv = 0.0f;
            float h, l;
            split_tf32(rv, h, l);
            vh[j] = __float_as_uint(h);
            vl[j] = __float_as_uint(l);
          }
        }
        tmem_st16(tm_s + lane_base + col, vh);
        tmem_st16(tm_lo + lane_base + col, vl);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();

    // ---- MMA2: G'[128 chains x 64 feats] += R·X, K = 128 rows, A from TMEM, 3xTF32 ----
    if (tid == 0) {
      tc_fence_after();
      const uint32_t b_h = smem_u32(B2), b_l = b_h + b2_half;
      for (int ks = 0; ks < kMcTileRows / 8; ++ks) {
        const uint32_t o = ks * 2048;
        tc_mma_ts(tm_g, tm_s + ks * 8, smem_desc(b_h + o, 1024, 128), id2, (it > 0 || ks > 0));
        tc_mma_ts(tm_g, tm_s + ks * 8, smem_desc(b_l + o, 1024, 128), id2, 1);
        tc_mma_ts(tm_g, tm_lo + ks * 8, smem_desc(b_h + o, 1024, 128), id2, 1);
      }
      tc_commit(bar_m2);
    }
  }

  // ---- write this row group's partial sums ----
  if (ntiles > 0) {
    mbar_wait_s(bar_m2, (ntiles - 1) & 1);
    tc_fence_after();
  }
  {
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int chain = cb + 32 * q + lane;
    float* pg = a.part_g + (static_cast<size_t>(blockIdx.x) * a.C + chain) * Dp;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int col = hf * 32 + cc * 16;
      uint32_t v[16];
      if (ntiles > 0) {
        tmem_ld16(tm_g + lane_base + col, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < Dp) pg[col + j] = __uint_as_float(v[j]);
    }
    lpc[hf * 128 + 32 * q + lane] = lp;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < kMcChainsPerCta) a.part_lp[static_cast<size_t>(blockIdx.x) * a.C + cb + tid] = lpc[tid] + lpc[128 + tid];
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Pipelined tensor-core pass (default). Same mathematics as k_mc_pass_tc, restructured so that the tensor
// pipe, the operand builders and the epilogue work on different tiles at the same time:
//   warp 0      control: TMA bulk loads of the raw X tiles + MMA1 issue (one thread); warp 13: MMA2 issue
//   warps 1-4   builders: raw tile → B1 (rows x K=Dp) and B2 (feats x K=rows) operand tiles, hi/lo split
//   warps 5-12  epilogue: Sᵀ (TMEM) → residual R (TMEM, hi in place + lo), log-likelihood
// 64-row tiles, every per-tile resource double-buffered (b = tile & 1):
//   TMEM  S[b] 64 cols at 64*b, R_lo[b] 64 cols at 128+64*b, G' 64 cols at 256   (512 allocated)
//   mbarriers  raw_full[b] (TMA) → b_ready[b] (128 builder arrivals) → s_ready[b] (tcgen05.commit after MMA1)
//              → r_ready[b] (256 epilogue arrivals) → mma2_done[b] (tcgen05.commit after MMA2, frees B[b])
// MMA1(i+2) may overwrite S[b] only after MMA2(i) has read R[b]: b_ready[b](i+2) implies it, because the
// builders wait for mma2_done[b](i) before they rebuild B[b].
// ------------------------------------------------------------------------------------------------
#define MC_DBG(slot) do { if (dbgp && i < 64) dbgp[i * 16 + (slot)] = clock64(); } while (0)
constexpr int kT2Rows = 64;
constexpr int kT2Threads = 14 * 32;  // control/MMA1, 4 builders, 8 epilogue, MMA2 issuer
constexpr int kT2Builders = 128;
constexpr int kT2Epilogue = 256;

__host__ __device__ inline int tc2_smem_layout(int Dp, int ldx, int* off /*8*/) {
  int o = 0;
  off[0] = o;  // raw[2] (also reused for the final logp combine)
  const int raw_bytes = (kT2Rows * ldx * 4 + 32 + 127) / 128 * 128;  // + slack: masked over-read of the last row
  o += 2 * raw_bytes;
  off[1] = o;  // ys[2][64]
  o += 2 * kT2Rows * 4;
  off[2] = o;  // A1 hi, lo
  o += 2 * (Dp / 4) * 2048;
  off[3] = o;  // B1[2] {hi, lo}
  o += 2 * 2 * (Dp / 4) * 1024;
  off[4] = o;  // B2[2] {hi, lo}
  o += 2 * 2 * (kT2Rows / 4) * 1024;
  off[5] = o;  // 10 mbarriers + tmem slot
  o += 128;
  off[6] = raw_bytes;
  return o;
}

__global__ void __launch_bounds__(kT2Threads, 1) k_mc_pass_tc2(const McArgs a, const float* theta, int gate) {
  if (gate && !*a.need_init) return;
  extern __shared__ __align__(128) unsigned char smem[];
  int off[8];
  const int ldx = static_cast<int>(a.ldx);
  const int Dp = a.Dp, D = a.D;
  tc2_smem_layout(Dp, ldx, off);
  const int raw_bytes = off[6];
  unsigned char* raw0 = smem + off[0];
  float* ysm = reinterpret_cast<float*>(smem + off[1]);
  unsigned char* A1 = smem + off[2];
  unsigned char* B1 = smem + off[3];
  unsigned char* B2 = smem + off[4];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off[5]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off[5] + 96);
  const int kc1 = Dp / 4;
  const int a1_half = kc1 * 2048;
  const int b1_half = kc1 * 1024, b1_buf = 2 * b1_half;
  const int b2_half = (kT2Rows / 4) * 1024, b2_buf = 2 * b2_half;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.y * kMcChainsPerCta;
  // barrier indices: 0,1 raw_full  2,3 b_ready  4,5 s_ready  6,7 r_ready  8,9 mma2_done
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int kind, int b) { return bar0 + static_cast<uint32_t>((kind * 2 + b) * 8); };

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[0 + b], 1);
      mbar_init(&bars[2 + b], kT2Builders);
      mbar_init(&bars[4 + b], 1);
      mbar_init(&bars[6 + b], kT2Epilogue);
      mbar_init(&bars[8 + b], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = tid; i < kMcChainsPerCta * kc1; i += kT2Threads) {
    const int c = i % kMcChainsPerCta, kc = i / kMcChainsPerCta;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kc * 4 + e;
      const float v = d < D ? theta[static_cast<size_t>(cb + c) * D + d] : 0.0f;
      split_tf32(v, hp[e], lq[e]);
    }
    *reinterpret_cast<float4*>(A1 + kc * 2048 + c * 16) = h;
    *reinterpret_cast<float4*>(A1 + a1_half + kc * 2048 + c * 16) = l;
  }
  for (int i = tid; i < 2 * b2_buf / 16; i += kT2Threads) reinterpret_cast<float4*>(B2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_g = tm + 256;

  // this row group's tiles of 64 rows (two per 128-row unit of the shared partition)
  long long u0, u1;
  tile_range(a.n_rows, a.n_rowgroups, blockIdx.x, u0, u1);
  const long long row_begin = u0 * kMcTileRows;
  long long row_end = u1 * kMcTileRows;
  if (row_end > a.n_rows) row_end = a.n_rows;
  const int nt = row_end > row_begin ? static_cast<int>((row_end - row_begin + kT2Rows - 1) / kT2Rows) : 0;
  const uint32_t id12 = idesc_tf32(128, kT2Rows);  // MMA1: N = 64 rows;  MMA2: N = 64 features
  double lp = 0.0;
  long long* dbgp = (a.dbg && !a.want_logp && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? a.dbg : nullptr;

  if (warp == 0) {
    // ================= control warp: TMA loads of raw tiles + MMA1 issue =================
    // The whole warp runs the loop (converged); one elected lane issues. Descriptors are built once; per
    // k-step only the 14-bit start-address field moves (+bytes/16).
    auto issue_raw = [&](int i) {
      const long long r0 = row_begin + static_cast<long long>(i) * kT2Rows;
      if (row_end - r0 >= kT2Rows) {
        if (elect_one()) {
          const int b = i & 1;
          const uint32_t bytes = kT2Rows * ldx * 4;
          fence_proxy_async_smem();
          mbar_arrive_expect_tx_s(BAR(0, b), bytes);
          bulk_g2s_s(smem_u32(raw0 + b * raw_bytes), a.X + r0 * a.ldx, bytes, BAR(0, b));
        }
        __syncwarp();
      }
    };
    for (int i = 0; i < nt && i < 2; ++i) issue_raw(i);
    const uint64_t da_h = smem_desc(smem_u32(A1), 2048, 128), da_l = smem_desc(smem_u32(A1) + a1_half, 2048, 128);
    const uint64_t db_h0 = smem_desc(smem_u32(B1), 1024, 128), db_l0 = smem_desc(smem_u32(B1) + b1_half, 1024, 128);
    const int nks = Dp / 8;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const uint32_t par = (i >> 1) & 1;
      mbar_wait_s(BAR(1, b), par);  // b_ready: operand tiles built, raw[b] consumed
      MC_DBG(0);
      if (i + 2 < nt) issue_raw(i + 2);
      tc_fence_after();
      const uint64_t db_h = db_h0 + static_cast<uint64_t>((b * b1_buf) >> 4), db_l = db_l0 + static_cast<uint64_t>((b * b1_buf) >> 4);
      const uint32_t ts = tm + 64 * b;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // Sᵀ[b] = Wᵀ·X(i)ᵀ, K = Dp = 8*nks
          if (ks < nks) {
            const uint64_t oa = static_cast<uint64_t>(ks * (4096 >> 4)), ob = static_cast<uint64_t>(ks * (2048 >> 4));
            tc_mma_ss(ts, da_h + oa, db_h + ob, id12, ks > 0);
            tc_mma_ss(ts, da_h + oa, db_l + ob, id12, 1);
            tc_mma_ss(ts, da_l + oa, db_h + ob, id12, 1);
          }
        }
        tc_commit(BAR(2, b));  // s_ready[b]
      }
      __syncwarp();
      MC_DBG(1);
    }
  } else if (warp == 13) {
    // ================= MMA2 issuer warp: G' += R(i)·X(i), A = R from TMEM =================
    const uint64_t d2_h0 = smem_desc(smem_u32(B2), 1024, 128), d2_l0 = smem_desc(smem_u32(B2) + b2_half, 1024, 128);
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      mbar_wait_s(BAR(3, b), (i >> 1) & 1);  // r_ready[b]
      MC_DBG(2);
      tc_fence_after();
      const uint64_t d2_h = d2_h0 + static_cast<uint64_t>((b * b2_buf) >> 4), d2_l = d2_l0 + static_cast<uint64_t>((b * b2_buf) >> 4);
      const uint32_t r_h = tm + 64 * b, r_l = tm + 128 + 64 * b;
      const uint32_t first = (i > 0) ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kT2Rows / 8; ++ks) {
          const uint64_t o = static_cast<uint64_t>(ks * (2048 >> 4));
          tc_mma_ts(tm_g, r_h + ks * 8, d2_h + o, id12, ks > 0 ? 1u : first);
          tc_mma_ts(tm_g, r_h + ks * 8, d2_l + o, id12, 1);
          tc_mma_ts(tm_g, r_l + ks * 8, d2_h + o, id12, 1);
        }
        tc_commit(BAR(4, b));  // mma2_done[b]: B[b], ys[b], S[b] are free again
      }
      __syncwarp();
      MC_DBG(3);
    }
    if (nt > 0) mbar_wait_s(BAR(4, (nt - 1) & 1), ((nt - 1) >> 1) & 1);  // the last commit covers every MMA2
  } else if (warp <= 4) {
    // ================= builders =================
    // Each work item is a 4x4 block (rows 4*mb.., features 4*kc..): loaded once from the raw tile, split once
    // into hi/lo, and written both as 4 row-chunks of B1 and as 4 feature-chunks of B2.
    const int bt = tid - 32;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const long long r0 = row_begin + static_cast<long long>(i) * kT2Rows;
      const int rows = row_end - r0 >= kT2Rows ? kT2Rows : static_cast<int>(row_end - r0);
      if (i >= 2) mbar_wait_s(BAR(4, b), par ^ 1u);  // MMA2(i-2) has finished reading B[b] / ys[b]
      if (warp == 1) MC_DBG(4);
      float* rawb = reinterpret_cast<float*>(raw0 + b * raw_bytes);
      if (rows == kT2Rows) {
        mbar_wait_s(BAR(0, b), par);
      } else {  // ragged last tile: stage it through the (idle) raw buffer with plain loads
        for (int e = bt; e < rows * ldx; e += kT2Builders) {
          const int m = e / ldx, d = e - m * ldx;
          rawb[e] = d < D ? a.X[(r0 + m) * a.ldx + d] : 0.0f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kT2Builders) : "memory");
      }
      if (warp == 1) MC_DBG(5);
      if (bt < kT2Rows) ysm[b * kT2Rows + bt] = bt < rows ? ld_y(a.y, a.y_dtype, r0 + bt) : 0.0f;
      unsigned char* b1 = B1 + b * b1_buf;
      unsigned char* b2 = B2 + b * b2_buf;
      // thread → (row block mb = bt & 15, feature chunks kc = (bt >> 4) + 8j): no divisions, 64-bit shared loads
      const int mb = bt & 15;
      for (int kc = bt >> 4; kc < kc1; kc += 8) {
        float v[4][4];
        const bool even = (ldx & 1) == 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int m = mb * 4 + r;
          const float* rp = rawb + m * ldx + kc * 4;
          if (even) {
            const float2 p0 = *reinterpret_cast<const float2*>(rp);
            const float2 p1 = *reinterpret_cast<const float2*>(rp + 2);
            v[r][0] = p0.x;
            v[r][1] = p0.y;
            v[r][2] = p1.x;
            v[r][3] = p1.y;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[r][e] = rp[e];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (m >= rows || kc * 4 + e >= D) v[r][e] = 0.0f;
        }
        float hi[4][4], lo[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(v[r][e], hi[r][e], lo[r][e]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // B1: row 4mb+r, features 4kc..4kc+3
          const int o = kc * 1024 + (mb * 4 + r) * 16;
          *reinterpret_cast<float4*>(b1 + o) = make_float4(hi[r][0], hi[r][1], hi[r][2], hi[r][3]);
          *reinterpret_cast<float4*>(b1 + b1_half + o) = make_float4(lo[r][0], lo[r][1], lo[r][2], lo[r][3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // B2: feature 4kc+e, rows 4mb..4mb+3
          const int o = mb * 1024 + (kc * 4 + e) * 16;
          *reinterpret_cast<float4*>(b2 + o) = make_float4(hi[0][e], hi[1][e], hi[2][e], hi[3][e]);
          *reinterpret_cast<float4*>(b2 + b2_half + o) = make_float4(lo[0][e], lo[1][e], lo[2][e], lo[3][e]);
        }
      }
      // features Dp..63 of B2 are never written by the loop above: they stay zero from the one-time fill
      if (warp == 1) MC_DBG(6);
      fence_proxy_async_smem();
      mbar_arrive(&bars[2 + b]);
      if (warp == 1) MC_DBG(7);
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3, hf = (warp - 5) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const long long r0 = row_begin + static_cast<long long>(i) * kT2Rows;
      const int rows = row_end - r0 >= kT2Rows ? kT2Rows : static_cast<int>(row_end - r0);
      mbar_wait_s(BAR(2, b), par);
      if (warp == 5) MC_DBG(8);
      tc_fence_after();
      const float* ys = ysm + b * kT2Rows;
      uint32_t vv[2][16];
      tmem_ld16(tm + 64 * b + lane_base + hf * 32, vv[0]);
      tmem_ld16(tm + 64 * b + lane_base + hf * 32 + 16, vv[1]);
      tmem_wait_ld();
      if (warp == 5) MC_DBG(9);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col = hf * 32 + cc * 16;
        uint32_t vh[16], vl[16];
        const uint32_t* v = vv[cc];
        if (a.want_logp || a.family != 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = col + j;
            float lpv, rv;
            row_terms(a.family, __uint_as_float(v[j]), ys[m], a.lik_scale, lpv, rv);
            if (m >= rows) {
              lpv = 0.0f;
              rv = 0.0f;
            }
            lp += static_cast<double>(lpv);
            float h, l;
            split_tf32(rv, h, l);
            vh[j] = __float_as_uint(h);
            vl[j] = __float_as_uint(l);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = col + j;
            const float eta = __uint_as_float(v[j]);
            const float e = expf(-fabsf(eta));
            const float qv = __fdividef(e, 1.0f + e);
            const float yv = ys[m];
            float rv = eta >= 0.0f ? (yv - 1.0f) + qv : yv - qv;
            if (m >= rows) rv = 0.0f;
            float h, l;
            split_tf32(rv, h, l);
            vh[j] = __float_as_uint(h);
            vl[j] = __float_as_uint(l);
          }
        }
        tmem_st16(tm + 64 * b + lane_base + col, vh);
        tmem_st16(tm + 128 + 64 * b + lane_base + col, vl);
      }
      if (warp == 5) MC_DBG(10);
      tmem_wait_st();
      if (warp == 5) MC_DBG(11);
      tc_fence_before();
      mbar_arrive(&bars[6 + b]);
      if (warp == 5) MC_DBG(12);
    }
  }

  // ---- write this row group's partial sums ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  double* lpc = reinterpret_cast<double*>(raw0);  // [2][128], the raw buffers are idle now
  if (warp >= 5 && warp <= 12) {
    const int q = warp & 3, hf = (warp - 5) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int chain = cb + 32 * q + lane;
    float* pg = a.part_g + (static_cast<size_t>(blockIdx.x) * a.C + chain) * Dp;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int col = hf * 32 + cc * 16;
      uint32_t v[16];
      if (nt > 0) {
        tmem_ld16(tm_g + lane_base + col, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < Dp) pg[col + j] = __uint_as_float(v[j]);
    }
    lpc[hf * 128 + 32 * q + lane] = lp;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < kMcChainsPerCta) a.part_lp[static_cast<size_t>(blockIdx.x) * a.C + cb + tid] = lpc[tid] + lpc[128 + tid];
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core pass, version 3 (default): no per-pass hi/lo splitting, no builder warps.
//
// X is re-laid once (k_mc_pretile, at the first many-chain call after bind) into the UMMA operand layout:
// per 64-row tile a hi plane and a lo plane (3xTF32 split), each [Dp/4][64 rows][4 floats] with a plane-row
// pitch of 1040 B (K-major, no-swizzle core matrices; the 16 B of padding per 1024 make the transposing
// reads below bank-conflict free). A tile is ONE contiguous TMA bulk copy and is used as it lands by
//   MMA1  Sᵀ = Wᵀ·Xᵀ   B = tile, N=64 rows x K=Dp, K-major (LBO = 1040: next 4 features, SBO = 128: next 8 rows)
// The second contraction needs the tile with rows as the K dimension. (An MN-major descriptor over the same
// bytes returned zeros on this part, see DESIGN.md.) The epilogue warps therefore transpose the landed tile
// once, 4x4 blocks through registers, into B2T[b] ([64/4][64 feats][4 rows], SBO = 144, LBO = 1152: padded so
// the transposing stores are conflict free), while MMA1 of the same tile runs:
//   MMA2  G' += R·X     A = R from TMEM, B = B2T[b], N=64 feats x K=64 rows, K-major
// Warp roles (20 warps): 0 TMA producer, 1 MMA1 issuer, 2 MMA2 issuer, 3 idle, 4-19 epilogue (TMEM lane quadrant =
// warp % 4, 16 of the tile's 64 columns each). NS-stage operand ring (3, or 2 when Dp = 64), S/R and B2T
// double-buffered (b = tile & 1).
//   x_full[s] (TMA) → {MMA1 → s_ready[b]} ‖ {transpose → B2T[b]} → epilogue → r_ready[b] (512 arrivals)
//   → MMA2 → x_empty[s] (commit).  Tile i may overwrite S[b] / B2T[b] only after x_empty of tile i-2.
// ------------------------------------------------------------------------------------------------
constexpr int kT3MaxStages = 3;
constexpr int kT3Threads = 20 * 32;
constexpr int kT3Epilogue = 16 * 32;
constexpr int kT3Pitch = 1040;      // bytes between feature chunks (kc) of a pre-tiled plane
constexpr int kT3B2Sbo = 144;       // bytes between 8-feature groups of B2T
constexpr int kT3B2Lbo = 8 * 144;   // bytes between 4-row groups of B2T
constexpr int kT3B2Plane = 16 * kT3B2Lbo;

__host__ __device__ inline int tc3_tile_bytes(int Dp) { return 2 * (Dp / 4) * kT3Pitch; }
__host__ __device__ inline int tc3_stages(int Dp) { return Dp > 56 ? 2 : 3; }

__host__ __device__ inline int tc3_smem_layout(int Dp, int* off /*8*/) {
  int o = 0;
  off[0] = o;  // operand ring
  const int stage_bytes = (tc3_tile_bytes(Dp) + 127) / 128 * 128;
  o += tc3_stages(Dp) * stage_bytes;
  off[1] = o;  // ys[stages][64]
  o += kT3MaxStages * kT2Rows * 4;
  off[2] = o;  // A1 hi, lo
  o += 2 * (Dp / 4) * 2048;
  off[3] = o;  // B2T[2] {hi, lo}
  o += 2 * 2 * kT3B2Plane;
  off[4] = o;  // mbarriers: x_full[3], x_empty[3], s_ready[2], r_ready[2] + tmem slot
  o += 128;
  off[5] = o;  // logp combine [4][128] doubles
  o += 4 * 128 * 8;
  off[6] = stage_bytes;
  return o;
}

// X → pre-tiled {hi, lo} operand planes + padded y. grid-stride; one thread per (tile, kc, m).
__global__ void k_mc_pretile(const McArgs a, float* xt, float* yt) {
  const int kc1 = a.Dp / 4;
  const long long ntiles = (a.n_rows + kT2Rows - 1) / kT2Rows;
  const long long total = ntiles * kc1 * kT2Rows;
  const size_t plane = static_cast<size_t>(kc1) * (kT3Pitch / 4);  // floats per plane
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i % kT2Rows);
    const long long r = i / kT2Rows;
    const int kc = static_cast<int>(r % kc1);
    const long long t = r / kc1;
    const long long row = t * kT2Rows + m;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kc * 4 + e;
      const float v = (row < a.n_rows && d < a.D) ? a.X[row * a.ldx + d] : 0.0f;
      split_tf32(v, hp[e], lq[e]);
    }
    float* base = xt + static_cast<size_t>(t) * 2 * plane + static_cast<size_t>(kc) * (kT3Pitch / 4) + m * 4;
    *reinterpret_cast<float4*>(base) = h;
    *reinterpret_cast<float4*>(base + plane) = l;
    if (m < 4) {  // the 16 bytes of padding behind each 1024-byte chunk
      float* pad = xt + static_cast<size_t>(t) * 2 * plane + static_cast<size_t>(kc) * (kT3Pitch / 4) + 256 + m;
      pad[0] = 0.0f;
      pad[plane] = 0.0f;
    }
    if (kc == 0) yt[t * kT2Rows + m] = row < a.n_rows ? ld_y(a.y, a.y_dtype, row) : 0.0f;
  }
}

__global__ void __launch_bounds__(kT3Threads, 1) k_mc_pass_tc3(const McArgs a, const float* theta, int gate) {
  if (gate && !*a.need_init) return;
  extern __shared__ __align__(128) unsigned char smem[];
  int off[8];
  const int Dp = a.Dp, D = a.D;
  tc3_smem_layout(Dp, off);
  const int stage_bytes = off[6];
  const int NS = tc3_stages(Dp);
  unsigned char* ring = smem + off[0];
  float* ysm = reinterpret_cast<float*>(smem + off[1]);
  unsigned char* A1 = smem + off[2];
  unsigned char* B2T = smem + off[3];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off[4]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off[4] + 112);
  double* lpc = reinterpret_cast<double*>(smem + off[5]);
  const int kc1 = Dp / 4;
  const int a1_half = kc1 * 2048;
  const int plane_bytes = kc1 * kT3Pitch;
  const int tile_bytes = 2 * plane_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.y * kMcChainsPerCta;
  // barrier indices: 0-2 x_full, 3-5 x_empty, 6-7 s_ready, 8-9 r_ready
  const uint32_t bar0 = smem_u32(bars);
  auto XFULL = [&](int s) { return bar0 + static_cast<uint32_t>(s * 8); };
  auto XEMPTY = [&](int s) { return bar0 + static_cast<uint32_t>((3 + s) * 8); };
  auto SREADY = [&](int b) { return bar0 + static_cast<uint32_t>((6 + b) * 8); };
  auto RREADY = [&](int b) { return bar0 + static_cast<uint32_t>((8 + b) * 8); };

  if (tid == 0) {
    for (int s = 0; s < kT3MaxStages; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[3 + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[6 + b], 1);
      mbar_init(&bars[8 + b], kT3Epilogue);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = tid; i < kMcChainsPerCta * kc1; i += kT3Threads) {
    const int c = i % kMcChainsPerCta, kc = i / kMcChainsPerCta;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kc * 4 + e;
      const float v = d < D ? theta[static_cast<size_t>(cb + c) * D + d] : 0.0f;
      split_tf32(v, hp[e], lq[e]);
    }
    *reinterpret_cast<float4*>(A1 + kc * 2048 + c * 16) = h;
    *reinterpret_cast<float4*>(A1 + a1_half + kc * 2048 + c * 16) = l;
  }
  // B2T: features Dp..63 and the padding are never written by the transpose: zero once
  for (int i = tid; i < 4 * kT3B2Plane / 16; i += kT3Threads) reinterpret_cast<float4*>(B2T)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_g = tm + 256;

  long long u0, u1;
  tile_range(a.n_rows, a.n_rowgroups, blockIdx.x, u0, u1);
  const long long row_begin = u0 * kMcTileRows;
  long long row_end = u1 * kMcTileRows;
  if (row_end > a.n_rows) row_end = a.n_rows;
  const long long tile0 = row_begin / kT2Rows;  // row_begin is a multiple of 128
  const int nt = row_end > row_begin ? static_cast<int>((row_end - row_begin + kT2Rows - 1) / kT2Rows) : 0;
  double lp = 0.0;

  if (warp == 0) {
    // ================= TMA producer =================
    for (int i = 0; i < nt; ++i) {
      const int s = i % NS;
      if (i >= NS) mbar_wait_s(XEMPTY(s), ((i / NS) - 1) & 1);
      if (elect_one()) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx_s(XFULL(s), tile_bytes + kT2Rows * 4);
        bulk_g2s_s(smem_u32(ring + s * stage_bytes), a.xt + static_cast<size_t>(tile0 + i) * (tile_bytes / 4), tile_bytes, XFULL(s));
        bulk_g2s_s(smem_u32(ysm + s * kT2Rows), a.yt + (tile0 + i) * kT2Rows, kT2Rows * 4, XFULL(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA1 issuer: Sᵀ[b] = Wᵀ·X(i)ᵀ =================
    const uint64_t da_h = smem_desc(smem_u32(A1), 2048, 128), da_l = smem_desc(smem_u32(A1) + a1_half, 2048, 128);
    const uint32_t id1 = idesc_tf32(128, kT2Rows);
    const int nks = Dp / 8;
    for (int i = 0; i < nt; ++i) {
      const int s = i % NS, b = i & 1;
      mbar_wait_s(XFULL(s), (i / NS) & 1);
      if (i >= 2) mbar_wait_s(XEMPTY((i - 2) % NS), ((i - 2) / NS) & 1);  // MMA2(i-2) has read R[b]
      tc_fence_after();
      const uint32_t xs = smem_u32(ring + s * stage_bytes);
      const uint64_t db_h = smem_desc(xs, kT3Pitch, 128), db_l = smem_desc(xs + plane_bytes, kT3Pitch, 128);
      const uint32_t ts = tm + 64 * b;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < nks) {
            const uint64_t oa = static_cast<uint64_t>(ks * (4096 >> 4)), ob = static_cast<uint64_t>(ks * ((2 * kT3Pitch) >> 4));
            tc_mma_ss(ts, da_h + oa, db_h + ob, id1, ks > 0);
            tc_mma_ss(ts, da_h + oa, db_l + ob, id1, 1);
            tc_mma_ss(ts, da_l + oa, db_h + ob, id1, 1);
          }
        }
        tc_commit(SREADY(b));
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ================= MMA2 issuer: G' += R(i)·X(i), A = R from TMEM, B = B2T[b] =================
    const uint32_t id2 = idesc_tf32(128, kTcN2);
    for (int i = 0; i < nt; ++i) {
      const int s = i % NS, b = i & 1;
      mbar_wait_s(RREADY(b), (i >> 1) & 1);
      tc_fence_after();
      const uint32_t bs = smem_u32(B2T + b * 2 * kT3B2Plane);
      const uint64_t d2_h = smem_desc(bs, kT3B2Lbo, kT3B2Sbo), d2_l = smem_desc(bs + kT3B2Plane, kT3B2Lbo, kT3B2Sbo);
      const uint32_t r_h = tm + 64 * b, r_l = tm + 128 + 64 * b;
      const uint32_t first = (i > 0) ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kT2Rows / 8; ++ks) {
          const uint64_t o = static_cast<uint64_t>(ks * ((2 * kT3B2Lbo) >> 4));  // 8 rows = 2 row groups
          tc_mma_ts(tm_g, r_h + ks * 8, d2_h + o, id2, ks > 0 ? 1u : first);
          tc_mma_ts(tm_g, r_h + ks * 8, d2_l + o, id2, 1);
          tc_mma_ts(tm_g, r_l + ks * 8, d2_h + o, id2, 1);
        }
        tc_commit(XEMPTY(s));  // stage s, ys[s], S[b]/R[b] and B2T[b] are free again
      }
      __syncwarp();
    }
    if (nt > 0) mbar_wait_s(XEMPTY((nt - 1) % NS), ((nt - 1) / NS) & 1);  // the last commit covers every MMA
  } else if (warp >= 4) {
    // ================= epilogue (16 warps): transpose the landed tile, then Sᵀ → R =================
    const int q = warp & 3, cg = (warp - 4) >> 2;  // TMEM lane quadrant, column group (16 columns)
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int col = cg * 16;
    // transpose work item of this thread: plane p (hi/lo), feature chunk kc, row block mb; kc fastest across lanes
    const int et = tid - 128;
    const int npk = 2 * kc1;              // (plane, kc) pairs
    const int pk = et % npk, mbt = et / npk;  // valid when mbt < 16
    const int tp = pk / kc1, tkc = pk - tp * kc1;
    const bool has_item = mbt < 16;
    for (int i = 0; i < nt; ++i) {
      const int s = i % NS, b = i & 1;
      const long long r0 = row_begin + static_cast<long long>(i) * kT2Rows;
      const int rows = row_end - r0 >= kT2Rows ? kT2Rows : static_cast<int>(row_end - r0);
      mbar_wait_s(XFULL(s), (i / NS) & 1);
      if (i >= 2) mbar_wait_s(XEMPTY((i - 2) % NS), ((i - 2) / NS) & 1);  // MMA2(i-2) has read B2T[b]
      if (has_item) {
        const unsigned char* src = ring + s * stage_bytes + tp * plane_bytes + tkc * kT3Pitch + mbt * 64;
        float4 v0 = *reinterpret_cast<const float4*>(src);
        float4 v1 = *reinterpret_cast<const float4*>(src + 16);
        float4 v2 = *reinterpret_cast<const float4*>(src + 32);
        float4 v3 = *reinterpret_cast<const float4*>(src + 48);
        const int d0 = tkc * 4;
        unsigned char* dst = B2T + b * 2 * kT3B2Plane + tp * kT3B2Plane + mbt * kT3B2Lbo + (d0 >> 3) * kT3B2Sbo + (d0 & 7) * 16;
        *reinterpret_cast<float4*>(dst) = make_float4(v0.x, v1.x, v2.x, v3.x);
        *reinterpret_cast<float4*>(dst + 16) = make_float4(v0.y, v1.y, v2.y, v3.y);
        *reinterpret_cast<float4*>(dst + 32) = make_float4(v0.z, v1.z, v2.z, v3.z);
        *reinterpret_cast<float4*>(dst + 48) = make_float4(v0.w, v1.w, v2.w, v3.w);
      }
      fence_proxy_async_smem();
      mbar_wait_s(SREADY(b), (i >> 1) & 1);
      tc_fence_after();
      const float* ys = ysm + s * kT2Rows;
      uint32_t v[16], vh[16], vl[16];
      tmem_ld16(tm + 64 * b + lane_base + col, v);
      tmem_wait_ld();
      if (a.want_logp || a.family != 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = col + j;
          float lpv, rv;
          row_terms(a.family, __uint_as_float(v[j]), ys[m], a.lik_scale, lpv, rv);
          if (m >= rows) {
            lpv = 0.0f;
            rv = 0.0f;
          }
          lp += static_cast<double>(lpv);
          float h, l;
          split_tf32(rv, h, l);
          vh[j] = __float_as_uint(h);
          vl[j] = __float_as_uint(l);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = col + j;
          const float eta = __uint_as_float(v[j]);
          const float e = expf(-fabsf(eta));
          const float qv = __fdividef(e, 1.0f + e);
          const float yv = ys[m];
          float rv = eta >= 0.0f ? (yv - 1.0f) + qv : yv - qv;
          if (m >= rows) rv = 0.0f;
          float h, l;
          split_tf32(rv, h, l);
          vh[j] = __float_as_uint(h);
          vl[j] = __float_as_uint(l);
        }
      }
      tmem_st16(tm + 64 * b + lane_base + col, vh);
      tmem_st16(tm + 128 + 64 * b + lane_base + col, vl);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bars[8 + b]);  // R[b] and B2T[b] are ready
    }
  }

  // ---- write this row group's partial sums ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= 4) {
    const int q = warp & 3, cg = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int chain = cb + 32 * q + lane;
    float* pg = a.part_g + (static_cast<size_t>(blockIdx.x) * a.C + chain) * Dp;
    const int col = cg * 16;
    uint32_t v[16];
    if (nt > 0) {
      tmem_ld16(tm_g + lane_base + col, v);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0u;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col + j < Dp) pg[col + j] = __uint_as_float(v[j]);
    lpc[cg * 128 + 32 * q + lane] = lp;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < kMcChainsPerCta)
    a.part_lp[static_cast<size_t>(blockIdx.x) * a.C + cb + tid] = (lpc[tid] + lpc[128 + tid]) + (lpc[256 + tid] + lpc[384 + tid]);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// per-chain kernels: grid = C blocks of 64 threads (thread d = feature d)
// ------------------------------------------------------------------------------------------------
constexpr int kMcChainThreads = 64;

__device__ __forceinline__ double chain_sum(double v, double* sh) {  // 64 threads, fixed order
  v = warp_sum_f64(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1];
}

// sums the row-group partials of (chain c, feature d) in a fixed order, float64
__device__ __forceinline__ double sum_part_g(const McArgs& a, int c, int d) {
  double s = 0.0;
  for (int rg = 0; rg < a.n_rowgroups; ++rg) s += static_cast<double>(a.part_g[(static_cast<size_t>(rg) * a.C + c) * a.Dp + d]);
  return s;
}
__device__ __forceinline__ double sum_part_lp(const McArgs& a, int c) {
  double s = 0.0;
  for (int rg = 0; rg < a.n_rowgroups; ++rg) s += a.part_lp[static_cast<size_t>(rg) * a.C + c];
  return s;
}

// gradient and log joint of chain c at `pos` from the pass partials + Normal prior (hmc.py:183-190)
__device__ __forceinline__ double mc_finish_gradient(const McArgs& a, int c, const float* pos, float* gout, double* sh) {
  const int d = threadIdx.x;
  double pl = 0.0;
  if (d < a.D) {
    const float loc = a.prior_loc[d], sc = a.prior_scale[d];
    const float zc = pos[static_cast<size_t>(c) * a.D + d];
    gout[static_cast<size_t>(c) * a.D + d] = static_cast<float>(sum_part_g(a, c, d) + prior_grad(zc, loc, sc));
    pl = prior_quad(zc, loc, sc);
  }
  return (chain_sum(pl, sh) - a.prior_const) + sum_part_lp(a, c);
}

__global__ void k_mc_check(const McArgs a) {  // one block of 256 threads
  __shared__ int s_need;
  if (threadIdx.x == 0) s_need = !*a.valid;
  __syncthreads();
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  const size_t n = static_cast<size_t>(a.C) * a.D;
  bool mismatch = false;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x)
    if (__float_as_uint(a.params[t_prev * n + i]) != __float_as_uint(a.zcur[i])) mismatch = true;
  if (mismatch) s_need = 1;
  __syncthreads();
  if (s_need)
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) a.zcur[i] = a.params[t_prev * n + i];
  if (threadIdx.x == 0) *a.need_init = s_need;
}

__global__ void __launch_bounds__(kMcChainThreads) k_mc_init_finish(const McArgs a) {
  __shared__ double sh[2];
  if (!*a.need_init) return;
  const int c = blockIdx.x;
  const double lp = mc_finish_gradient(a, c, a.zcur, a.gcur, sh);
  if (threadIdx.x == 0) a.logp_cur[c] = lp;
  if (c == 0 && threadIdx.x == 0) *a.valid = 1;
}

__device__ __forceinline__ float mc_kick(float r, float h, float g) { return __fadd_rn(r, __fmul_rn(h, g)); }
__device__ __forceinline__ float mc_drift(float z, float e, float r) { return __fadd_rn(z, __fmul_rn(e, r)); }

__device__ __forceinline__ void mc_finish_transition(const McArgs& a, int c, long long it, double logp_new, double* sh) {
  const int d = threadIdx.x;
  const size_t cd = static_cast<size_t>(c) * a.D + d;
  const float rr = d < a.D ? a.r[cd] : 0.0f;
  const double k_new = 0.5 * chain_sum(static_cast<double>(__fmul_rn(rr, rr)), sh);
  const double logp_cur = a.logp_cur[c], k_old = a.k_old[c], log_u = a.log_u[c];
  const double ratio = ((k_old - k_new) + logp_new) - logp_cur;  // hmc.py:100-105
  const bool accept = log_u < ratio;                              // hmc.py:108-109
  __syncthreads();
  const long long t = a.t0 + it;
  if (d < a.D) {
    if (accept) {
      a.zcur[cd] = a.z[cd];
      a.gcur[cd] = a.g[cd];
    }
    a.params[(static_cast<size_t>(t) * a.C + c) * a.D + d] = accept ? a.z[cd] : a.zcur[cd];
  }
  if (d == 0) {
    if (accept) {
      a.logp_cur[c] = logp_new;
      a.n_accept[c] += 1;
    }
    if (a.trace) {
      double* tr = a.trace + (static_cast<size_t>(it) * a.C + c) * 8;
      tr[0] = logp_cur;
      tr[1] = logp_new;
      tr[2] = k_old;
      tr[3] = k_new;
      tr[4] = ratio;
      tr[5] = log_u;
      tr[6] = accept ? 1.0 : 0.0;
      tr[7] = 0.0;
    }
  }
}

// start of transition `it` for every chain: momentum, kinetic energy, first half kick + drift (hmc.py:88-97,201-204)
__global__ void __launch_bounds__(kMcChainThreads) k_mc_begin(const McArgs a, long long it) {
  __shared__ double sh[2];
  const int c = blockIdx.x, d = threadIdx.x;
  const long long t = a.t0 + it;
  const size_t cd = static_cast<size_t>(c) * a.D + d;
  float rv = 0.0f;
  if (d < a.D) {
    rv = a.r0 ? a.r0[(static_cast<size_t>(it) * a.C + c) * a.D + d] : philox_normal(a.seed + 0x9E3779B97F4A7C15ull * (c + 1), t, d);
    float zz = a.zcur[cd];
    float rr = rv;
    const float gg = a.gcur[cd];
    if (a.L > 0) {
      rr = mc_kick(rv, a.half_eps, gg);
      zz = mc_drift(zz, a.eps, rr);
    }
    a.r[cd] = rr;
    a.z[cd] = zz;
    a.g[cd] = gg;
  }
  const double k_old = 0.5 * chain_sum(static_cast<double>(__fmul_rn(rv, rv)), sh);
  if (d == 0) {
    const float u = a.u ? a.u[static_cast<size_t>(it) * a.C + c] : philox_uniform(a.seed + 0x9E3779B97F4A7C15ull * (c + 1), t);
    a.k_old[c] = k_old;
    a.log_u[c] = static_cast<double>(logf(u));
  }
  if (a.L == 0) {
    __syncthreads();
    mc_finish_transition(a, c, it, a.logp_cur[c], sh);
  }
}

// after the pass of leapfrog step s: second half kick, then the next step's first half or the MH accept
__global__ void __launch_bounds__(kMcChainThreads) k_mc_leap(const McArgs a, long long it, int s) {
  __shared__ double sh[2];
  const int c = blockIdx.x, d = threadIdx.x;
  const size_t cd = static_cast<size_t>(c) * a.D + d;
  const bool last = (s == a.L - 1);
  double logp_new = 0.0;
  if (last) {
    logp_new = mc_finish_gradient(a, c, a.z, a.g, sh);
  } else if (d < a.D) {
    a.g[cd] = static_cast<float>(sum_part_g(a, c, d) + prior_grad(a.z[cd], a.prior_loc[d], a.prior_scale[d]));
  }
  if (d < a.D) {
    const float gg = a.g[cd];
    float rr = mc_kick(a.r[cd], a.half_eps, gg);
    if (!last) {
      rr = mc_kick(rr, a.half_eps, gg);
      a.z[cd] = mc_drift(a.z[cd], a.eps, rr);
    }
    a.r[cd] = rr;
  }
  if (last) {
    __syncthreads();
    mc_finish_transition(a, c, it, logp_new, sh);
  }
}

__global__ void __launch_bounds__(kMcChainThreads) k_mc_logp_grad_finish(const McArgs a, const float* theta, double* logp,
                                                                         float* grad) {
  __shared__ double sh[2];
  const int c = blockIdx.x;
  const double lp = mc_finish_gradient(a, c, theta, grad, sh);
  if (threadIdx.x == 0) logp[c] = lp;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int mc_smem_bytes_tc(int Dp) {
  int off[8];
  return tc_smem_layout(Dp, kMcMaxD, off);
}

size_t mc_pretile_bytes(long long n_rows, int Dp, size_t* yt_bytes) {
  const long long ntiles = (n_rows + kT2Rows - 1) / kT2Rows;
  *yt_bytes = static_cast<size_t>(ntiles) * kT2Rows * sizeof(float);
  return static_cast<size_t>(ntiles) * tc3_tile_bytes(Dp);
}

cudaError_t mc_launch_pretile(const McArgs& a, float* xt, float* yt, cudaStream_t s) {
  k_mc_pretile<<<148 * 8, 256, 0, s>>>(a, xt, yt);
  return cudaGetLastError();
}

cudaError_t mc_prepare_tc() {
  int off[8];
  {
    int mx = 0;
    for (int dp = 8; dp <= kMcMaxD; dp += 8) {
      const int b = tc3_smem_layout(dp, off);
      if (b > mx) mx = b;
    }
    cudaError_t e3 = cudaFuncSetAttribute(k_mc_pass_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    if (e3 != cudaSuccess) return e3;
  }
  cudaError_t e = cudaFuncSetAttribute(k_mc_pass_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       tc2_smem_layout(kMcMaxD, kMcMaxD, off));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_mc_pass_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem_bytes_tc(kMcMaxD));
}

cudaError_t mc_launch_pass(const McArgs& a, const float* theta, int use_tc, int gate, cudaStream_t s) {
  dim3 grid(a.n_rowgroups, a.C / kMcChainsPerCta);
  if (use_tc == 3) {
    int off[8];
    const int smem = tc3_smem_layout(a.Dp, off);
    k_mc_pass_tc3<<<grid, kT3Threads, smem, s>>>(a, theta, gate);
  } else if (use_tc == 2) {
    int off[8];
    const int smem = tc2_smem_layout(a.Dp, static_cast<int>(a.ldx), off);
    k_mc_pass_tc2<<<grid, kT2Threads, smem, s>>>(a, theta, gate);
  } else if (use_tc == 1) {
    int off[8];
    const int smem = tc_smem_layout(a.Dp, static_cast<int>(a.ldx), off);
    k_mc_pass_tc<<<grid, kTcThreads, smem, s>>>(a, theta, gate);
  } else {
    k_mc_pass_simple<<<grid, kMcChainsPerCta, 0, s>>>(a, theta, gate);
  }
  return cudaGetLastError();
}
cudaError_t mc_launch_check(const McArgs& a, cudaStream_t s) {
  k_mc_check<<<1, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t mc_launch_init_finish(const McArgs& a, cudaStream_t s) {
  k_mc_init_finish<<<a.C, kMcChainThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t mc_launch_begin(const McArgs& a, long long it, cudaStream_t s) {
  k_mc_begin<<<a.C, kMcChainThreads, 0, s>>>(a, it);
  return cudaGetLastError();
}
cudaError_t mc_launch_leap(const McArgs& a, long long it, int step, cudaStream_t s) {
  k_mc_leap<<<a.C, kMcChainThreads, 0, s>>>(a, it, step);
  return cudaGetLastError();
}
cudaError_t mc_launch_logp_grad_finish(const McArgs& a, const float* theta, double* logp, float* grad, cudaStream_t s) {
  k_mc_logp_grad_finish<<<a.C, kMcChainThreads, 0, s>>>(a, theta, logp, grad);
  return cudaGetLastError();
}

}  // namespace edhmc
