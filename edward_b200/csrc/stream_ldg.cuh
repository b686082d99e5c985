// stream_ldg.cuh — the streaming pass WITHOUT shared-memory staging (ring mode 2), for narrow rows (D <= 64).
//
// Ring mode 1 moves every byte of X through shared memory twice (TMA write, then LDS): at cfg 2 that is 14k of the
// 22k cycles of a pass on the 128 B/clk shared-memory port, and neither more warps nor more lanes per row helped
// (profiles/README round 2). Here X is re-laid ONCE at bind time (k_relay_tiles) into column-pair-major 32-row tiles,
//
//     Xt[tile][pair i < Kact][row r < 32] = float2(X[32*tile + r][2i], X[32*tile + r][2i+1])     (zero padded)
//     Yt[tile][r]                         = float(y[32*tile + r])
//
// so that lane r of a warp reads ITS row with Kact coalesced LDG.64 (256 contiguous bytes per warp instruction) straight
// into registers: one lane per row as before — the same dot / link / gradient arithmetic as row_group<1, 2, K> — but no
// TMA, no mbarrier, no ring bookkeeping, and the bytes cross the L1/shared-memory array once. theta is read from shared
// memory (broadcast LDS.64), which leaves room for 12 warps per SM; the probe kernel of this access pattern
// (edhmc_probe_read modes 2-5, tools/probe_gemv.py) streams cfg 2 in 8.3 us per pass with 12 warps against 10.2 us with 8.
//
// Rows: a CTA owns a contiguous range of tiles, warp w of it the tiles w, w + NW, ...; odd passes walk them backwards
// (zig-zag, L2 reuse). Padded rows of the last tile are masked (their residual and log-likelihood are zero).
//
// On the persistent plan the pass needs neither shared memory for staging nor the SM's tensor memory, so as much of X as
// fits is copied there ONCE per launch (ldg_load_resident) and read from on-chip memory in every pass: the first tiles of
// the CTA's range from shared memory (LDS.64), the last ones from tensor memory (tcgen05.ld.32x32b straight into the tile
// registers, TmemTiles / ldg_tmem_tiles). cfg 2: 29 + 36 of 123 tiles per SM, 53 % of X; 16.1 -> 14.2 us per leapfrog step.
#pragma once
#include "stream.cuh"
#include "tc.cuh"

namespace edhmc {

// Re-lay of X / y into the tile layout above (once per edhmc_bind_data). One thread per (tile, pair, row) element.
static __global__ void __launch_bounds__(256) k_relay_tiles(const float* __restrict__ X, long long n_rows, long long ldx, int D,
                                                     const void* __restrict__ y, int y_dtype, int Kact, long long n_tiles,
                                                     float2* __restrict__ Xt, float* __restrict__ Yt) {
  const long long total = n_tiles * Kact * 32;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e & 31);
    const long long q = e >> 5;
    const int i = static_cast<int>(q % Kact);
    const long long t = q / Kact;
    const long long row = t * 32 + r;
    float2 v = make_float2(0.0f, 0.0f);
    if (row < n_rows) {
      const float* src = X + row * ldx + 2 * i;
      v.x = src[0];  // 2i < D always (Kact = ceil(D / 2))
      if (2 * i + 1 < D) v.y = src[1];
    }
    Xt[e] = v;
    if (i == 0) {
      float yv = 0.0f;
      if (row < n_rows)
        yv = y_dtype == 1 ? reinterpret_cast<const float*>(y)[row] : static_cast<float>(reinterpret_cast<const int*>(y)[row]);
      Yt[row] = yv;
    }
  }
}

// Tiles a warp keeps in flight per round: rows of few columns are batched so that every lane has ~24 LDG.64 outstanding
// whatever D is (a warp with one 8-column tile in flight has 1 KB outstanding and is bound by the L2 / HBM latency:
// 16M x 8 ran at 47 % of the HBM peak with J = 1).
template <int K>
struct LdgBatch {
  static constexpr int J = K >= 14 ? 1 : (K >= 10 ? 2 : (K >= 8 ? 3 : (K >= 6 ? 4 : 6)));
};

// Measured and rejected (profiles/README round 2): two tiles per warp double-buffered in registers (8 warps for K >= 24, 12
// below) — slower on every shape (4M x 32: 100.5 -> 86.7 % of the HBM copy peak, 3M x 64: 103 -> 57 %, cfg 2 18.6 us per step
// against 17.6 for the loop below with resident tiles). A conditionally loaded register array ends up in local memory; with
// unconditional (clamped) loads the arrays stay in registers but the 255-register variants lose more to spilled chain state
// and 8 warps than the second buffer gains.
// Tensor-memory tiles. The sampler issues no tcgen05.mma, so the SM's 256 KB of tensor memory (128 lanes x 512 columns of
// 32 bits) is free: a persistent launch parks further 32-row tiles there — TMEM lane = row of the tile, 2K consecutive
// columns = the row — and reads a row back with tcgen05.ld.32x32b straight into the registers the arithmetic uses. A warp
// can only reach the 32 lanes of its quarter (warp % 4), so TMEM tile tt lives in quarter tt % 4, slot tt / 4, and is
// handled, in every pass and in both directions, by the warps of that quarter (slot % (NW / 4) == warp / 4); y of those
// tiles sits in shared memory behind the resident tiles. cfg 2: 36 tiles (1,152 rows, 249 KB) per SM next to the 29 in
// shared memory — 53 % of X never leaves the SM. Only for K >= 14 (one tile per warp and round).
template <int K>
struct TmemTiles {
  static constexpr int kCols = (2 * K + 7) / 8 * 8;  // column stride of a slot
  static constexpr int kSlots = 512 / kCols;         // slots per lane quarter
  static constexpr int kMax = LdgBatch<K>::J == 1 ? 4 * kSlots : 0;
};

template <int K, int FAM, int LPM>
__device__ __forceinline__ void ldg_row_math(const float2 (&x)[K], float yv, bool valid_row, const float2* __restrict__ theta2,
                                             float bias, float lik_scale, float2 (&g)[K], float& gb, double& lp) {
  float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
  const float4* theta4 = reinterpret_cast<const float4*>(theta2);
#pragma unroll
  for (int i = 0; i < K; i += 2) {
    const float4 w = theta4[i >> 1];  // broadcast LDS.128: two pairs per shared-memory wavefront (entries >= D are zero)
    a0 = fma2(x[i], make_float2(w.x, w.y), a0);
    if (i + 1 < K) a1 = fma2(x[i + 1], make_float2(w.z, w.w), a1);
  }
  const float eta = ((a0.x + a0.y) + (a1.x + a1.y)) + bias;
  float lpv = 0.0f, rv;
  if (LPM)
    row_terms(FAM, eta, yv, lik_scale, lpv, rv);
  else if (FAM == 0)
    rv = bernoulli_resid_fast(eta, yv);
  else
    rv = row_resid(FAM, eta, yv, lik_scale);
  if (!valid_row) {  // padded rows of the last tile
    lpv = 0.0f;
    rv = 0.0f;
  }
  if (LPM) lp += static_cast<double>(lpv);
  gb += rv;
  const float2 r2 = make_float2(rv, rv);
#pragma unroll
  for (int i = 0; i < K; ++i) g[i] = fma2(r2, x[i], g[i]);
}

// The tiles of the CTA's range that a launch with tensor-memory tiles keeps there: the LAST n_tm of its cnt tiles. The
// same tiles are visited by the same warps in the same order when they are read from global memory instead (tm_base ==
// kNoTmem: stepwise plan, one pass per launch), so both plans add the same numbers in the same order.
constexpr uint32_t kNoTmem = 0xFFFFFFFFu;
template <int K, int NW, int FAM, int LPM>
__device__ __forceinline__ void ldg_tmem_tiles(const float2* __restrict__ Xt, const float* __restrict__ Yt, int Kact, long long tile0,
                                               int n_tm, uint32_t tm_base, const float* __restrict__ tm_y, long long n_rows,
                                               const float2* __restrict__ theta2, float bias, float lik_scale, float2 (&g)[K],
                                               float& gb, double& lp) {
  if constexpr (TmemTiles<K>::kMax > 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3;
    const long long stride_t = static_cast<long long>(Kact) * 32;
    for (int slot = warp >> 2; 4 * slot + q < n_tm; slot += NW / 4) {
      const int tt = 4 * slot + q;
      float2 x[K];
      float yv;
      if (tm_base != kNoTmem) {
        uint32_t v[2 * K];
        tmem_ld_cols<2 * K>(tm_base + (static_cast<uint32_t>(q * 32) << 16) + slot * TmemTiles<K>::kCols, v);
        yv = tm_y[tt * 32 + lane];
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < K; ++i) x[i] = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      } else {
        const float2* p = Xt + (tile0 + tt) * stride_t + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) x[i] = (i < Kact) ? __ldcg(p + i * 32) : make_float2(0.0f, 0.0f);
        yv = __ldcg(Yt + (tile0 + tt) * 32 + lane);
      }
      ldg_row_math<K, FAM, LPM>(x, yv, (tile0 + tt) * 32 + lane < n_rows, theta2, bias, lik_scale, g, gb, lp);
    }
  }
}

template <int K, int NW, int FAM, int LPM>
__device__ __forceinline__ void stream_pass_ldg_tiles(const float2* __restrict__ Xt, const float* __restrict__ Yt, int Kact,
                                                      long long t0, long long cnt, bool backward, long long n_rows,
                                                      const float2* __restrict__ theta2, float bias, float lik_scale,
                                                      const float* __restrict__ res, int n_res, int n_tm, uint32_t tm_base,
                                                      const float* __restrict__ tm_y,
                                                      float2* __restrict__ gout, float* gbout, double* lpout) {
  constexpr int J = LdgBatch<K>::J;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 g[K];
#pragma unroll
  for (int i = 0; i < K; ++i) g[i] = make_float2(0.0f, 0.0f);
  float gb = 0.0f;
  double lp = 0.0;
  const long long stride_t = static_cast<long long>(Kact) * 32;
  for (long long k0 = static_cast<long long>(warp) * J; k0 < cnt; k0 += NW * J) {
    float2 x[J][K];
    float yv[J];
    long long row[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const bool live = J == 1 || k0 + j < cnt;
      const long long t = backward ? (t0 + cnt - 1 - (k0 + j)) : (t0 + k0 + j);
      const long long rel = t - t0;
      if (live && rel < n_res) {  // resident tile (warp-uniform): shared memory, conflict-free LDS.64
        const float2* p = reinterpret_cast<const float2*>(res) + rel * stride_t + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) x[j][i] = (i < Kact) ? p[i * 32] : make_float2(0.0f, 0.0f);
        yv[j] = res[static_cast<long long>(n_res) * stride_t * 2 + rel * 32 + lane];
      } else {
        const float2* p = Xt + t * stride_t + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) x[j][i] = (live && i < Kact) ? __ldcg(p + i * 32) : make_float2(0.0f, 0.0f);
        yv[j] = live ? __ldcg(Yt + t * 32 + lane) : 0.0f;
      }
      row[j] = live ? t * 32 + lane : n_rows;
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
      const float4* theta4 = reinterpret_cast<const float4*>(theta2);
#pragma unroll
      for (int i = 0; i < K; i += 2) {
        const float4 w = theta4[i >> 1];  // broadcast LDS.128: two pairs per shared-memory wavefront (entries >= D are zero)
        a0 = fma2(x[j][i], make_float2(w.x, w.y), a0);
        if (i + 1 < K) a1 = fma2(x[j][i + 1], make_float2(w.z, w.w), a1);
      }
      const float eta = ((a0.x + a0.y) + (a1.x + a1.y)) + bias;
      float lpv = 0.0f, rv;
      if (LPM)
        row_terms(FAM, eta, yv[j], lik_scale, lpv, rv);
      else if (FAM == 0)
        rv = bernoulli_resid_fast(eta, yv[j]);
      else
        rv = row_resid(FAM, eta, yv[j], lik_scale);
      if (row[j] >= n_rows) {  // padded rows of the last tile, absent tiles of the last round
        lpv = 0.0f;
        rv = 0.0f;
      }
      if (LPM) lp += static_cast<double>(lpv);
      gb += rv;
      const float2 r2 = make_float2(rv, rv);
#pragma unroll
      for (int i = 0; i < K; ++i) g[i] = fma2(r2, x[j][i], g[i]);
    }
  }
  // `cnt` above excludes the tensor-memory tiles, which follow it in the CTA's range
  ldg_tmem_tiles<K, NW, FAM, LPM>(Xt, Yt, Kact, t0 + cnt, n_tm, tm_base, tm_y, n_rows, theta2, bias, lik_scale, g, gb, lp);
#pragma unroll
  for (int i = 0; i < K; ++i) gout[i] = g[i];
  *gbout = gb;
  *lpout = lp;
}

// Resident tiles: X does not change during a persistent launch, so part of the CTA's range is copied ONCE per launch into
// on-chip memory that ring mode 2 does not otherwise need and is read from there in every pass: the first n_res tiles into
// shared memory (~210 KB per SM), the last n_tm tiles into tensor memory (256 KB per SM, see TmemTiles). Those rows never
// cross the SM's L2 port again, and an L2-sized X (cfg 2: 127.8 MB against 126 MB of L2) shrinks to a working set that
// fits the L2 with room to spare. Shared-memory layout: [n_res][Kact][32] float2, [n_res][32] float (y), [n_tm][32] float
// (y of the tensor-memory tiles). The order in which a warp visits its tiles does not depend on where a tile lives, so
// the sums are bit-identical to a launch that reads everything from global memory (stepwise plan).
struct LdgRange {
  long long t0;    // first tile of the CTA
  long long cnt;   // tiles visited by the main loop (shared-memory-resident ones first)
  int n_res, n_tm;
};
template <int K>
__device__ __forceinline__ LdgRange ldg_range(const KArgs& a) {
  LdgRange r;
  const long long per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  r.t0 = blockIdx.x * per;
  long long cnt = a.n_tiles - r.t0;
  if (cnt > per) cnt = per;
  if (cnt < 0) cnt = 0;
  const long long tm_cap = TmemTiles<K>::kMax < a.n_tm ? TmemTiles<K>::kMax : a.n_tm;
  r.n_tm = static_cast<int>(cnt < tm_cap ? cnt : tm_cap);
  r.cnt = cnt - r.n_tm;
  r.n_res = static_cast<int>(r.cnt < a.n_res ? r.cnt : a.n_res);
  return r;
}

template <int K, int NW>
__device__ __forceinline__ void ldg_load_resident(const KArgs& a, const SmemLayout& sm, uint32_t tm_base) {
  if (a.n_res <= 0 && tm_base == kNoTmem) return;
  const LdgRange r = ldg_range<K>(a);
  const long long stride_t = static_cast<long long>(a.Kact) * 32;  // float2 per tile
  if (a.n_res > 0) {
    const float4* src = reinterpret_cast<const float4*>(a.Xt + r.t0 * stride_t);  // tiles start 256-byte aligned
    float4* dst = reinterpret_cast<float4*>(sm.ring);
    const long long n4 = static_cast<long long>(r.n_res) * stride_t / 2;
    for (long long i = threadIdx.x; i < n4; i += NW * 32) dst[i] = __ldcg(src + i);
    float* dy = sm.ring + static_cast<long long>(r.n_res) * stride_t * 2;
    for (int i = threadIdx.x; i < r.n_res * 32; i += NW * 32) dy[i] = __ldcg(a.Yt + r.t0 * 32 + i);
  }
  if constexpr (TmemTiles<K>::kMax > 0) {
    if (tm_base != kNoTmem) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int q = warp & 3;
      float* tm_y = sm.ring + static_cast<long long>(r.n_res) * (stride_t * 2 + 32);
      const long long tile0 = r.t0 + r.cnt;
      for (int slot = warp >> 2; 4 * slot + q < r.n_tm; slot += NW / 4) {
        const int tt = 4 * slot + q;
        const float2* p = a.Xt + (tile0 + tt) * stride_t + lane;
        uint32_t v[2 * K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float2 xv = (i < a.Kact) ? __ldcg(p + i * 32) : make_float2(0.0f, 0.0f);
          v[2 * i] = __float_as_uint(xv.x);
          v[2 * i + 1] = __float_as_uint(xv.y);
        }
        tmem_st_cols<2 * K>(tm_base + (static_cast<uint32_t>(q * 32) << 16) + slot * TmemTiles<K>::kCols, v);
        tm_y[tt * 32 + lane] = __ldcg(a.Yt + (tile0 + tt) * 32 + lane);
      }
      tmem_wait_st();  // a warp only ever reads back the tiles it stored itself
    }
  }
  __syncthreads();
}

// One pass of this CTA over its tiles (ring mode 2). Same contract as stream_pass / stream_pass_cta: on return
// cta_acc[0..P] holds the CTA's float64 sums, reduced in a fixed order; ends with a __syncthreads().
template <int K, int NW>
__device__ __forceinline__ void stream_pass_ldg(const KArgs& a, const SmemLayout& sm, float bias, bool want_lp, bool backward,
                                                uint32_t tm_base) {
  const LdgRange r = ldg_range<K>(a);
  const float2* theta2 = reinterpret_cast<const float2*>(sm.theta_s);
  const float* res = sm.ring;
  const float* tm_y = sm.ring + static_cast<long long>(r.n_res) * (static_cast<long long>(a.Kact) * 64 + 32);
  float2 g[K];
  float gb;
  double lp;
  const int fam = a.family;
#define EDHMC_LDG_ARGS a.Xt, a.Yt, a.Kact, r.t0, r.cnt, backward, a.n_rows, theta2, bias, a.lik_scale, res, r.n_res, r.n_tm, tm_base, tm_y, g, &gb, &lp
  if (want_lp) {
    if (fam == 0)
      stream_pass_ldg_tiles<K, NW, 0, 1>(EDHMC_LDG_ARGS);
    else if (fam == 1)
      stream_pass_ldg_tiles<K, NW, 1, 1>(EDHMC_LDG_ARGS);
    else
      stream_pass_ldg_tiles<K, NW, 2, 1>(EDHMC_LDG_ARGS);
  } else {
    if (fam == 0)
      stream_pass_ldg_tiles<K, NW, 0, 0>(EDHMC_LDG_ARGS);
    else if (fam == 1)
      stream_pass_ldg_tiles<K, NW, 1, 0>(EDHMC_LDG_ARGS);
    else
      stream_pass_ldg_tiles<K, NW, 2, 0>(EDHMC_LDG_ARGS);
  }
#undef EDHMC_LDG_ARGS
  PlanRegs pr_unused = {};
  int wt_unused = 0, ring_unused = 0;
  pass_reduce<1, 2, K, NW, true>(a, pr_unused, wt_unused, ring_unused, sm, g, gb, lp, false, false, 0u, 0ull);
}

}  // namespace edhmc
