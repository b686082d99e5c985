// chains_wide.cu — many vectorised chains for wide models / row shards (see chains_wide.cuh): two hand-written
// tcgen05 GEMMs in 3xTF32 per leapfrog step, operands as pre-tiled hi/lo planes moved by TMA bulk copies.
//
// k_mcw_gemm<1>   Sᵀ[128 chains x 128 rows] per tile: A = Wᵀ planes, B = X planes (K = features, 16 per stage);
//                 TWO accumulators per tile — hi·hi into the main one, the 3xTF32 correction terms hi·lo + lo·hi into a
//                 second one, summed in the epilogue — because the tensor core truncates the fp32 accumulator after
//                 every instruction: with all three terms in one accumulator K = 1,000 features are 375 truncations
//                 of the logit (a systematic 8e-6 shrink, 1e-5 gradient error at 1.25M rows); now 125;
//                 both double-buffered in TMEM (2 x (128 + 128) columns); epilogue R = y − σ(Sᵀ) (one thread per
//                 chain = TMEM lane, so the per-chain log-likelihood needs no cross-thread reduction), written as
//                 hi/lo planes in the K-major layout GEMM 2 reads as its A operand.
// k_mcw_gemm<2>   G'[128 chains x NB2 features] per (chain tile, feature tile, K split): A = R planes, B = Xᵀ planes
//                 (K = rows, 16 per stage); every 8,192 rows the fp32 TMEM accumulator is flushed into a float64
//                 partial (bounds the fp32 accumulation error; the other TMEM buffer keeps the MMAs running).
// Both: 12 warps — TMA producer, MMA issuer (elect.sync in a converged warp), 8 epilogue warps; 4-stage ring of
// {A hi, A lo, B hi, B lo}; mbarrier chain full → (tcgen05.commit) empty, acc_full → acc_empty.
#include "chains_wide.cuh"

#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace edhmc {

constexpr int kMwThreads = 384;
constexpr int kMwStages = 4;
constexpr int kMwAPlane = (kMwKC / 4) * 128 * 16;     // 8 KB
constexpr int kMwBPlaneMax = (kMwKC / 4) * 256 * 16;  // 16 KB
constexpr int kMwStageBytes = 2 * kMwAPlane + 2 * kMwBPlaneMax;
constexpr int kMwEpilogue = 256;
constexpr int kMwSmemBytes = kMwStages * kMwStageBytes + 128 + 2 * 128 * 8;
constexpr int kMwChainThreads = 256;

__device__ __forceinline__ float mcw_ld_y(const void* y, int y_dtype, long long i) {
  return y_dtype == 0 ? static_cast<float>(reinterpret_cast<const int*>(y)[i]) : reinterpret_cast<const float*>(y)[i];
}

// ------------------------------------------------------------------------------------------------
// operand planes
// ------------------------------------------------------------------------------------------------
// X as the B operand of GEMM 1: per 256-row tile {hi, lo} x [Kp1/4][256][4]. One thread per (tile, row, k-group),
// k-group fastest: coalesced reads along the rows of X.
__global__ void k_mcw_pretile_xk(const McwArgs a) {
  const int kg1 = a.Kp1 / 4;
  const long long total = static_cast<long long>(a.nrt) * kMwRowTile * kg1;
  const size_t plane = static_cast<size_t>(a.Kp1) * kMwRowTile;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kg = static_cast<int>(i % kg1);
    const long long rr = i / kg1;
    const int m = static_cast<int>(rr % kMwRowTile);
    const long long t = rr / kMwRowTile;
    const long long row = t * kMwRowTile + m;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kg * 4 + e;
      const float v = (row < a.n_rows && d < a.D) ? (d < a.Dx ? a.X[row * a.ldx + d] : 1.0f) : 0.0f;  // d == Dx: bias column
      split_tf32(v, hp[e], lq[e]);
    }
    float* base = a.xk + static_cast<size_t>(t) * 2 * plane + (static_cast<size_t>(kg) * kMwRowTile + m) * 4;
    *reinterpret_cast<float4*>(base) = h;
    *reinterpret_cast<float4*>(base + plane) = l;
    if (kg == 0) a.yt[row] = row < a.n_rows ? mcw_ld_y(a.y, a.y_dtype, row) : 0.0f;
  }
}

// Xᵀ as the B operand of GEMM 2: per feature tile {hi, lo} x [rowsP/4][NB2][4] (K = rows). One thread per
// (feature tile, row group, feature), feature fastest: coalesced reads and writes.
__global__ void k_mcw_pretile_xt(const McwArgs a) {
  const long long rg_n = a.rowsP / 4;
  const long long total = static_cast<long long>(a.nft) * rg_n * a.NB2;
  const size_t plane = static_cast<size_t>(a.rowsP) * a.NB2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i % a.NB2);
    const long long rr = i / a.NB2;
    const long long rg = rr % rg_n;
    const int ft = static_cast<int>(rr / rg_n);
    const int d = ft * a.NB2 + f;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long row = rg * 4 + e;
      const float v = (row < a.n_rows && d < a.D) ? (d < a.Dx ? a.X[row * a.ldx + d] : 1.0f) : 0.0f;  // d == Dx: bias column
      split_tf32(v, hp[e], lq[e]);
    }
    float* base = a.xt + static_cast<size_t>(ft) * 2 * plane + (static_cast<size_t>(rg) * a.NB2 + f) * 4;
    *reinterpret_cast<float4*>(base) = h;
    *reinterpret_cast<float4*>(base + plane) = l;
  }
}

// Wᵀ as the A operand of GEMM 1: per chain tile {hi, lo} x [Kp1/4][128][4], from theta [C][D].
__global__ void k_mcw_wtile(const McwArgs a, const float* theta, int gate) {
  if (gate && !*a.need_init) return;
  const int kg1 = a.Kp1 / 4;
  const int total = a.nct * kg1 * 128;
  const size_t plane = static_cast<size_t>(a.Kp1) * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kg = i % kg1;
    const int rr = i / kg1;
    const int c = rr % 128;
    const int ct = rr / 128;
    float4 h, l;
    float* hp = &h.x;
    float* lq = &l.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = kg * 4 + e;
      const float v = d < a.D ? theta[static_cast<size_t>(ct * 128 + c) * a.D + d] : 0.0f;
      split_tf32(v, hp[e], lq[e]);
    }
    float* base = a.wt + static_cast<size_t>(ct) * 2 * plane + (static_cast<size_t>(kg) * 128 + c) * 4;
    *reinterpret_cast<float4*>(base) = h;
    *reinterpret_cast<float4*>(base + plane) = l;
  }
}

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kMwThreads, 1) k_mcw_gemm(const McwArgs a, int gate) {
  if (gate && !*a.need_init) return;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMwStages * kMwStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kMwStages * kMwStageBytes + 96);
  double* lpc = reinterpret_cast<double*>(smem + kMwStages * kMwStageBytes + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + static_cast<uint32_t>(s * 8); };
  auto EMPTY = [&](int s) { return bar0 + static_cast<uint32_t>((4 + s) * 8); };
  auto ACCF = [&](int b) { return bar0 + static_cast<uint32_t>((8 + b) * 8); };
  auto ACCE = [&](int b) { return bar0 + static_cast<uint32_t>((10 + b) * 8); };
  if (tid == 0) {
    for (int s = 0; s < kMwStages; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[4 + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[8 + b], 1);
      mbar_init(&bars[10 + b], kMwEpilogue);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  // ---- work of this CTA ----
  const int ct = blockIdx.x;
  const int NB = MODE == 1 ? kMwRowTile : a.NB2;
  const uint32_t b_plane = static_cast<uint32_t>((kMwKC / 4) * NB * 16);
  int n_items;
  long long c0 = 0, c1 = 0;  // MODE 2: stage range of this split
  const int nch1 = a.Kp1 / kMwKC;
  if (MODE == 1) {
    n_items = static_cast<int>(blockIdx.y) < a.nrt ? (a.nrt - static_cast<int>(blockIdx.y) + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y) : 0;
  } else {
    const long long nck = a.rowsP / kMwKC;
    c0 = nck * blockIdx.z / gridDim.z;
    c1 = nck * (blockIdx.z + 1) / gridDim.z;
    n_items = static_cast<int>((c1 - c0 + kMwSegChunks - 1) / kMwSegChunks);
  }
  auto item_chunks = [&](int i) -> int {
    if (MODE == 1) return nch1;
    const long long b = c0 + static_cast<long long>(i) * kMwSegChunks;
    return static_cast<int>((c1 - b) < kMwSegChunks ? (c1 - b) : kMwSegChunks);
  };
  // operand sources, in floats
  const size_t a_plane_f = MODE == 1 ? static_cast<size_t>(a.Kp1) * 128 : static_cast<size_t>(a.rowsP) * 128;
  const float* a_base = (MODE == 1 ? a.wt : a.rp) + static_cast<size_t>(ct) * 2 * a_plane_f;
  const size_t b_plane_f = MODE == 1 ? static_cast<size_t>(a.Kp1) * kMwRowTile : static_cast<size_t>(a.rowsP) * a.NB2;
  const size_t b_chunk_f = static_cast<size_t>(kMwKC) * NB;

  if (warp == 0) {
    // ================= TMA producer =================
    long long q = 0;
    for (int i = 0; i < n_items; ++i) {
      const int nch = item_chunks(i);
      const float* asrc;
      const float* bsrc;
      if (MODE == 1) {
        const long long t = blockIdx.y + static_cast<long long>(i) * gridDim.y;
        asrc = a_base;
        bsrc = a.xk + static_cast<size_t>(t) * 2 * b_plane_f;
      } else {
        const long long cb = c0 + static_cast<long long>(i) * kMwSegChunks;
        asrc = a_base + static_cast<size_t>(cb) * (kMwKC * 128);
        bsrc = a.xt + static_cast<size_t>(blockIdx.y) * 2 * b_plane_f + static_cast<size_t>(cb) * b_chunk_f;
      }
      for (int c = 0; c < nch; ++c, ++q) {
        const int s = static_cast<int>(q % kMwStages);
        if (q >= kMwStages) mbar_wait_s(EMPTY(s), static_cast<uint32_t>(((q / kMwStages) - 1) & 1));
        if (elect_one()) {
          const uint32_t dst = smem_u32(ring + s * kMwStageBytes);
          mbar_arrive_expect_tx_s(FULL(s), 2 * kMwAPlane + 2 * b_plane);
          const float* ap = asrc + static_cast<size_t>(c) * (kMwKC * 128);
          const float* bp = bsrc + static_cast<size_t>(c) * b_chunk_f;
          bulk_g2s_s(dst, ap, kMwAPlane, FULL(s));
          bulk_g2s_s(dst + kMwAPlane, ap + a_plane_f, kMwAPlane, FULL(s));
          bulk_g2s_s(dst + 2 * kMwAPlane, bp, b_plane, FULL(s));
          bulk_g2s_s(dst + 2 * kMwAPlane + kMwBPlaneMax, bp + b_plane_f, b_plane, FULL(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(128, NB);
    const uint32_t b_lbo = static_cast<uint32_t>(NB * 16);
    long long q = 0;
    for (int i = 0; i < n_items; ++i) {
      const int b = i & 1;
      if (i >= 2) mbar_wait_s(ACCE(b), static_cast<uint32_t>(((i >> 1) - 1) & 1));  // epilogue of item i-2 has drained buffer b
      tc_fence_after();
      const uint32_t d = tm + static_cast<uint32_t>(b * 256);
      const int nch = item_chunks(i);
      for (int c = 0; c < nch; ++c, ++q) {
        const int s = static_cast<int>(q % kMwStages);
        mbar_wait_s(FULL(s), static_cast<uint32_t>((q / kMwStages) & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(ring + s * kMwStageBytes);
        const uint64_t a_h = smem_desc(base, 2048, 128), a_l = smem_desc(base + kMwAPlane, 2048, 128);
        const uint64_t b_h = smem_desc(base + 2 * kMwAPlane, b_lbo, 128);
        const uint64_t b_l = smem_desc(base + 2 * kMwAPlane + kMwBPlaneMax, b_lbo, 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kMwKC / 8; ++ks) {
            const uint64_t oa = static_cast<uint64_t>(ks * (4096 >> 4)), ob = static_cast<uint64_t>(ks * ((2 * b_lbo) >> 4));
            const uint32_t accf = (c > 0 || ks > 0) ? 1u : 0u;
            tc_mma_ss(d, a_h + oa, b_h + ob, idesc, accf);
            if (MODE == 1) {  // correction terms into their own accumulator (columns 128..255 of the buffer)
              tc_mma_ss(d + 128, a_h + oa, b_l + ob, idesc, accf);
              tc_mma_ss(d + 128, a_l + oa, b_h + ob, idesc, 1);
            } else {
              tc_mma_ss(d, a_h + oa, b_l + ob, idesc, 1);
              tc_mma_ss(d, a_l + oa, b_h + ob, idesc, 1);
            }
          }
          tc_commit(EMPTY(s));
          if (c == nch - 1) tc_commit(ACCF(b));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (8 warps: TMEM lane quadrant q, column half hf) =================
    const int q = warp & 3, hf = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int cl = 32 * q + lane;  // chain within the tile
    double lp = 0.0;
    for (int i = 0; i < n_items; ++i) {
      const int b = i & 1;
      mbar_wait_s(ACCF(b), static_cast<uint32_t>((i >> 1) & 1));
      tc_fence_after();
      const uint32_t acc = tm + static_cast<uint32_t>(b * 256) + lane_base;
      if (MODE == 1) {
        const long long t = blockIdx.y + static_cast<long long>(i) * gridDim.y;
        const size_t rplane = static_cast<size_t>(a.rowsP) * 128;
        float* rbase = a.rp + static_cast<size_t>(ct) * 2 * rplane + static_cast<size_t>(cl) * 4;
        for (int gq = 0; gq < kMwRowTile / 32; ++gq) {
          const int col0 = hf * (kMwRowTile / 2) + gq * 16;
          uint32_t v[16], vc[16];
          tmem_ld16(acc + col0, v);
          tmem_ld16(acc + 128 + col0, vc);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(vc[j]));
          const long long row0 = t * kMwRowTile + col0;
          float rv[16];
          float lps = 0.0f;  // 16 terms in fp32, then one float64 add: the FP64 pipe is narrow
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float eta = __uint_as_float(v[j]);
            const float yv = __ldg(a.yt + row0 + j);
            float lpv = 0.0f, r;
            // Bernoulli: the special-function forms of common.cuh (three / two MUFU ops per element) — with the general
            // row_terms (expf, logf, a division) the epilogue of a log-likelihood step outlasted the MMAs of the next tile
            // (tensor pipe 56 % active on those steps against 91 % on gradient-only ones, profiles r01f)
            if (a.family == 0) {
              if (a.want_logp)
                bernoulli_terms_fast(eta, yv, lpv, r);
              else
                r = bernoulli_resid_direct(eta, yv);
            } else if (a.want_logp) {
              row_terms(a.family, eta, yv, a.lik_scale, lpv, r);
            } else {
              r = row_resid(a.family, eta, yv, a.lik_scale);
            }
            if (row0 + j >= a.n_rows) {
              lpv = 0.0f;
              r = 0.0f;
            }
            lps += lpv;
            rv[j] = r;
          }
          lp += static_cast<double>(lps);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 h, l;
            split_tf32(rv[j4 * 4 + 0], h.x, l.x);
            split_tf32(rv[j4 * 4 + 1], h.y, l.y);
            split_tf32(rv[j4 * 4 + 2], h.z, l.z);
            split_tf32(rv[j4 * 4 + 3], h.w, l.w);
            float* dst = rbase + (static_cast<size_t>(row0 >> 2) + j4) * (128 * 4);
            *reinterpret_cast<float4*>(dst) = h;
            *reinterpret_cast<float4*>(dst + rplane) = l;
          }
        }
      } else {
        const int ncg = NB / 16;
        // float64 partial of (split, feature, chain), chain fastest: the 32 chains of a warp are 256 contiguous bytes per
        // feature. The first segment stores, later ones add with fire-and-forget reductions (no load round trip).
        const size_t cs = static_cast<size_t>(a.C);
        double* out = a.part_g64 + (static_cast<size_t>(blockIdx.z) * a.Dp2 + static_cast<size_t>(blockIdx.y) * NB) * cs + ct * 128 + cl;
        for (int gq = hf; gq < ncg; gq += 2) {
          uint32_t v[16];
          tmem_ld16(acc + gq * 16, v);
          tmem_wait_ld();
          double* o = out + static_cast<size_t>(gq * 16) * cs;
          if (i == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j * cs] = static_cast<double>(__uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(o + j * cs, static_cast<double>(__uint_as_float(v[j])));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[10 + b]);
    }
    if (MODE == 1) {
      lpc[hf * 128 + cl] = lp;
    } else if (n_items == 0) {
      const int ncg = NB / 16;
      const size_t cs = static_cast<size_t>(a.C);
      double* out = a.part_g64 + (static_cast<size_t>(blockIdx.z) * a.Dp2 + static_cast<size_t>(blockIdx.y) * NB) * cs + ct * 128 + cl;
      for (int gq = hf; gq < ncg; gq += 2)
        for (int j = 0; j < 16; ++j) out[static_cast<size_t>(gq * 16 + j) * cs] = 0.0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (MODE == 1 && tid < 128) a.part_lp[static_cast<size_t>(blockIdx.y) * a.C + ct * 128 + tid] = lpc[tid] + lpc[128 + tid];
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------------
// float64 fold of the per-split / per-CTA partials: gsum[c] = {Σ G'[c, 0..D), Σ logp[c]}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMwChainThreads) k_mcw_fold(const McwArgs a, int gate) {
  if (gate && !*a.need_init) return;
  const int c = blockIdx.x;
  double* out = a.gsum + static_cast<size_t>(c) * (a.D + 1);
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
    double s = 0.0;
    for (int sp = 0; sp < a.splits; ++sp) s += a.part_g64[(static_cast<size_t>(sp) * a.Dp2 + d) * a.C + c];
    out[d] = s;
  }
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < a.g1; ++j) s += a.part_lp[static_cast<size_t>(j) * a.C + c];
    out[a.D] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// per-chain kernels: grid = C blocks of 256 threads, features strided over the threads. They consume a.gsum
// (already summed over row shards), add the Normal prior and run leapfrog / kinetic energy / MH accept /
// Empirical write exactly like the narrow kernels of chains.cu (hmc.py:81-130,195-210).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mcw_chain_sum(double v, double* sh) {  // fixed order
  v = warp_sum_f64(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kMwChainThreads / 32; ++w) t += sh[w];
  return t;
}

__device__ __forceinline__ double mcw_finish_gradient(const McwArgs& a, int c, const float* pos, float* gout, double* sh) {
  const double* gs = a.gsum + static_cast<size_t>(c) * (a.D + 1);
  double pl = 0.0;
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
    const float loc = a.prior_loc[d], sc = a.prior_scale[d];
    const float zc = pos[static_cast<size_t>(c) * a.D + d];
    gout[static_cast<size_t>(c) * a.D + d] = static_cast<float>(gs[d] + prior_grad(zc, loc, sc));
    pl += prior_quad(zc, loc, sc);
  }
  return (mcw_chain_sum(pl, sh) - a.prior_const) + gs[a.D];
}

// Does the cached state (zcur, gcur, logp_cur) describe row max(t0-1,0) of the Empirical store? Two grid-wide
// steps over the C*D elements: compare (any mismatch raises a.valid[2]), then re-seed zcur and publish need_init.
__global__ void k_mcw_check_compare(const McwArgs a) {
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  const size_t n = static_cast<size_t>(a.Cu) * a.D;  // the caller's chains are the first Cu rows of zcur
  bool mismatch = false;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    if (__float_as_uint(a.params[t_prev * n + i]) != __float_as_uint(a.zcur[i])) mismatch = true;
  if (__syncthreads_or(mismatch ? 1 : 0) && threadIdx.x == 0) atomicExch(a.valid + 2, 1);
}
__global__ void k_mcw_check_apply(const McwArgs a) {
  const int need = (!a.valid[0] || a.valid[2]) ? 1 : 0;
  const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
  const size_t n = static_cast<size_t>(a.Cu) * a.D;  // the caller's chains are the first Cu rows of zcur
  if (need)
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
      a.zcur[i] = a.params[t_prev * n + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.need_init = need;  // nobody in this grid reads need_init
}

__global__ void __launch_bounds__(kMwChainThreads) k_mcw_init_finish(const McwArgs a) {
  __shared__ double sh[8];
  if (!*a.need_init) return;
  const int c = blockIdx.x;
  const double lp = mcw_finish_gradient(a, c, a.zcur, a.gcur, sh);
  if (threadIdx.x == 0) a.logp_cur[c] = lp;
  if (c == 0 && threadIdx.x == 0) *a.valid = 1;
}

__device__ __forceinline__ float mcw_kick(float r, float h, float g) { return __fadd_rn(r, __fmul_rn(h, g)); }
__device__ __forceinline__ float mcw_drift(float z, float e, float r) { return __fadd_rn(z, __fmul_rn(e, r)); }

__device__ __forceinline__ void mcw_finish_transition(const McwArgs& a, int c, long long it, double logp_new, double* sh) {
  double ks = 0.0;
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
    const float rr = a.r[static_cast<size_t>(c) * a.D + d];
    ks += static_cast<double>(__fmul_rn(rr, rr));
  }
  const double k_new = 0.5 * mcw_chain_sum(ks, sh);
  const double logp_cur = a.logp_cur[c], k_old = a.k_old[c], log_u = a.log_u[c];
  const double ratio = ((k_old - k_new) + logp_new) - logp_cur;  // hmc.py:100-105
  const bool accept = log_u < ratio;                              // hmc.py:108-109
  __syncthreads();
  const long long t = a.t0 + it;
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
    const size_t cd = static_cast<size_t>(c) * a.D + d;
    if (accept) {
      a.zcur[cd] = a.z[cd];
      a.gcur[cd] = a.g[cd];
    }
    a.params[(static_cast<size_t>(t) * a.Cu + c) * a.D + d] = accept ? a.z[cd] : a.zcur[cd];
  }
  if (threadIdx.x == 0) {
    if (accept) {
      a.logp_cur[c] = logp_new;
      a.n_accept[c] += 1;
    }
    if (a.trace) {
      double* tr = a.trace + (static_cast<size_t>(it) * a.Cu + c) * 8;
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

__global__ void __launch_bounds__(kMwChainThreads) k_mcw_begin(const McwArgs a, long long it) {
  __shared__ double sh[8];
  const int c = blockIdx.x;
  const long long t = a.t0 + it;
  const unsigned long long cseed = a.seed + 0x9E3779B97F4A7C15ull * (c + 1);
  double ks = 0.0;
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
    const size_t cd = static_cast<size_t>(c) * a.D + d;
    const float rv = a.r0 ? a.r0[(static_cast<size_t>(it) * a.Cu + c) * a.D + d] : philox_normal(cseed, t, d);
    float zz = a.zcur[cd];
    float rr = rv;
    const float gg = a.gcur[cd];
    if (a.L > 0) {
      rr = mcw_kick(rv, a.half_eps, gg);
      zz = mcw_drift(zz, a.eps, rr);
    }
    a.r[cd] = rr;
    a.z[cd] = zz;
    a.g[cd] = gg;
    ks += static_cast<double>(__fmul_rn(rv, rv));
  }
  const double k_old = 0.5 * mcw_chain_sum(ks, sh);
  if (threadIdx.x == 0) {
    const float u = a.u ? a.u[static_cast<size_t>(it) * a.Cu + c] : philox_uniform(cseed, t);
    a.k_old[c] = k_old;
    a.log_u[c] = static_cast<double>(logf(u));
  }
  if (a.L == 0) {
    __syncthreads();
    mcw_finish_transition(a, c, it, a.logp_cur[c], sh);
  }
}

__global__ void __launch_bounds__(kMwChainThreads) k_mcw_leap(const McwArgs a, long long it, int s) {
  __shared__ double sh[8];
  const int c = blockIdx.x;
  const bool last = (s == a.L - 1);
  double logp_new = 0.0;
  if (last) {
    logp_new = mcw_finish_gradient(a, c, a.z, a.g, sh);
  } else {
    const double* gs = a.gsum + static_cast<size_t>(c) * (a.D + 1);
    for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {
      const size_t cd = static_cast<size_t>(c) * a.D + d;
      a.g[cd] = static_cast<float>(gs[d] + prior_grad(a.z[cd], a.prior_loc[d], a.prior_scale[d]));
    }
  }
  for (int d = threadIdx.x; d < a.D; d += kMwChainThreads) {  // each thread re-reads only what it wrote
    const size_t cd = static_cast<size_t>(c) * a.D + d;
    const float gg = a.g[cd];
    float rr = mcw_kick(a.r[cd], a.half_eps, gg);
    if (!last) {
      rr = mcw_kick(rr, a.half_eps, gg);
      a.z[cd] = mcw_drift(a.z[cd], a.eps, rr);
    }
    a.r[cd] = rr;
  }
  if (last) {
    __syncthreads();
    mcw_finish_transition(a, c, it, logp_new, sh);
  }
}

__global__ void __launch_bounds__(kMwChainThreads) k_mcw_logp_grad_finish(const McwArgs a, const float* theta, double* logp,
                                                                          float* grad) {
  __shared__ double sh[8];
  const int c = blockIdx.x;
  const double lp = mcw_finish_gradient(a, c, theta, grad, sh);
  if (threadIdx.x == 0) logp[c] = lp;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void mcw_plan(McwArgs& a, int num_sms) {
  a.nct = a.C / 128;
  a.Kp1 = (a.D + kMwKC - 1) / kMwKC * kMwKC;
  a.nrt = static_cast<int>((a.n_rows + kMwRowTile - 1) / kMwRowTile);
  if (a.nrt < 1) a.nrt = 1;
  a.rowsP = static_cast<long long>(a.nrt) * kMwRowTile;
  // GEMM 2 tiling: feature tiles of NB2 <= 256 columns (multiple of 16) x K splits; pick the tile count that keeps
  // most SMs busy after padding (e.g. C=1024, D=1000: 6 tiles of 176 x 3 splits = 144 CTAs instead of 4 x 256 x 4 = 128)
  {
    const int nft0 = (a.D + 255) / 256;
    double best = -1.0;
    for (int nft = nft0; nft <= nft0 + 6; ++nft) {
      const int nb = ((a.D + nft - 1) / nft + 15) / 16 * 16;
      if (nb < 128 && nft > nft0) break;
      int sp = num_sms / (a.nct * nft);
      if (sp < 1) sp = 1;
      const double ctas = static_cast<double>(a.nct) * nft * sp;
      const double waves = ctas <= num_sms ? 1.0 : static_cast<double>((static_cast<int>(ctas) + num_sms - 1) / num_sms);
      const double eff = ctas / (waves * num_sms) * a.D / (static_cast<double>(nft) * nb);
      if (eff > best + 1e-9) {
        best = eff;
        a.nft = nft;
        a.NB2 = nb;
      }
    }
  }
  a.Dp2 = a.nft * a.NB2;
  int g1 = num_sms / a.nct;
  if (g1 < 1) g1 = 1;
  if (g1 > a.nrt) g1 = a.nrt;
  a.g1 = g1;
  int sp = num_sms / (a.nct * a.nft);
  if (sp < 1) sp = 1;
  const long long nck = a.rowsP / kMwKC;
  if (sp > nck) sp = static_cast<int>(nck);
  a.splits = sp;
}

McwSizes mcw_sizes(const McwArgs& a) {
  McwSizes z;
  z.xk = static_cast<size_t>(a.nrt) * 2 * a.Kp1 * kMwRowTile * sizeof(float);
  z.xt = static_cast<size_t>(a.nft) * 2 * a.rowsP * a.NB2 * sizeof(float);
  z.yt = static_cast<size_t>(a.rowsP) * sizeof(float);
  z.wt = static_cast<size_t>(a.nct) * 2 * a.Kp1 * 128 * sizeof(float);
  z.rp = static_cast<size_t>(a.nct) * 2 * a.rowsP * 128 * sizeof(float);
  z.part_g64 = static_cast<size_t>(a.splits) * a.C * a.Dp2 * sizeof(double);
  z.part_lp = static_cast<size_t>(a.g1) * a.C * sizeof(double);
  z.gsum = static_cast<size_t>(a.C) * (a.D + 1) * sizeof(double);
  return z;
}

cudaError_t mcw_prepare() {
  cudaError_t e = cudaFuncSetAttribute(k_mcw_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMwSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_mcw_gemm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMwSmemBytes);
}

cudaError_t mcw_launch_pretile(const McwArgs& a, cudaStream_t s) {
  k_mcw_pretile_xk<<<148 * 8, 256, 0, s>>>(a);
  k_mcw_pretile_xt<<<148 * 8, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t mcw_launch_pass(const McwArgs& a, const float* theta, int gate, cudaStream_t s) {
  k_mcw_wtile<<<148, 256, 0, s>>>(a, theta, gate);
  k_mcw_gemm<1><<<dim3(a.nct, a.g1, 1), kMwThreads, kMwSmemBytes, s>>>(a, gate);
  k_mcw_gemm<2><<<dim3(a.nct, a.nft, a.splits), kMwThreads, kMwSmemBytes, s>>>(a, gate);
  k_mcw_fold<<<a.C, kMwChainThreads, 0, s>>>(a, gate);
  return cudaGetLastError();
}
cudaError_t mcw_launch_check(const McwArgs& a, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(a.valid + 2, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_mcw_check_compare<<<148, 256, 0, s>>>(a);
  k_mcw_check_apply<<<148, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t mcw_launch_init_finish(const McwArgs& a, cudaStream_t s) {
  k_mcw_init_finish<<<a.Cu, kMwChainThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t mcw_launch_begin(const McwArgs& a, long long it, cudaStream_t s) {
  k_mcw_begin<<<a.Cu, kMwChainThreads, 0, s>>>(a, it);
  return cudaGetLastError();
}
cudaError_t mcw_launch_leap(const McwArgs& a, long long it, int step, cudaStream_t s) {
  k_mcw_leap<<<a.Cu, kMwChainThreads, 0, s>>>(a, it, step);
  return cudaGetLastError();
}
cudaError_t mcw_launch_logp_grad_finish(const McwArgs& a, const float* theta, double* logp, float* grad, cudaStream_t s) {
  k_mcw_logp_grad_finish<<<a.Cu, kMwChainThreads, 0, s>>>(a, theta, logp, grad);
  return cudaGetLastError();
}

}  // namespace edhmc
