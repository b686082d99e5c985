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
// tcgen05 helpers (inline PTX; SASS: UTCHMMA-family / LDTM / STTM)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t mbar_s) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_s) : "memory");
}
// D[tmem] (+)= A[smem] · B[smem]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte
// units): a core matrix is 8 rows x 16 bytes stored contiguously (128 B); SBO = byte step between 8-row
// groups along M/N, LBO = byte step between core matrices along K. Bits: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (SWIZZLE_NONE) [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr_s, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr_s >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor kind::tf32: D=f32 (bits[4,6)=1), A=B=tf32 (bits[7,10)=bits[10,13)=2), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // the 19 bits the tensor core reads
  lo = x - hi;                                               // exact in fp32
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

cudaError_t mc_prepare_tc() {
  return cudaFuncSetAttribute(k_mc_pass_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem_bytes_tc(kMcMaxD));
}

cudaError_t mc_launch_pass(const McArgs& a, const float* theta, int use_tc, int gate, cudaStream_t s) {
  dim3 grid(a.n_rowgroups, a.C / kMcChainsPerCta);
  if (use_tc) {
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
