// chains.cu — many vectorised HMC chains over one design matrix (see chains.cuh).
//
//   k_mc_pass_tc3     the dense contraction on the 5th-generation tensor cores, hand-written tcgen05:
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
      xs[m][d] = m < nr ? (d < a.Dx ? a.X[(base + m) * a.ldx + d] : 1.0f) : 0.0f;  // d == Dx: the bias latent's column of ones
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
  double* pg = a.part_g + static_cast<size_t>(blockIdx.x) * a.Dp * a.C + c;  // [row group][feature][chain]
#pragma unroll
  for (int d = 0; d < kMcMaxD; ++d)
    if (d < a.Dp) pg[static_cast<size_t>(d) * a.C] = d < D ? static_cast<double>(g[d]) : 0.0;
  a.part_lp[static_cast<size_t>(blockIdx.x) * a.C + c] = lp;
}

constexpr int kTcN2 = 64;     // N of the second contraction: features, padded to 64
constexpr int kT2Rows = 64;    // rows of X per tile of the pipelined pass

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
// The tensor core accumulates in fp32 with truncation: every tcgen05.mma into the same TMEM accumulator loses up to an
// ulp of the running sum, always toward zero. Over a CTA's whole row range (cfg 3: 7,851 rows = 2,900 accumulations) that
// is a systematic 2.4e-5 shrink of the gradient (found by tests/test_gpu_fullsize.py); G' is therefore flushed into a
// float64 partial every kT3Seg tiles (512 rows, 192 accumulations; measured gradient error 4.2e-6 at cfg 3 against 5.9e-5
// without the flush, 4 % of the pass time) from one of two TMEM buffers while
// the MMAs of the next segment run into the other.
constexpr int kT3Seg = 8;  // default of McArgs::seg_tiles

__host__ __device__ inline int tc3_tile_bytes(int Dp) { return 2 * (Dp / 4) * kT3Pitch; }
__host__ __device__ inline int tc3_stages(int Dp) { return Dp > 56 ? 2 : 3; }

__host__ __device__ inline int tc3_smem_layout(int Dp, int* off /*8*/) {
  int o = 0;
  off[0] = o;  // operand ring
  const int stage_bytes = (tc3_tile_bytes(Dp) + 127) / 128 * 128;
  o += tc3_stages(Dp) * stage_bytes;
  off[1] = o;  // ys[stages][64]
  o += kT3MaxStages * kT2Rows * 4;
  off[2] = o;  // (formerly Wᵀ hi / lo as a shared-memory operand; Wᵀ now lives in TMEM)
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
      const float v = (row < a.n_rows && d < a.D) ? (d < a.Dx ? a.X[row * a.ldx + d] : 1.0f) : 0.0f;  // d == Dx: bias column
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

// development timeline (edhmc_set_chain_debug): clock64 of CTA (0,0) per tile and role, [64 tiles][16 slots]:
// 0 TMA issued | 1 MMA1 inputs ready, 2 MMA1 issued | 3 MMA2 inputs ready, 4 MMA2 issued |
// 5 epilogue: tile landed, 6 transposed, 7 S ready, 8 S loaded, 9 R stored and signalled
#define T3_DBG(slot)                                                                       \
  do {                                                                                     \
    if (dbg_on && i < 64) a.dbg[i * 16 + (slot)] = clock64();                              \
  } while (0)
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
  unsigned char* B2T = smem + off[3];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off[4]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off[4] + 112);
  double* lpc = reinterpret_cast<double*>(smem + off[5]);
  const int kc1 = Dp / 4;
  const int plane_bytes = kc1 * kT3Pitch;
  const int tile_bytes = 2 * plane_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.y * kMcChainsPerCta;
  const bool dbg_on = a.dbg != nullptr && (a.dbg_lp != 0) == (a.want_logp != 0) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp < 3 || warp == 4);
  // barrier indices: 0-2 x_full, 3-5 x_empty, 6-7 s_ready, 8-9 r_ready
  const uint32_t bar0 = smem_u32(bars);
  auto XFULL = [&](int s) { return bar0 + static_cast<uint32_t>(s * 8); };
  auto XEMPTY = [&](int s) { return bar0 + static_cast<uint32_t>((3 + s) * 8); };
  auto SREADY = [&](int b) { return bar0 + static_cast<uint32_t>((6 + b) * 8); };
  auto RREADY = [&](int b) { return bar0 + static_cast<uint32_t>((8 + b) * 8); };
  auto GFULL = [&](int g) { return bar0 + static_cast<uint32_t>((10 + g) * 8); };   // segment's MMAs into G'[g] are done
  auto GEMPTY = [&](int g) { return bar0 + static_cast<uint32_t>((12 + g) * 8); };  // G'[g] has been flushed

  if (tid == 0) {
    for (int s = 0; s < kT3MaxStages; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&bars[3 + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[6 + b], 1);
      mbar_init(&bars[8 + b], kT3Epilogue);
      mbar_init(&bars[10 + b], 1);
      mbar_init(&bars[12 + b], kT3Epilogue);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  // B2T: features Dp..63 and the padding are never written by the transpose: zero once
  for (int i = tid; i < 4 * kT3B2Plane / 16; i += kT3Threads) reinterpret_cast<float4*>(B2T)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_g = tm + 256;  // G'[0] at +256, G'[1] at +320
  // Wᵀ (this CTA's 128 chains x Dp features, split hi / lo) goes to TMEM columns 384..447 / 448..511 and is the A operand
  // of MMA1 from there: read from shared memory instead, every MMA1 instruction fetched 4 KB of Wᵀ next to 2 KB of X and
  // ran at 53 cycles (the per-role timeline, profiles/r02_cfg3_*) against 36 for MMA2, whose A operand is in TMEM.
  const uint32_t tm_w = tm + 384;
  if (warp >= 4) {
    const int q = warp & 3, cg = (warp - 4) >> 2;
    const int c = cb + 32 * q + lane;
    uint32_t wh[16], wl[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int d = cg * 16 + j;
      const float v = d < D ? theta[static_cast<size_t>(c) * D + d] : 0.0f;
      float h, l;
      split_tf32(v, h, l);
      wh[j] = __float_as_uint(h);
      wl[j] = __float_as_uint(l);
    }
    const uint32_t lb = static_cast<uint32_t>(32 * q) << 16;
    tmem_st16(tm_w + lb + cg * 16, wh);
    tmem_st16(tm_w + 64 + lb + cg * 16, wl);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  long long u0, u1;
  tile_range(a.n_rows, a.n_rowgroups, blockIdx.x, u0, u1);
  const long long row_begin = u0 * kMcTileRows;
  long long row_end = u1 * kMcTileRows;
  if (row_end > a.n_rows) row_end = a.n_rows;
  const long long tile0 = row_begin / kT2Rows;  // row_begin is a multiple of 128
  const int nt = row_end > row_begin ? static_cast<int>((row_end - row_begin + kT2Rows - 1) / kT2Rows) : 0;
  double lp = 0.0;
  const int seg_tiles = a.seg_tiles > 0 ? a.seg_tiles : kT3Seg;

  if (warp == 0) {
    // ================= TMA producer =================
    int s = 0, ph = 0;  // stage of tile i, parity of the phase its x_empty wait refers to ((i / NS) - 1) & 1
    for (int i = 0; i < nt; ++i) {
      if (i >= NS) mbar_wait_s(XEMPTY(s), ph ^ 1);
      if (elect_one()) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx_s(XFULL(s), tile_bytes + kT2Rows * 4);
        bulk_g2s_s(smem_u32(ring + s * stage_bytes), a.xt + static_cast<size_t>(tile0 + i) * (tile_bytes / 4), tile_bytes, XFULL(s));
        bulk_g2s_s(smem_u32(ysm + s * kT2Rows), a.yt + (tile0 + i) * kT2Rows, kT2Rows * 4, XFULL(s));
      }
      __syncwarp();
      T3_DBG(0);
      if (++s == NS) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (both contractions) =================
    // One thread issues MMA1 and MMA2: tcgen05.mma instructions of one thread execute in issue order, so MMA1(i+2), which
    // overwrites S/R[b], can be issued right behind MMA2(i), which reads R[b], without waiting for it to COMPLETE. With
    // two issuer warps that hazard needed a completion wait (x_empty), and R(i) → MMA2(i) done → MMA1(i+2) issued → done
    // → epilogue(i+2) was the critical path of the pass: 4,400 cycles per two tiles (per-role timeline, profiles/r02_cfg3_*).
    //   MMA1  Sᵀ[b] = Wᵀ·X(j)ᵀ   A = Wᵀ from TMEM, B = the landed tile
    //   MMA2  G'   += R(i)·X(i)  A = R from TMEM,  B = B2T[b]
    const uint32_t id1 = idesc_tf32(128, kT2Rows);
    const uint32_t id2 = idesc_tf32(128, kTcN2);
    const int nks = Dp / 8;
    int s1 = 0, ph1 = 0;  // (stage, x_full parity) of the next tile MMA1 is issued for
    auto issue_mma1 = [&](int i) {
      const int b = i & 1;
      mbar_wait_s(XFULL(s1), ph1);
      tc_fence_after();
      T3_DBG(1);
      const uint32_t xs = smem_u32(ring + s1 * stage_bytes);
      const uint64_t db_h = smem_desc(xs, kT3Pitch, 128), db_l = smem_desc(xs + plane_bytes, kT3Pitch, 128);
      const uint32_t ts = tm + 64 * b;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (ks < nks) {
            const uint64_t ob = static_cast<uint64_t>(ks * ((2 * kT3Pitch) >> 4));
            tc_mma_ts(ts, tm_w + ks * 8, db_h + ob, id1, ks > 0);
            tc_mma_ts(ts, tm_w + ks * 8, db_l + ob, id1, 1);
            tc_mma_ts(ts, tm_w + 64 + ks * 8, db_h + ob, id1, 1);
          }
        }
        tc_commit(SREADY(b));
      }
      __syncwarp();
      T3_DBG(2);
      if (++s1 == NS) {
        s1 = 0;
        ph1 ^= 1;
      }
    };
    for (int i = 0; i < 2 && i < nt; ++i) issue_mma1(i);
    int s = 0, ph = 0, sg = 0, in_seg = 0;  // stage / phase of tile i; its segment and position inside it
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int gb = sg & 1;
      const bool seg_first = in_seg == 0, seg_last = in_seg == seg_tiles - 1 || i == nt - 1;
      mbar_wait_s(RREADY(b), (i >> 1) & 1);
      if (seg_first && sg >= 2) mbar_wait_s(GEMPTY(gb), ((sg >> 1) - 1) & 1);  // the epilogue has flushed segment sg-2
      tc_fence_after();
      T3_DBG(3);
      const uint32_t bs = smem_u32(B2T + b * 2 * kT3B2Plane);
      const uint64_t d2_h = smem_desc(bs, kT3B2Lbo, kT3B2Sbo), d2_l = smem_desc(bs + kT3B2Plane, kT3B2Lbo, kT3B2Sbo);
      const uint32_t r_h = tm + 64 * b, r_l = tm + 128 + 64 * b;
      const uint32_t acc = tm_g + 64 * gb;
      const uint32_t first = seg_first ? 0u : 1u;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kT2Rows / 8; ++ks) {
          const uint64_t o = static_cast<uint64_t>(ks * ((2 * kT3B2Lbo) >> 4));  // 8 rows = 2 row groups
          tc_mma_ts(acc, r_h + ks * 8, d2_h + o, id2, ks > 0 ? 1u : first);
          tc_mma_ts(acc, r_h + ks * 8, d2_l + o, id2, 1);
          tc_mma_ts(acc, r_l + ks * 8, d2_h + o, id2, 1);
        }
        tc_commit(XEMPTY(s));  // stage s, ys[s], S[b]/R[b] and B2T[b] are free again
        if (seg_last) tc_commit(GFULL(gb));
      }
      __syncwarp();
      T3_DBG(4);
      if (i + 2 < nt) issue_mma1(i + 2);  // into S/R[b], behind MMA2(i) in the same in-order pipe
      if (++in_seg == seg_tiles) {
        in_seg = 0;
        ++sg;
      }
      if (i + 1 < nt && ++s == NS) {
        s = 0;
        ph ^= 1;
      }
    }
    if (nt > 0) mbar_wait_s(XEMPTY(s), ph);  // the last commit covers every MMA
  } else if (warp >= 4) {
    // ================= epilogue (16 warps): transpose the landed tile, then Sᵀ → R =================
    // (Two groups of 8 warps converting alternate tiles concurrently were measured and are slower, 318 vs 288 us per step:
    // R overwrites S in place and there are two S/R buffers, so MMA1(i+2) waits for MMA2(i), which waits for the epilogue
    // of tile i — the epilogue LATENCY of one tile is on the critical path, and 8 warps take twice as long as 16.)
    const int q = warp & 3, cg = (warp - 4) >> 2;  // TMEM lane quadrant, column group (16 columns)
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int col = cg * 16;
    // transpose work item of this thread: plane p (hi/lo), feature chunk kc, row block mb; kc fastest across lanes
    const int et = tid - 128;
    const int npk = 2 * kc1;              // (plane, kc) pairs
    const int pk = et % npk, mbt = et / npk;  // valid when mbt < 16
    const int tp = pk / kc1, tkc = pk - tp * kc1;
    const bool has_item = mbt < 16;
    const int it_src = tp * plane_bytes + tkc * kT3Pitch + mbt * 64;
    const int it_dst = tp * kT3B2Plane + mbt * kT3B2Lbo + ((tkc * 4) >> 3) * kT3B2Sbo + ((tkc * 4) & 7) * 16;
    // float64 running sums of this thread's (chain, 16 features) in global memory (L2-resident), laid out
    // [row group][feature][chain] so that the 32 chains of a warp are 256 contiguous bytes per feature: segment 0
    // stores, later segments add with fire-and-forget reductions (no load round trip in the epilogue warps)
    double* pg = a.part_g + static_cast<size_t>(blockIdx.x) * Dp * a.C + cb + 32 * q + lane;
    const size_t pgs = static_cast<size_t>(a.C);
    auto flush_segment = [&](int sg) {
      const int gb = sg & 1;
      mbar_wait_s(GFULL(gb), (sg >> 1) & 1);
      tc_fence_after();
      uint32_t gv[16];
      tmem_ld16(tm_g + 64 * gb + lane_base + col, gv);
      tmem_wait_ld();
      if (a.seg_mode == 1) {
        // (development) synchronisation only
      } else if (sg == 0 || a.seg_mode == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < Dp) pg[(col + j) * pgs] = static_cast<double>(__uint_as_float(gv[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < Dp) atomicAdd(pg + (col + j) * pgs, static_cast<double>(__uint_as_float(gv[j])));
      }
      tc_fence_before();
      mbar_arrive(&bars[12 + gb]);
    };
    // cursors kept incrementally (integer divisions by run-time values cost ~20 instructions each in every warp)
    int s = 0, xph = 0, sm1 = 0, phm1 = 0, sm2 = 0, phm2 = 0;  // (stage, x_full parity) of tiles i, i-1, i-2
    int sg = 0, in_seg = 0;
    const bool plain = !(a.want_logp || a.family != 0);  // gradient-only Bernoulli pass: the short residual
    const int flush_at = seg_tiles > 1 ? 1 : 0;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      // The float64 flush of segment sg-1 waits for ITS last MMA2, which the tensor pipe may still be executing when
      // the first tile of segment sg arrives here; flushing one tile later (G' is double-buffered, the buffer is needed
      // again only a whole segment later) keeps the epilogue from idling behind the tensor pipe.
      if (sg > 0 && in_seg == flush_at) flush_segment(sg - 1);
      const long long r0 = row_begin + static_cast<long long>(i) * kT2Rows;
      const int rows = row_end - r0 >= kT2Rows ? kT2Rows : static_cast<int>(row_end - r0);
      mbar_wait_s(XFULL(s), xph);
      if (i >= 2) mbar_wait_s(XEMPTY(sm2), phm2);  // MMA2(i-2) has read B2T[b]
      T3_DBG(5);
      if (has_item) {
        const unsigned char* src = ring + s * stage_bytes + it_src;
        float4 v0 = *reinterpret_cast<const float4*>(src);
        float4 v1 = *reinterpret_cast<const float4*>(src + 16);
        float4 v2 = *reinterpret_cast<const float4*>(src + 32);
        float4 v3 = *reinterpret_cast<const float4*>(src + 48);
        unsigned char* dst = B2T + b * 2 * kT3B2Plane + it_dst;
        *reinterpret_cast<float4*>(dst) = make_float4(v0.x, v1.x, v2.x, v3.x);
        *reinterpret_cast<float4*>(dst + 16) = make_float4(v0.y, v1.y, v2.y, v3.y);
        *reinterpret_cast<float4*>(dst + 32) = make_float4(v0.z, v1.z, v2.z, v3.z);
        *reinterpret_cast<float4*>(dst + 48) = make_float4(v0.w, v1.w, v2.w, v3.w);
      }
      fence_proxy_async_smem();
      T3_DBG(6);
      mbar_wait_s(SREADY(b), (i >> 1) & 1);
      tc_fence_after();
      T3_DBG(7);
      const float4* ys4 = reinterpret_cast<const float4*>(ysm + s * kT2Rows + col);
      float yv[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 t = ys4[j4];
        yv[4 * j4] = t.x;
        yv[4 * j4 + 1] = t.y;
        yv[4 * j4 + 2] = t.z;
        yv[4 * j4 + 3] = t.w;
      }
      uint32_t v[16], vh[16], vl[16];
      tmem_ld16(tm + 64 * b + lane_base + col, v);
      tmem_wait_ld();
      T3_DBG(8);
      if (a.family == 0 && !plain) {
        // last leapfrog step of a trajectory, Bernoulli: residual AND log-likelihood from three special-function ops per
        // element (common.cuh bernoulli_terms_fast); the 16 terms of a tile are summed in float32, one float64 add per tile
        // (the general path below paced this pass at 8,300 cycles per tile against 2,150 for a gradient-only one)
        float lpt = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = col + j;
          float lpv, rv;
          bernoulli_terms_fast(__uint_as_float(v[j]), yv[j], lpv, rv);
          if (m >= rows) {
            lpv = 0.0f;
            rv = 0.0f;
          }
          lpt += lpv;
          float h, l;
          split_tf32(rv, h, l);
          vh[j] = __float_as_uint(h);
          vl[j] = __float_as_uint(l);
        }
        lp += static_cast<double>(lpt);
      } else if (!plain) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = col + j;
          float lpv, rv;
          row_terms(a.family, __uint_as_float(v[j]), yv[j], a.lik_scale, lpv, rv);
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
      } else if (rows == kT2Rows) {
        // full tile, gradient only: y - 1 / (1 + 2^(-eta log2 e)), five instructions and two special-function ops per
        // element (absolute error <= 1.2e-7 per residual, random in sign: far inside the 1e-5 of the gradient)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float rv = bernoulli_resid_direct(__uint_as_float(v[j]), yv[j]);
          float h, l;
          split_tf32(rv, h, l);
          vh[j] = __float_as_uint(h);
          vl[j] = __float_as_uint(l);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = col + j;
          float rv = bernoulli_resid_direct(__uint_as_float(v[j]), yv[j]);
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
      T3_DBG(9);
      sm2 = sm1;
      phm2 = phm1;
      sm1 = s;
      phm1 = xph;
      if (++s == NS) {
        s = 0;
        xph ^= 1;
      }
      if (++in_seg == seg_tiles) {
        in_seg = 0;
        ++sg;
      }
    }
    if (nt > 0) {
      // segments whose deferred flush did not come up inside the loop: the one before the last if the last segment has
      // a single tile, then the last one
      const int last_seg = (nt - 1) / seg_tiles;
      if (last_seg > 0 && flush_at == 1 && (nt - 1) - last_seg * seg_tiles < 1) flush_segment(last_seg - 1);
      flush_segment(last_seg);
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
    const int col = cg * 16;
    (void)lane_base;
    if (nt == 0) {  // a row group without rows still publishes its (zero) partial
      double* pgz = a.part_g + static_cast<size_t>(blockIdx.x) * Dp * a.C + chain;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < Dp) pgz[static_cast<size_t>(col + j) * a.C] = 0.0;
    }
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
  for (int rg = 0; rg < a.n_rowgroups; ++rg) s += a.part_g[(static_cast<size_t>(rg) * a.Dp + d) * a.C + c];
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
  const size_t n = static_cast<size_t>(a.Cu) * a.D;  // the caller's chains are the first Cu rows of zcur
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
    a.params[(static_cast<size_t>(t) * a.Cu + c) * a.D + d] = accept ? a.z[cd] : a.zcur[cd];
  }
  if (d == 0) {
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

// start of transition `it` for every chain: momentum, kinetic energy, first half kick + drift (hmc.py:88-97,201-204)
__global__ void __launch_bounds__(kMcChainThreads) k_mc_begin(const McArgs a, long long it) {
  __shared__ double sh[2];
  const int c = blockIdx.x, d = threadIdx.x;
  const long long t = a.t0 + it;
  const size_t cd = static_cast<size_t>(c) * a.D + d;
  float rv = 0.0f;
  if (d < a.D) {
    rv = a.r0 ? a.r0[(static_cast<size_t>(it) * a.Cu + c) * a.D + d] : philox_normal(a.seed + 0x9E3779B97F4A7C15ull * (c + 1), t, d);
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
    const float u = a.u ? a.u[static_cast<size_t>(it) * a.Cu + c] : philox_uniform(a.seed + 0x9E3779B97F4A7C15ull * (c + 1), t);
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
  return cudaSuccess;
}

cudaError_t mc_launch_pass(const McArgs& a, const float* theta, int use_tc, int gate, cudaStream_t s) {
  dim3 grid(a.n_rowgroups, a.C / kMcChainsPerCta);
  if (use_tc == 3) {
    int off[8];
    const int smem = tc3_smem_layout(a.Dp, off);
    k_mc_pass_tc3<<<grid, kT3Threads, smem, s>>>(a, theta, gate);
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
  k_mc_init_finish<<<a.Cu, kMcChainThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t mc_launch_begin(const McArgs& a, long long it, cudaStream_t s) {
  k_mc_begin<<<a.Cu, kMcChainThreads, 0, s>>>(a, it);
  return cudaGetLastError();
}
cudaError_t mc_launch_leap(const McArgs& a, long long it, int step, cudaStream_t s) {
  k_mc_leap<<<a.Cu, kMcChainThreads, 0, s>>>(a, it, step);
  return cudaGetLastError();
}
cudaError_t mc_launch_logp_grad_finish(const McArgs& a, const float* theta, double* logp, float* grad, cudaStream_t s) {
  k_mc_logp_grad_finish<<<a.Cu, kMcChainThreads, 0, s>>>(a, theta, logp, grad);
  return cudaGetLastError();
}

}  // namespace edhmc
