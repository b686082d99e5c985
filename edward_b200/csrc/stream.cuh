// stream.cuh — the fused streaming pass over X: one read of every row computes the linear predictor
// X·w (+b), the likelihood term, its residual d logp/d eta, and the accumulation Xᵀ·residual.
//
// Replaces, per evaluation, the reference's CheckNumerics(X) + MatMul + ~10 elementwise ops + Sum +
// backward elementwise ops + MatMul(Xᵀ) chain built by ed.dot (util/tensorflow.py:10-45),
// Bernoulli.log_prob and tf.gradients (hmc.py:199,206).
//
// Layout: each warp owns a contiguous range of rows and a private ring of S shared-memory stages.
// A stage holds one tile = RT consecutive rows of X (contiguous in HBM, so the tile is ONE 1-D TMA
// bulk copy, cp.async.bulk → UBLKCP) followed by the tile's y slice (cp.async 4-byte, folded into the
// same mbarrier). The warp that consumes a stage re-arms it itself, so there are no "empty" barriers,
// no CTA-wide synchronisation in the main loop, and the ring keeps prefetching the NEXT pass's tiles
// while the grid synchronises on the current pass's reduction.
//
// Work decomposition inside a warp: a row is handled by G lanes (G = 1..32, power of two); lane `lg` of
// the group owns vector chunks k*G+lg (k < K, K a compile-time tier) of V floats each, i.e. 128-/64-/
// 32-bit shared loads, and the arithmetic is packed FFMA2 (two fp32 FMAs per instruction) when V >= 2.
// theta is zero-padded to G*K*V entries, so padded columns contribute 0 to the dot product and their
// gradient accumulators are simply never written out. The planner picks the smallest G whose per-lane
// slice (K*V floats) keeps the kernel under 128 registers, so that 16 warps per SM hide the latency of the
// dependent FMA / MUFU chains; wide rows fall back to 12 or 8 warps.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace edhmc {

// Per-warp description of its rows, in registers.
struct WarpTiles {
  const float* x0;  // first float of the warp's first row
  const char* y0;   // first y byte of the warp's first row
  int nt;           // tiles per pass
  int rows_last;    // rows in tile nt-1
  int tail;         // 1: tile nt-1 ends at the last row of X (must not read past its D-th float)
};

__device__ __forceinline__ WarpTiles warp_tiles(const KArgs& a, int gw, int total_w) {
  WarpTiles w;
  if (a.interleave) {
    // tile t belongs to warp t % total_w: at any moment the grid reads one compact window of X that moves forward
    // (DRAM page locality of a grid-stride loop) instead of total_w separate streams
    const long long ntiles = (a.n_rows + a.RT - 1) / a.RT;
    w.nt = gw < ntiles ? static_cast<int>((ntiles - gw + total_w - 1) / total_w) : 0;
    w.x0 = a.X + static_cast<long long>(gw) * a.tl;
    w.y0 = reinterpret_cast<const char*>(a.y) + static_cast<long long>(gw) * a.RT * 4;
    const bool owns_last = w.nt > 0 && gw + static_cast<long long>(w.nt - 1) * total_w == ntiles - 1;
    w.rows_last = w.nt ? (owns_last ? static_cast<int>(a.n_rows - (ntiles - 1) * a.RT) : a.RT) : 0;
    w.tail = owns_last ? 1 : 0;
    return w;
  }
  const long long units = (a.n_rows + 3) >> 2;  // 4-row units keep every warp's first row 16-byte aligned
  const long long u0 = units * gw / total_w, u1 = units * (gw + 1) / total_w;
  long long begin = u0 * 4;
  long long end = u1 * 4 < a.n_rows ? u1 * 4 : a.n_rows;
  if (end < begin) end = begin;
  w.x0 = a.X + begin * a.ldx;
  w.y0 = reinterpret_cast<const char*>(a.y) + begin * 4;
  const long long n = end - begin;
  w.nt = static_cast<int>((n + a.RT - 1) / a.RT);
  w.rows_last = w.nt ? static_cast<int>(n - static_cast<long long>(w.nt - 1) * a.RT) : 0;
  w.tail = (n > 0 && end == a.n_rows) ? 1 : 0;
  return w;
}

struct Ring {
  uint32_t base_s;  // shared-window address of this warp's first stage
  uint32_t bars_s;  // shared-window address of this warp's S mbarriers
  // consumer cursor
  int stage;
  uint32_t parity;
  int cpass;  // parity source of the pass being consumed (zig-zag order)
  // producer cursor
  long long qi, q_total;  // next tile to issue / tiles in this launch
  int ipass, ik, istage;
};

__device__ __forceinline__ void ring_init(Ring& ring, float* base, uint64_t* bars) {
  ring.base_s = smem_u32(base);
  ring.bars_s = smem_u32(bars);
  ring.stage = 0;
  ring.parity = 0;
  ring.cpass = 0;
  ring.qi = 0;
  ring.q_total = 0;
  ring.ipass = 0;
  ring.ik = 0;
  ring.istage = 0;
}

// Loop-invariant plan constants, hoisted into registers once per kernel.
struct PlanRegs {
  int stage_bytes, y_off_bytes, RT, tm, tl, S, ldx, D, zigzag, l2_hint;
  long long xstride;  // floats between consecutive tiles of a warp
  long long ystride;  // bytes of y between them
};
__device__ __forceinline__ PlanRegs plan_regs(const KArgs& a) {
  PlanRegs p;
  p.stage_bytes = a.stage_floats * 4;
  p.y_off_bytes = a.y_off * 4;
  p.RT = a.RT;
  p.tm = a.tm;
  p.tl = a.tl;
  p.S = a.S;
  p.ldx = a.ldx_i;
  p.D = a.D;
  p.zigzag = a.zigzag;
  p.l2_hint = a.l2_hint;
  const long long every = a.interleave ? static_cast<long long>(gridDim.x) * (blockDim.x >> 5) : 1;
  p.xstride = every * a.tl;
  p.ystride = every * a.RT * 4;
  return p;
}

// Issues the next tile of this warp's sequence into stage `istage`. Called by ALL lanes of the warp
// (converged): every lane copies its share of the y slice, lane 0 launches the bulk copy of X.
// A tile whose first float is not 16-byte aligned (possible only when ldx % 4 != 0) is copied from the
// aligned address below it; the consumer skips the same `m` leading floats.
__device__ __forceinline__ void ring_issue(const PlanRegs& pr, const WarpTiles& wt, Ring& ring, int lane, uint64_t policy) {
  const int kk = (pr.zigzag && (ring.ipass & 1)) ? (wt.nt - 1 - ring.ik) : ring.ik;
  const bool last = (kk == wt.nt - 1);
  const int rows = last ? wt.rows_last : pr.RT;
  const uint32_t sb = ring.base_s + ring.istage * pr.stage_bytes;
  const uint32_t bar = ring.bars_s + ring.istage * 8;
  const char* ysrc = wt.y0 + kk * pr.ystride + lane * 4;
  const uint32_t ydst = sb + pr.y_off_bytes + lane * 4;
  if (lane < rows) cp_async4_s(ydst, ysrc);
  for (int e = 32; lane + e < rows; e += 32) cp_async4_s(ydst + e * 4, ysrc + e * 4);
  cp_async_mbar_arrive_noinc_s(bar);
  if (lane == 0) {
    const float* tile = wt.x0 + kk * pr.xstride;
    const int m = static_cast<int>(reinterpret_cast<uintptr_t>(tile) >> 2) & 3;  // X itself is 16-byte aligned
    const float* src = tile - m;
    uint32_t bytes;
    if (last && wt.tail) {
      // never read past the last valid float of X: bulk-copy the 16-byte multiple, finish with scalar copies
      const int nfl = m + (rows - 1) * pr.ldx + pr.D;
      bytes = static_cast<uint32_t>(nfl * 4) & ~15u;
      for (int i = bytes >> 2; i < nfl; ++i) sts_f32(sb + i * 4, __ldg(src + i));
    } else {
      bytes = (static_cast<uint32_t>((m + rows * pr.ldx) * 4) + 15u) & ~15u;
    }
    fence_proxy_async_smem();
    mbar_arrive_expect_tx_s(bar, bytes);
    if (bytes) {
      if (pr.l2_hint)
        bulk_g2s_hint_s(sb, src, bytes, bar, policy);
      else
        bulk_g2s_s(sb, src, bytes, bar);
    }
  }
  ++ring.qi;
  if (++ring.ik == wt.nt) {
    ring.ik = 0;
    ++ring.ipass;
  }
  if (++ring.istage == pr.S) ring.istage = 0;
}

// Fills the ring at the start of a launch (all lanes).
__device__ __forceinline__ void ring_prologue(const PlanRegs& pr, const WarpTiles& wt, Ring& ring, long long n_passes,
                                              int lane, uint64_t policy) {
  ring.q_total = n_passes * wt.nt;
  for (int s = 0; s < pr.S && ring.qi < ring.q_total; ++s) ring_issue(pr, wt, ring, lane, policy);
}

// ---- packed helpers -------------------------------------------------------------------------------
template <int V>
struct Acc {
  using type = float2;
};
template <>
struct Acc<1> {
  using type = float;
};

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// Loads K chunks of V floats starting at p with a chunk stride of `cstride` floats into packed registers.
template <int V, int K>
__device__ __forceinline__ void load_chunks(const float* p, int cstride, typename Acc<V>::type* out) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if constexpr (V == 1) {
      out[k] = p[k * cstride];
    } else if constexpr (V == 2) {
      out[k] = *reinterpret_cast<const float2*>(p + k * cstride);
    } else {
      const float4 t = *reinterpret_cast<const float4*>(p + k * cstride);
      out[2 * k] = make_float2(t.x, t.y);
      out[2 * k + 1] = make_float2(t.z, t.w);
    }
  }
}

// Same, from a 32-bit shared-window address with compile-time chunk stride CS (floats): explicit ld.shared
// with immediate offsets, so the hot loop carries one address register per row.
template <int V, int K, int CS>
struct ChunkLoader {
  template <int k>
  static __device__ __forceinline__ void step(uint32_t addr, typename Acc<V>::type* out) {
    if constexpr (k < K) {
      if constexpr (V == 1) {
        out[k] = lds_f32<k * CS * 4>(addr);
      } else if constexpr (V == 2) {
        out[k] = lds_f32x2<k * CS * 4>(addr);
      } else {
        const float4 t = lds_f32x4<k * CS * 4>(addr);
        out[2 * k] = make_float2(t.x, t.y);
        out[2 * k + 1] = make_float2(t.z, t.w);
      }
      step<k + 1>(addr, out);
    }
  }
  static __device__ __forceinline__ void load(uint32_t addr, typename Acc<V>::type* out) { step<0>(addr, out); }
};

template <int V, int NA>
__device__ __forceinline__ float flat(const typename Acc<V>::type* g, int i) {
  if constexpr (V == 1) {
    return g[i];
  } else {
    return (i & 1) ? g[i >> 1].y : g[i >> 1].x;
  }
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// Shared-memory carve-up.
struct SmemLayout {
  float* ring;
  uint64_t* bars;
  double* cta_acc;
  double* red;
  double* comb;
  float* theta_s;
  float* state;  // 5 * ppad floats
  float* xw;     // kXwFloats floats + 2*kMaxWarps doubles: one-shot cross-warp reduction scratch
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t smem_layout_bytes(int nw, int S, int stage_floats, int P, int wpad, size_t* offs /*8*/) {
  size_t off = 0;
  offs[0] = off;
  off += static_cast<size_t>(nw) * S * stage_floats * 4;
  off = align_up(off, 128);
  offs[1] = off;
  off += static_cast<size_t>(kMaxWarps) * kMaxStages * 8;  // mbarriers
  off += static_cast<size_t>(kMaxWarps) * kMaxStages * 4;  // consumer counters of the warp-group rings (ring mode 1)
  offs[2] = off;
  off += align_up(static_cast<size_t>(P + 1) * 8, 16);
  offs[3] = off;
  off += 64 * 8;
  offs[4] = off;
  off += static_cast<size_t>(kMaxWarps) * 32 * 8;
  offs[5] = off;
  off += align_up(static_cast<size_t>(wpad) * 4, 16);
  offs[6] = off;
  off += 5 * align_up(static_cast<size_t>(P) * 4, 16);
  offs[7] = off;
  off += static_cast<size_t>(kXwFloats) * 4 + 2 * kMaxWarps * 8;
  return align_up(off, 128);
}

__device__ __forceinline__ SmemLayout carve_smem(unsigned char* raw, const KArgs& a, int nw) {
  size_t offs[8];
  smem_layout_bytes(nw, a.S, a.stage_floats, a.P, a.wpad, offs);
  SmemLayout L;
  L.ring = reinterpret_cast<float*>(raw + offs[0]);
  L.bars = reinterpret_cast<uint64_t*>(raw + offs[1]);
  L.cta_acc = reinterpret_cast<double*>(raw + offs[2]);
  L.red = reinterpret_cast<double*>(raw + offs[3]);
  L.comb = reinterpret_cast<double*>(raw + offs[4]);
  L.theta_s = reinterpret_cast<float*>(raw + offs[5]);
  L.state = reinterpret_cast<float*>(raw + offs[6]);
  L.xw = reinterpret_cast<float*>(raw + offs[7]);
  return L;
}

// Zero the ring (so padded / stale columns are finite), init the mbarriers, zero theta_s.
__device__ __forceinline__ void smem_setup(const SmemLayout& L, const KArgs& a, int nw) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nring = nw * a.S * a.stage_floats;
  for (int i = tid; i < nring; i += nthr) L.ring[i] = 0.0f;
  for (int i = tid; i < a.wpad; i += nthr) L.theta_s[i] = 0.0f;
  if (tid < kMaxWarps * kMaxStages) mbar_init(L.bars + tid, 33);  // 32 cp.async arrivals + 1 expect_tx arrival
  fence_mbar_init();
  fence_proxy_async_smem();
  __syncthreads();
}

// Lane map of a warp. Interleaved (SPLIT = false): lane = grp*G + lg — the G lanes of a row are adjacent (per-warp
// rings). Split (SPLIT = true): lane = lg*RPS + grp — lanes of one shared-memory wavefront read the same chunk of
// consecutive rows, so the bank pattern is that of one lane per row whatever G is (CTA-wide ring).
template <int G, bool SPLIT>
struct LaneMap {
  static constexpr int RPS = 32 / G;
  static __device__ __forceinline__ int lg(int lane) { return SPLIT ? lane / RPS : (lane & (G - 1)); }
  static __device__ __forceinline__ int grp(int lane) { return SPLIT ? (lane & (RPS - 1)) : lane / G; }
  // xor distance of lane bit `b` of lg
  static __device__ __forceinline__ constexpr int lg_xor(int b) { return SPLIT ? b * RPS : b; }
};

// Balanced binary tree over 16 float64 slots ((v0+v1)+(v2+v3))+..., the fixed summation order of every cross-CTA and
// cross-warp reduction of the narrow-model path: float64 adds have a long latency on this part (a chain of 8 costs ~600
// cycles, measured with the per-pass timeline), a tree of 16 is four levels deep. Absent slots hold 0.0 (x + 0.0 == x).
__device__ __forceinline__ double tree16(double (&v)[16]) {
#pragma unroll
  for (int w = 1; w < 16; w <<= 1) {
#pragma unroll
    for (int i = 0; i < 16; i += 2 * w) v[i] += v[i + w];
  }
  return v[0];
}

// End of a pass: sums the per-lane gradient accumulators g[], the bias gradient gb and the log-likelihood lp over the
// warp (halving butterfly) and then over the warps of the CTA (fixed order, float64) into sm.cta_acc[0..P].
template <int G, int V, int K, int NW, bool SPLIT, typename WT, typename RING>
__device__ __forceinline__ void pass_reduce(const KArgs& a, const PlanRegs& pr, const WT& wt, RING& ring, const SmemLayout& sm,
                                            const typename Acc<V>::type* g, float gb, double lp, bool park, bool deferred,
                                            uint32_t park_s, uint64_t policy, long long* tlx = nullptr) {
  constexpr int RPS = 32 / G;
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = LaneMap<G, SPLIT>::lg(lane), grp = LaneMap<G, SPLIT>::grp(lane);
  double* cta_acc = sm.cta_acc;
  // ---- reduce across the RPS row groups of the warp with a halving butterfly: after stage `st` a lane
  //      keeps the half of the accumulators selected by its own bit, so 32+16+8+4+2 shuffles sum 64
  //      accumulators over 32 lanes (instead of 64*5). ----
  constexpr int NST = ilog2(RPS);
  constexpr int LP = (KV + RPS - 1) / RPS * RPS;  // padded length, divisible by RPS = 2^NST
  constexpr int LPF = LP / RPS;                   // accumulators a lane ends up owning
  float h[LP];
#pragma unroll
  for (int i = 0; i < LP; ++i) h[i] = (i < KV) ? flat<V, NA>(g, i) : 0.0f;
#pragma unroll
  for (int st = 0; st < NST; ++st) {
    const int off = SPLIT ? ((RPS >> 1) >> st) : (16 >> st);
    const int half = LP >> (st + 1);
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? h[i] : h[i + half];
      const float keep = upper ? h[i + half] : h[i];
      h[i] = keep + __shfl_xor_sync(kFull, send, off);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) gb += __shfl_xor_sync(kFull, gb, off);
  lp = warp_sum_f64(lp);

  // ---- reduce across the warps of the CTA: fixed order, float64. Lane (grp, lg) owns flat accumulator
  //      indices grp*LPF + j, i.e. chunk k = idx / V, element v = idx % V, column (k*G+lg)*V+v. ----
  const int D = a.D, P = a.P;
  double* xwd = reinterpret_cast<double*>(sm.xw + kXwFloats);  // [2][kMaxWarps]: bias-gradient and logp per warp
  if (!park) {
    // one-shot: every warp publishes its column sums, then each column is summed over the warps in order
#pragma unroll
    for (int j = 0; j < LPF; ++j) {
      const int idx = grp * LPF + j;
      const int col = ((idx / V) * G + lg) * V + (idx % V);
      if (idx < KV && col < D) sm.xw[warp * D + col] = h[j];
    }
    if (lane == 0) {
      xwd[warp] = static_cast<double>(gb);
      xwd[kMaxWarps + warp] = lp;
    }
    if (tlx != nullptr && threadIdx.x == 0) tlx[10] = clock64();  // slot 18: warp 0 reduced and published its sums
    __syncthreads();
    if (tlx != nullptr && threadIdx.x == 0) tlx[11] = clock64();  // slot 19: all warps did
    for (int c = threadIdx.x; c <= P; c += NW * 32) {
      double v[16];
#pragma unroll
      for (int wq = 0; wq < 16; ++wq) v[wq] = 0.0;
      if (c < D) {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = static_cast<double>(sm.xw[wq * D + c]);
      } else if (c == P) {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = xwd[kMaxWarps + wq];
      } else {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = xwd[wq];
      }
      cta_acc[c] = tree16(v);
    }
    __syncthreads();
  } else {
    // parked one-shot: same arithmetic as above (float64 sum over the warps in ascending order), the scratch of
    // warp w being the D floats at the start of its parked stage
    uint32_t* park_tab = reinterpret_cast<uint32_t*>(sm.xw);
#pragma unroll
    for (int j = 0; j < LPF; ++j) {
      const int idx = grp * LPF + j;
      const int col = ((idx / V) * G + lg) * V + (idx % V);
      if (idx < KV && col < D) sts_f32(park_s + col * 4, h[j]);
    }
    if (lane == 0) {
      park_tab[warp] = park_s;
      xwd[warp] = static_cast<double>(gb);
      xwd[kMaxWarps + warp] = lp;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= P; c += NW * 32) {
      double v[16];
#pragma unroll
      for (int wq = 0; wq < 16; ++wq) v[wq] = 0.0;
      if (c < D) {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = static_cast<double>(lds_f32<0>(park_tab[wq] + c * 4));
      } else if (c == P) {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = xwd[kMaxWarps + wq];
      } else {
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) v[wq] = xwd[wq];
      }
      cta_acc[c] = tree16(v);
    }
    fence_proxy_async_smem();  // the parked stages go back to the TMA (async proxy) after the barrier
    __syncthreads();
    if constexpr (!SPLIT) {
      if (deferred) ring_issue(pr, wt, ring, lane, policy);
    }
  }
}


// One row group: the warp's 32/G concurrent rows at shared address xaddr (already offset to this lane's row and
// chunk). Computes the linear predictor, the likelihood term and its residual, and accumulates residual * x.
// FAM >= 0 / LPM >= 0 fix the likelihood family and whether the log-likelihood is accumulated at compile time (the
// CTA-ring pass dispatches once per pass, so the tile loop carries no uniform branches); -1 = decided at run time.
// FAST: gradient-only Bernoulli passes use the special-function-unit residual (common.cuh bernoulli_resid_fast).
template <int G, int V, int K, bool WREG, bool SPLIT, bool CHECK = true, int NCH = 2, int FAM = -1, int LPM = -1,
          bool FAST = false>
__device__ __forceinline__ void row_group(uint32_t xaddr, uint32_t ys, int lr, int rows, int lg,
                                          const typename Acc<V>::type* w, uint32_t theta_lane_s, const float* theta_s,
                                          float bias, int family, float lik_scale, int y_dtype, bool want_lp,
                                          typename Acc<V>::type* g, float& gb, double& lp) {
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;
  using acc_t = typename Acc<V>::type;
  acc_t x[NA];
  ChunkLoader<V, K, G * V>::load(xaddr, x);
  const bool valid = CHECK ? (lr < rows) : true;
  const float yv = y_from_bits(lds_u32(ys + (valid ? lr : 0) * 4), y_dtype);
  float dotv;
  if constexpr (V == 1) {
    float a0 = 0.0f, a1 = 0.0f;
    float wl[WREG ? 1 : NA];
    if constexpr (!WREG) ChunkLoader<V, K, G * V>::load(theta_lane_s, wl);
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      float wi;
      if constexpr (WREG)
        wi = w[i];
      else
        wi = wl[i];
      if (i & 1)
        a1 = fmaf(x[i], wi, a1);
      else
        a0 = fmaf(x[i], wi, a0);
    }
    dotv = a0 + a1;
  } else {
    float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
    if constexpr (WREG && NCH == 4) {
      // four independent accumulation chains: halves the dependent-FMA latency of a long row slice
      float2 a2 = make_float2(0.0f, 0.0f), a3 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        if ((i & 3) == 0)
          a0 = fma2(x[i], w[i], a0);
        else if ((i & 3) == 1)
          a1 = fma2(x[i], w[i], a1);
        else if ((i & 3) == 2)
          a2 = fma2(x[i], w[i], a2);
        else
          a3 = fma2(x[i], w[i], a3);
      }
      a0.x += a2.x;
      a0.y += a2.y;
      a1.x += a3.x;
      a1.y += a3.y;
    } else if constexpr (WREG) {
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        if (i & 1)
          a1 = fma2(x[i], w[i], a1);
        else
          a0 = fma2(x[i], w[i], a0);
      }
    } else {
      // theta re-read from shared memory in 4 batches (keeps the live range short)
      constexpr int KB = (K + 3) / 4;
#pragma unroll
      for (int kb = 0; kb < K; kb += KB) {
        constexpr int NB = KB * (V / 2);
        acc_t wk[NB];
#pragma unroll
        for (int k = 0; k < KB; ++k)
          if (kb + k < K) load_chunks<V, 1>(theta_s + ((kb + k) * G + lg) * V, 0, &wk[k * (V / 2)]);
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const int i = kb * (V / 2) + q;
          if (i < NA) {
            if (q & 1)
              a1 = fma2(x[i], wk[q], a1);
            else
              a0 = fma2(x[i], wk[q], a0);
          }
        }
      }
    }
    dotv = (a0.x + a0.y) + (a1.x + a1.y);
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) dotv += __shfl_xor_sync(kFull, dotv, LaneMap<G, SPLIT>::lg_xor(off));
  const float eta = dotv + bias;
  float lpv = 0.0f, rv;
  const int fam = FAM >= 0 ? FAM : family;
  const bool wlp = LPM >= 0 ? (LPM != 0) : want_lp;
  if (wlp)  // CTA-uniform
    row_terms(fam, eta, yv, lik_scale, lpv, rv);
  else if (FAST && fam == 0)
    rv = bernoulli_resid_fast(eta, yv);
  else
    rv = row_resid(fam, eta, yv, lik_scale);
  if (CHECK && !valid) {
    lpv = 0.0f;
    rv = 0.0f;
  }
  if (lg == 0) {
    if (wlp) lp += static_cast<double>(lpv);
    gb += rv;
  }
  if constexpr (V == 1) {
#pragma unroll
    for (int i = 0; i < NA; ++i) g[i] = fmaf(rv, x[i], g[i]);
  } else {
    const float2 r2 = make_float2(rv, rv);
#pragma unroll
    for (int i = 0; i < NA; ++i) g[i] = fma2(r2, x[i], g[i]);
  }
}

// One pass of this CTA over its rows. On return cta_acc[0..D) = Σ r_n·X[n,:], cta_acc[D] = Σ r_n (if
// has_bias), cta_acc[P] = Σ log p(y_n|eta_n) over the CTA's rows, all float64, reduced in a fixed
// order (bitwise reproducible). Ends with a __syncthreads().
// Register cap per thread for NW warps of 32 threads with one CTA per SM.
__host__ __device__ constexpr int reg_cap(int nw) { return nw >= 16 ? 128 : (nw >= 12 ? 168 : 255); }

template <int G, int V, int K, int NW>
__device__ __forceinline__ void stream_pass(const KArgs& a, const PlanRegs& pr, const WarpTiles& wt, Ring& ring,
                                            const SmemLayout& sm, float bias, uint64_t policy, bool want_lp) {
  constexpr int RPS = 32 / G;  // rows processed concurrently by a warp
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;          // packed accumulators per lane
  constexpr bool WREG = (3 * KV + 40 <= reg_cap(NW));  // theta slice held in registers, else re-read from smem
  using acc_t = typename Acc<V>::type;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane & (G - 1), grp = lane / G;
  const int family = a.family;
  const float lik_scale = a.lik_scale;
  const int y_dtype = a.y_dtype;
  const float* theta_s = sm.theta_s;
  double* cta_acc = sm.cta_acc;

  acc_t g[NA];
  acc_t w[WREG ? NA : 1];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    if constexpr (V == 1)
      g[i] = 0.0f;
    else
      g[i] = make_float2(0.0f, 0.0f);
  }
  if constexpr (WREG) load_chunks<V, K>(theta_s + lg * V, G * V, w);
  const uint32_t theta_lane_s = smem_u32(theta_s) + lg * V * 4;
  float gb = 0.0f;
  double lp = 0.0;

  // Wide models (NW*D floats exceed the xw scratch): the cross-warp reduction parks each warp's column sums in
  // the ring stage that warp consumed last, so the re-arm of that one stage waits until the sums are read.
  constexpr bool kMayPark = NW * G * K * V > kXwFloats;  // D <= G*K*V: narrow kernels never park
  const bool park = kMayPark && NW * pr.D > kXwFloats;
  bool deferred = false;
  uint32_t park_s = ring.base_s;  // a warp without tiles never arms its ring: stage 0 is free
  const bool backward = pr.zigzag && (ring.cpass & 1);
  const int row_bytes = pr.ldx * 4;
  const uint32_t lane_off = grp * row_bytes + lg * V * 4;
  for (int kt = 0; kt < wt.nt; ++kt) {
    const int kk = backward ? (wt.nt - 1 - kt) : kt;
    const int rows = (kk == wt.nt - 1) ? wt.rows_last : pr.RT;
    const int m = static_cast<int>(reinterpret_cast<uintptr_t>(wt.x0 + kk * pr.xstride) >> 2) & 3;
    const uint32_t sb = ring.base_s + ring.stage * pr.stage_bytes;
    mbar_wait_s(ring.bars_s + ring.stage * 8, ring.parity);
    const uint32_t ys = sb + pr.y_off_bytes;
    uint32_t xaddr = sb + m * 4 + lane_off;

    for (int j0 = 0; j0 < rows; j0 += RPS, xaddr += RPS * row_bytes) {  // warp-uniform
      row_group<G, V, K, WREG, false>(xaddr, ys, j0 + grp, rows, lg, w, theta_lane_s, theta_s, bias, family, lik_scale,
                                      y_dtype, want_lp, g, gb, lp);
    }
    __syncwarp();
    if (++ring.stage == pr.S) {
      ring.stage = 0;
      ring.parity ^= 1u;
    }
    if (park && kt == wt.nt - 1) {
      park_s = sb;
      deferred = ring.qi < ring.q_total;
    } else if (ring.qi < ring.q_total) {
      ring_issue(pr, wt, ring, lane, policy);
    }
  }
  ++ring.cpass;

  pass_reduce<G, V, K, NW, false>(a, pr, wt, ring, sm, g, gb, lp, park, deferred, park_s, policy);
}

// Sums the per-CTA partials [ncta][P+1] (global, float64) in a fixed order into cta_acc[0..P].
// Every CTA that calls this gets bit-identical totals.
// Canonical order (shared with the leader protocol of chain.cuh, which computes the first level on group-leader CTAs):
//   level 1: G_g[c] = tree16 over the CTAs g*16 .. g*16+15 of column c;   level 2: total[c] = tree16 over G_0 .. G_15.
// `scratch` holds kXwFloats/2 doubles (the cross-warp scratch of pass_reduce, free by now). When 16 group sums of every
// (padded) column fit in it, thread (c, q) computes the groups q, q+nsl, ... (two groups = 32 independent loads in flight
// at a time) and thread c combines them; otherwise thread c walks all groups itself.
template <int B = 2>
__device__ __forceinline__ void reduce_partials(const double* part, int ncta, int P, double* cta_acc, double* scratch) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ncol = P + 1;
  const int ngroups = (ncta + kLlGroup - 1) / kLlGroup;
  int cpad = 32;
  while (cpad < ncol && cpad < nthr) cpad <<= 1;
  if (cpad >= ncol && cpad * kLlMaxGroups <= kXwFloats / 2) {
    const int nsl = nthr / cpad;
    const int c = tid % cpad, q = tid / cpad;
    if (c < ncol) {
      for (int g0 = q; g0 < kLlMaxGroups; g0 += B * nsl) {
        double v[B][16];
#pragma unroll
        for (int i = 0; i < B; ++i) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const int cc = (g0 + i * nsl) * kLlGroup + m;
            v[i][m] = (g0 + i * nsl < ngroups && cc < ncta) ? __ldcg(part + static_cast<size_t>(cc) * ncol + c) : 0.0;
          }
        }
#pragma unroll
        for (int i = 0; i < B; ++i)
          if (g0 + i * nsl < kLlMaxGroups) scratch[(g0 + i * nsl) * cpad + c] = tree16(v[i]);
      }
    }
    __syncthreads();
    if (tid < ncol) {
      double G[16];
#pragma unroll
      for (int g = 0; g < kLlMaxGroups; ++g) G[g] = scratch[g * cpad + tid];
      cta_acc[tid] = tree16(G);
    }
    __syncthreads();
    return;
  }
  for (int c = tid; c < ncol; c += nthr) {
    double G[kLlMaxGroups];
#pragma unroll
    for (int g = 0; g < kLlMaxGroups; ++g) G[g] = 0.0;
#pragma unroll
    for (int g0 = 0; g0 < kLlMaxGroups; g0 += B) {
      if (g0 < ngroups) {
        double v[B][16];
#pragma unroll
        for (int i = 0; i < B; ++i) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const int cc = (g0 + i) * kLlGroup + m;
            v[i][m] = cc < ncta ? __ldcg(part + static_cast<size_t>(cc) * ncol + c) : 0.0;
          }
        }
#pragma unroll
        for (int i = 0; i < B; ++i) G[g0 + i] = tree16(v[i]);
      }
    }
    cta_acc[c] = tree16(G);
  }
  __syncthreads();
}

}  // namespace edhmc
