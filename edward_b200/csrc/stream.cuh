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
// Work decomposition inside a warp: a row is handled by G lanes (G = 1, 4 or 32); lane `lg` of the
// group owns vector chunks k*G+lg (k < K, K a compile-time tier) of V floats each, i.e. 128-/64-/32-bit
// shared loads, and the arithmetic is packed FFMA2 (two fp32 FMAs per instruction) when V >= 2.
// theta is zero-padded to G*K*V entries, so padded columns contribute 0 to the dot product and their
// gradient accumulators are simply never written out.
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
  const long long units = (a.n_rows + 3) >> 2;  // 4-row units keep every warp's first row 16-byte aligned
  const long long u0 = units * gw / total_w, u1 = units * (gw + 1) / total_w;
  long long begin = u0 * 4;
  long long end = u1 * 4 < a.n_rows ? u1 * 4 : a.n_rows;
  if (end < begin) end = begin;
  WarpTiles w;
  w.x0 = a.X + begin * a.ldx;
  w.y0 = reinterpret_cast<const char*>(a.y) + begin * 4;
  const long long n = end - begin;
  w.nt = static_cast<int>((n + a.RT - 1) / a.RT);
  w.rows_last = w.nt ? static_cast<int>(n - static_cast<long long>(w.nt - 1) * a.RT) : 0;
  w.tail = (n > 0 && end == a.n_rows) ? 1 : 0;
  return w;
}

struct Ring {
  float* base;     // first stage of this warp
  uint64_t* bars;  // S mbarriers of this warp
  // consumer cursor
  int stage;
  uint32_t parity;
  int cpass;  // pass index of the tile being consumed (for the zig-zag order)
  // producer cursor
  long long qi, q_total;  // next tile to issue / tiles in this launch
  int ipass, ik, istage;
};

__device__ __forceinline__ void ring_init(Ring& ring, float* base, uint64_t* bars) {
  ring.base = base;
  ring.bars = bars;
  ring.stage = 0;
  ring.parity = 0;
  ring.cpass = 0;
  ring.qi = 0;
  ring.q_total = 0;
  ring.ipass = 0;
  ring.ik = 0;
  ring.istage = 0;
}

// Issues the next tile of this warp's sequence into stage `istage`. Called by ALL lanes of the warp
// (converged): every lane copies its share of the y slice, lane 0 launches the bulk copy of X.
// A tile whose first float is not 16-byte aligned (possible only when ldx % 4 != 0) is copied from the
// aligned address below it; the consumer skips the same `m` leading floats.
__device__ __forceinline__ void ring_issue(const KArgs& a, const WarpTiles& wt, Ring& ring, int lane, uint64_t policy) {
  const int kk = (a.zigzag && (ring.ipass & 1)) ? (wt.nt - 1 - ring.ik) : ring.ik;
  const bool last = (kk == wt.nt - 1);
  const int rows = last ? wt.rows_last : a.RT;
  float* sb = ring.base + ring.istage * a.stage_floats;
  uint64_t* bar = ring.bars + ring.istage;
  const char* ysrc = wt.y0 + static_cast<long long>(kk) * (a.RT * 4);
  for (int e = lane; e < rows; e += 32) cp_async4(sb + a.y_off + e, ysrc + e * 4);
  cp_async_mbar_arrive_noinc(bar);
  if (lane == 0) {
    const int m = (kk * a.tm) & 3;
    const float* src = wt.x0 + static_cast<long long>(kk) * a.tl - m;
    uint32_t bytes;
    if (last && wt.tail) {
      // never read past the last valid float of X: bulk-copy the 16-byte multiple, finish with scalar copies
      const int nfl = m + (rows - 1) * a.ldx_i + a.D;
      bytes = static_cast<uint32_t>(nfl * 4) & ~15u;
      for (int i = bytes >> 2; i < nfl; ++i) sb[i] = __ldg(src + i);
    } else {
      bytes = (static_cast<uint32_t>((m + rows * a.ldx_i) * 4) + 15u) & ~15u;
    }
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bar, bytes);
    if (bytes) {
      if (a.l2_hint)
        bulk_g2s_hint(sb, src, bytes, bar, policy);
      else
        bulk_g2s(sb, src, bytes, bar);
    }
  }
  ++ring.qi;
  if (++ring.ik == wt.nt) {
    ring.ik = 0;
    ++ring.ipass;
  }
  if (++ring.istage == a.S) ring.istage = 0;
}

// Fills the ring at the start of a launch (all lanes).
__device__ __forceinline__ void ring_prologue(const KArgs& a, const WarpTiles& wt, Ring& ring, long long n_passes, int lane,
                                              uint64_t policy) {
  ring.q_total = n_passes * wt.nt;
  for (int s = 0; s < a.S && ring.qi < ring.q_total; ++s) ring_issue(a, wt, ring, lane, policy);
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

template <int V, int NA>
__device__ __forceinline__ float flat(const typename Acc<V>::type* g, int i) {
  if constexpr (V == 1) {
    return g[i];
  } else {
    return (i & 1) ? g[i >> 1].y : g[i >> 1].x;
  }
}

// One pass of this CTA over its rows. On return cta_acc[0..D) = Σ r_n·X[n,:], cta_acc[D] = Σ r_n (if
// has_bias), cta_acc[P] = Σ log p(y_n|eta_n) over the CTA's rows, all float64, reduced in a fixed
// order (bitwise reproducible). Ends with a __syncthreads().
template <int G, int V, int K>
__device__ __forceinline__ void stream_pass(const KArgs& a, const WarpTiles& wt, Ring& ring, const float* theta_s,
                                            float bias, uint64_t policy, double* cta_acc) {
  constexpr int RPS = 32 / G;  // rows processed concurrently by a warp
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;  // packed accumulators per lane
  constexpr bool WREG = (KV <= 56);           // theta slice held in registers
  using acc_t = typename Acc<V>::type;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane & (G - 1), grp = lane / G;
  const int ldx = a.ldx_i;
  const int family = a.family;
  const float lik_scale = a.lik_scale;

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
  float gb = 0.0f;
  double lp = 0.0;

  const bool backward = a.zigzag && (ring.cpass & 1);
  for (int kt = 0; kt < wt.nt; ++kt) {
    const int kk = backward ? (wt.nt - 1 - kt) : kt;
    const int rows = (kk == wt.nt - 1) ? wt.rows_last : a.RT;
    const int m = (kk * a.tm) & 3;
    const float* sb = ring.base + ring.stage * a.stage_floats;
    mbar_wait(ring.bars + ring.stage, ring.parity);
    const uint32_t* ys = reinterpret_cast<const uint32_t*>(sb + a.y_off);
    const float* xbase = sb + m + grp * ldx + lg * V;

    for (int j0 = 0; j0 < rows; j0 += RPS) {  // warp-uniform
      const int lr = j0 + grp;
      acc_t x[NA];
      load_chunks<V, K>(xbase + j0 * ldx, G * V, x);
      float dotv;
      if constexpr (V == 1) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          float wi;
          if constexpr (WREG)
            wi = w[i];
          else
            wi = theta_s[(i * G + lg)];
          if (i & 1)
            a1 = fmaf(x[i], wi, a1);
          else
            a0 = fmaf(x[i], wi, a0);
        }
        dotv = a0 + a1;
      } else {
        float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
        if constexpr (WREG) {
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            if (i & 1)
              a1 = fma2(x[i], w[i], a1);
            else
              a0 = fma2(x[i], w[i], a0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            acc_t wk[V / 2];
            load_chunks<V, 1>(theta_s + (k * G + lg) * V, 0, wk);
#pragma unroll
            for (int q = 0; q < V / 2; ++q) {
              if (q & 1)
                a1 = fma2(x[k * (V / 2) + q], wk[q], a1);
              else
                a0 = fma2(x[k * (V / 2) + q], wk[q], a0);
            }
          }
        }
        dotv = (a0.x + a0.y) + (a1.x + a1.y);
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) dotv += __shfl_xor_sync(kFull, dotv, off);
      const float eta = dotv + bias;
      const bool valid = lr < rows;
      const float yv = y_from_bits(ys[valid ? lr : 0], a.y_dtype);
      float lpv, rv;
      row_terms(family, eta, yv, lik_scale, lpv, rv);
      if (!valid) {
        lpv = 0.0f;
        rv = 0.0f;
      }
      if (lg == 0) {
        lp += static_cast<double>(lpv);
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
    __syncwarp();
    if (++ring.stage == a.S) {
      ring.stage = 0;
      ring.parity ^= 1u;
    }
    if (ring.qi < ring.q_total) ring_issue(a, wt, ring, lane, policy);
  }
  ++ring.cpass;

  // ---- reduce across the RPS row groups of the warp with a halving butterfly: after stage `st` a lane
  //      keeps the half of the accumulators selected by its own bit, so 32+16+8+4+2 shuffles sum 64
  //      accumulators over 32 lanes (instead of 64*5). ----
  constexpr int NST = (RPS == 32) ? 5 : (RPS == 8 ? 3 : 0);
  constexpr int LP = (KV + RPS - 1) / RPS * RPS;  // padded length, divisible by RPS = 2^NST
  constexpr int LPF = LP / RPS;                   // accumulators a lane ends up owning
  float h[LP];
#pragma unroll
  for (int i = 0; i < LP; ++i) h[i] = (i < KV) ? flat<V, NA>(g, i) : 0.0f;
#pragma unroll
  for (int st = 0; st < NST; ++st) {
    const int off = 16 >> st;
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
  for (int wq = 0; wq < kWarpsPerCta; ++wq) {
    if (warp == wq) {
#pragma unroll
      for (int j = 0; j < LPF; ++j) {
        const int idx = grp * LPF + j;
        const int col = ((idx / V) * G + lg) * V + (idx % V);
        if (idx < KV && col < a.D) cta_acc[col] = (wq ? cta_acc[col] : 0.0) + static_cast<double>(h[j]);
      }
      if (lane == 0) {
        if (a.has_bias) cta_acc[a.D] = (wq ? cta_acc[a.D] : 0.0) + static_cast<double>(gb);
        cta_acc[a.P] = (wq ? cta_acc[a.P] : 0.0) + lp;
      }
    }
    __syncthreads();
  }
}

// Sums the per-CTA partials [ncta][P+1] (global, float64) in a fixed order into cta_acc[0..P].
// Every CTA that calls this gets bit-identical totals. `comb` holds kThreads doubles.
__device__ __forceinline__ void reduce_partials(const double* part, int ncta, int P, double* cta_acc, double* comb) {
  const int tid = threadIdx.x;
  const int ncol = P + 1;
  int cpad = 32;
  while (cpad < ncol) cpad <<= 1;
  if (cpad <= kThreads) {
    const int nsl = kThreads / cpad;
    const int c = tid % cpad, q = tid / cpad;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (c < ncol) {
      int cta = q;
      for (; cta + 3 * nsl < ncta; cta += 4 * nsl) {
        s0 += __ldcg(part + static_cast<size_t>(cta) * ncol + c);
        s1 += __ldcg(part + static_cast<size_t>(cta + nsl) * ncol + c);
        s2 += __ldcg(part + static_cast<size_t>(cta + 2 * nsl) * ncol + c);
        s3 += __ldcg(part + static_cast<size_t>(cta + 3 * nsl) * ncol + c);
      }
      for (; cta < ncta; cta += nsl) s0 += __ldcg(part + static_cast<size_t>(cta) * ncol + c);
    }
    comb[q * cpad + c] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < ncol) {
      double t = 0.0;
      for (int qq = 0; qq < nsl; ++qq) t += comb[qq * cpad + tid];
      cta_acc[tid] = t;
    }
    __syncthreads();
  } else {
    for (int c = tid; c < ncol; c += kThreads) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int cta = 0;
      for (; cta + 3 < ncta; cta += 4) {
        s0 += __ldcg(part + static_cast<size_t>(cta) * ncol + c);
        s1 += __ldcg(part + static_cast<size_t>(cta + 1) * ncol + c);
        s2 += __ldcg(part + static_cast<size_t>(cta + 2) * ncol + c);
        s3 += __ldcg(part + static_cast<size_t>(cta + 3) * ncol + c);
      }
      for (; cta < ncta; ++cta) s0 += __ldcg(part + static_cast<size_t>(cta) * ncol + c);
      cta_acc[c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
  }
}

// Shared-memory carve-up.
struct SmemLayout {
  float* ring;
  uint64_t* bars;
  double* cta_acc;
  double* red;
  double* comb;
  float* theta_s;
  float* state;  // 5 * ppad floats
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t smem_layout_bytes(int S, int stage_floats, int P, int wpad, size_t* offs /*7*/) {
  size_t off = 0;
  offs[0] = off;
  off += static_cast<size_t>(kWarpsPerCta) * S * stage_floats * 4;
  off = align_up(off, 128);
  offs[1] = off;
  off += static_cast<size_t>(kWarpsPerCta) * kMaxStages * 8;
  offs[2] = off;
  off += align_up(static_cast<size_t>(P + 1) * 8, 16);
  offs[3] = off;
  off += 64 * 8;
  offs[4] = off;
  off += static_cast<size_t>(kThreads) * 8;
  offs[5] = off;
  off += align_up(static_cast<size_t>(wpad) * 4, 16);
  offs[6] = off;
  off += 5 * align_up(static_cast<size_t>(P) * 4, 16);
  return align_up(off, 128);
}

__device__ __forceinline__ SmemLayout carve_smem(unsigned char* raw, const KArgs& a) {
  size_t offs[7];
  smem_layout_bytes(a.S, a.stage_floats, a.P, a.wpad, offs);
  SmemLayout L;
  L.ring = reinterpret_cast<float*>(raw + offs[0]);
  L.bars = reinterpret_cast<uint64_t*>(raw + offs[1]);
  L.cta_acc = reinterpret_cast<double*>(raw + offs[2]);
  L.red = reinterpret_cast<double*>(raw + offs[3]);
  L.comb = reinterpret_cast<double*>(raw + offs[4]);
  L.theta_s = reinterpret_cast<float*>(raw + offs[5]);
  L.state = reinterpret_cast<float*>(raw + offs[6]);
  return L;
}

// Zero the ring (so padded / stale columns are finite), init the mbarriers, zero theta_s.
__device__ __forceinline__ void smem_setup(const SmemLayout& L, const KArgs& a) {
  const int tid = threadIdx.x;
  const int nring = kWarpsPerCta * a.S * a.stage_floats;
  for (int i = tid; i < nring; i += kThreads) L.ring[i] = 0.0f;
  for (int i = tid; i < a.wpad; i += kThreads) L.theta_s[i] = 0.0f;
  if (tid < kWarpsPerCta * kMaxStages) mbar_init(L.bars + tid, 33);  // 32 cp.async arrivals + 1 expect_tx arrival
  fence_mbar_init();
  fence_proxy_async_smem();
  __syncthreads();
}

}  // namespace edhmc
