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
// group owns vector chunks k*G+lg (k < Kact) of V floats each, i.e. 128-/64-/32-bit shared loads.
// theta sits zero-padded in shared memory, so padded columns contribute 0 to the dot product.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace edhmc {

struct WarpRows {
  long long begin, end;  // rows [begin, end) of this warp
  int nt;                // tiles per pass
};

__device__ __forceinline__ WarpRows warp_rows(const KArgs& a, int gw, int total_w) {
  const long long units = (a.n_rows + 3) >> 2;  // 4-row units keep every tile start 16-byte aligned
  const long long u0 = units * gw / total_w, u1 = units * (gw + 1) / total_w;
  WarpRows w;
  w.begin = u0 * 4;
  w.end = u1 * 4 < a.n_rows ? u1 * 4 : a.n_rows;
  if (w.end < w.begin) w.end = w.begin;
  w.nt = static_cast<int>((w.end - w.begin + a.RT - 1) / a.RT);
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

__device__ __forceinline__ void tile_geometry(const KArgs& a, const WarpRows& wr, int pass, int kt, long long& row0,
                                              int& rows) {
  const int kk = (a.zigzag && (pass & 1)) ? (wr.nt - 1 - kt) : kt;
  row0 = wr.begin + static_cast<long long>(kk) * a.RT;
  const long long left = wr.end - row0;
  rows = left < a.RT ? static_cast<int>(left) : a.RT;
}

// Issues the next tile of this warp's sequence into stage `istage`. Called by ALL lanes of the warp
// (converged): every lane copies its share of the y slice, lane 0 launches the bulk copy of X.
__device__ __forceinline__ void ring_issue(const KArgs& a, const WarpRows& wr, Ring& ring, int lane, uint64_t policy) {
  long long row0;
  int rows;
  tile_geometry(a, wr, ring.ipass, ring.ik, row0, rows);
  float* sb = ring.base + static_cast<size_t>(ring.istage) * a.stage_floats;
  uint64_t* bar = ring.bars + ring.istage;
  const char* ysrc = reinterpret_cast<const char*>(a.y) + row0 * 4;
  for (int e = lane; e < rows; e += 32) cp_async4(sb + a.y_off + e, ysrc + static_cast<size_t>(e) * 4);
  cp_async_mbar_arrive_noinc(bar);
  if (lane == 0) {
    const float* src = a.X + row0 * a.ldx;
    // never read past the last valid float of X (the last row has only D valid floats)
    const long long nfl = (row0 + rows == a.n_rows) ? static_cast<long long>(rows - 1) * a.ldx + a.D
                                                    : static_cast<long long>(rows) * a.ldx;
    const uint32_t b16 = static_cast<uint32_t>((nfl * 4) & ~15LL);
    for (long long i = b16 >> 2; i < nfl; ++i) sb[i] = __ldg(src + i);
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bar, b16);
    if (b16) {
      if (a.l2_hint)
        bulk_g2s_hint(sb, src, b16, bar, policy);
      else
        bulk_g2s(sb, src, b16, bar);
    }
  }
  ++ring.qi;
  if (++ring.ik == wr.nt) {
    ring.ik = 0;
    ++ring.ipass;
  }
  if (++ring.istage == a.S) ring.istage = 0;
}

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

// Fills the ring at the start of a launch (all lanes).
__device__ __forceinline__ void ring_prologue(const KArgs& a, const WarpRows& wr, Ring& ring, long long n_passes, int lane,
                                              uint64_t policy) {
  ring.q_total = n_passes * wr.nt;
  for (int s = 0; s < a.S && ring.qi < ring.q_total; ++s) ring_issue(a, wr, ring, lane, policy);
}

template <int V>
__device__ __forceinline__ void lds_vec(const float* p, float* out) {
  if constexpr (V == 1) {
    out[0] = *p;
  } else if constexpr (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    out[0] = t.x;
    out[1] = t.y;
  } else {
    const float4 t = *reinterpret_cast<const float4*>(p);
    out[0] = t.x;
    out[1] = t.y;
    out[2] = t.z;
    out[3] = t.w;
  }
}

// One pass of this CTA over its rows. On return cta_acc[0..D) = Σ r_n·X[n,:], cta_acc[D] = Σ r_n (if
// has_bias), cta_acc[P] = Σ log p(y_n|eta_n) over the CTA's rows, all float64, reduced in a fixed
// order (bitwise reproducible). Ends with a __syncthreads().
template <int G, int V, int KMAX>
__device__ __forceinline__ void stream_pass(const KArgs& a, const WarpRows& wr, Ring& ring, const float* theta_s,
                                            float bias, uint64_t policy, double* cta_acc) {
  constexpr int RPS = 32 / G;  // rows processed concurrently by a warp
  constexpr int KV = KMAX * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane % G, grp = lane / G;
  const int ldx = static_cast<int>(a.ldx);
  const int Kact = a.Kact;

  float g[KV];
#pragma unroll
  for (int i = 0; i < KV; ++i) g[i] = 0.0f;
  float gb = 0.0f;
  double lp = 0.0;

  for (int kt = 0; kt < wr.nt; ++kt) {
    long long row0;
    int rows;
    tile_geometry(a, wr, ring.cpass, kt, row0, rows);
    const float* sb = ring.base + static_cast<size_t>(ring.stage) * a.stage_floats;
    mbar_wait(ring.bars + ring.stage, ring.parity);
    const uint32_t* ys = reinterpret_cast<const uint32_t*>(sb + a.y_off);

    for (int j = 0; j < a.J; ++j) {
      if (j * RPS >= rows) break;  // warp-uniform
      const int lr = j * RPS + grp;
      const float* xr = sb + lr * ldx + lg * V;
      float x[KV];
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < Kact) lds_vec<V>(xr + k * G * V, &x[k * V]);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < Kact) {
          float wv[V];
          lds_vec<V>(theta_s + (k * G + lg) * V, wv);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (((k * V + v) & 1) == 0)
              acc0 = fmaf(x[k * V + v], wv[v], acc0);
            else
              acc1 = fmaf(x[k * V + v], wv[v], acc1);
          }
        }
      float dotv = acc0 + acc1;
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) dotv += __shfl_xor_sync(kFull, dotv, off);
      const float eta = dotv + bias;
      const bool valid = lr < rows;
      const float yv = y_from_bits(ys[valid ? lr : 0], a.y_dtype);
      float lpv, rv;
      row_terms(a.family, eta, yv, a.lik_scale, lpv, rv);
      if (!valid) {
        lpv = 0.0f;
        rv = 0.0f;
      }
      if (lg == 0) {
        lp += static_cast<double>(lpv);
        gb += rv;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < Kact) {
#pragma unroll
          for (int v = 0; v < V; ++v) g[k * V + v] = fmaf(rv, x[k * V + v], g[k * V + v]);
        }
    }
    __syncwarp();
    if (++ring.stage == a.S) {
      ring.stage = 0;
      ring.parity ^= 1u;
    }
    if (ring.qi < ring.q_total) ring_issue(a, wr, ring, lane, policy);
  }
  ++ring.cpass;

  // ---- reduce across the row groups of the warp (lanes with equal lg) ----
#pragma unroll
  for (int off = G; off < 32; off <<= 1) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < Kact) {
#pragma unroll
        for (int v = 0; v < V; ++v) g[k * V + v] += __shfl_xor_sync(kFull, g[k * V + v], off);
      }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) gb += __shfl_xor_sync(kFull, gb, off);
  lp = warp_sum_f64(lp);

  // ---- reduce across the warps of the CTA: fixed order, float64 ----
  for (int wq = 0; wq < kWarpsPerCta; ++wq) {
    if (warp == wq) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < Kact) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            // every replica of the lane group holds the same totals: spread the writes over replicas
            if (grp == ((k * V + v) % RPS)) {
              const int col = (k * G + lg) * V + v;
              if (col < a.D) cta_acc[col] = (wq ? cta_acc[col] : 0.0) + static_cast<double>(g[k * V + v]);
            }
          }
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

// Shared-memory carve-up common to both plans.
struct SmemLayout {
  float* ring;
  uint64_t* bars;
  double* cta_acc;
  double* red;
  double* comb;
  float* theta_s;
  float* state;  // 5 * ppad floats (persistent plan only)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t smem_layout_bytes(int S, int stage_floats, int P, int wpad, bool with_state,
                                                    size_t* offs /*7*/) {
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
  if (with_state) off += 5 * align_up(static_cast<size_t>(P) * 4, 16);
  return align_up(off, 128);
}

__device__ __forceinline__ SmemLayout carve_smem(unsigned char* raw, const KArgs& a, bool with_state) {
  size_t offs[7];
  smem_layout_bytes(a.S, a.stage_floats, a.P, a.wpad, with_state, offs);
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
  const size_t nring = static_cast<size_t>(kWarpsPerCta) * a.S * a.stage_floats;
  for (size_t i = tid; i < nring; i += kThreads) L.ring[i] = 0.0f;
  for (int i = tid; i < a.wpad; i += kThreads) L.theta_s[i] = 0.0f;
  if (tid < kWarpsPerCta * kMaxStages) mbar_init(L.bars + tid, 33);  // 32 cp.async arrivals + 1 expect_tx arrival
  fence_mbar_init();
  fence_proxy_async_smem();
  __syncthreads();
}

}  // namespace edhmc
