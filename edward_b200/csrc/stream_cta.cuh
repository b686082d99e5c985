// stream_cta.cuh — the streaming pass with ONE ring per CTA (ring mode 1), for narrow rows.
//
// With the per-warp rings of stream.cuh every 32-row tile of a narrow matrix (cfg 2: 54 floats per row, one row per
// lane) costs the consuming warp ~300 instructions of ring bookkeeping next to ~130 instructions of arithmetic, and the
// pass is bound by instruction issue at two warps per scheduler (ncu, profiles/r01g_cfg2_*). Here a stage holds
// NW * (32/G) * J consecutive rows — one bulk copy of X and one of y per stage — and every warp consumes its own
// J row groups of each stage. The warp that finishes a stage LAST (a shared-memory counter) re-arms it with the tile
// S positions further down the CTA's sequence, so there is no producer warp, no "empty" barrier to wait on and the
// consumers' per-tile overhead is a barrier wait, one atomic and the cursor update.
//
// Lane map: split (stream.cuh LaneMap<G, true>): lanes of one shared-memory wavefront read the same chunk of consecutive
// rows, so spreading a row over G = 2 or 4 lanes (to get under 128 registers and run 16 warps per SM) keeps the
// conflict-free bank pattern of one lane per row.
//
// The CTA's rows are one contiguous range in units of 4 rows, so every tile starts 16-byte aligned in X and in y
// whatever the row stride is.
#pragma once
#include "stream.cuh"

namespace edhmc {

struct CtaRing {
  uint32_t base_s;  // shared-window address of stage 0
  uint32_t bars_s;  // S "full" mbarriers (one arrival: the arming thread's expect_tx)
  uint32_t cnt_s;   // S consumer counters (u32)
  int stage;
  uint32_t parity;
  int cpass;            // parity source of the pass being consumed (zig-zag direction)
  int par0;             // cpass of the first pass of this launch
  long long q, q_total;  // sequence number of the tile being consumed / tiles in this launch
  int npass, nk;         // (pass, tile) coordinates of sequence number q + S: the tile that refills the current stage
};

// Arms stage `stage` with tile `k` of a pass of direction parity `par`. Executed by ONE thread.
__device__ __forceinline__ void cta_ring_issue(const PlanRegs& pr, const WarpTiles& ct, const CtaRing& ring, int stage, int par,
                                               int k, uint64_t policy) {
  const int kk = (pr.zigzag && (par & 1)) ? (ct.nt - 1 - k) : k;
  const bool last = (kk == ct.nt - 1);
  const int rows = last ? ct.rows_last : pr.RT;
  const uint32_t sb = ring.base_s + stage * pr.stage_bytes;
  const uint32_t bar = ring.bars_s + stage * 8;
  const float* src = ct.x0 + kk * pr.xstride;
  const char* ysrc = ct.y0 + kk * pr.ystride;
  uint32_t xb, yb;
  if (last && ct.tail) {
    // never read past the last valid float of X / entry of y: bulk-copy the 16-byte multiples, finish with scalar copies
    const int nfl = (rows - 1) * pr.ldx + pr.D;
    xb = static_cast<uint32_t>(nfl * 4) & ~15u;
    for (int i = xb >> 2; i < nfl; ++i) sts_f32(sb + i * 4, __ldg(src + i));
    yb = static_cast<uint32_t>(rows * 4) & ~15u;
    for (int i = yb >> 2; i < rows; ++i)
      sts_f32(sb + pr.y_off_bytes + i * 4, __ldg(reinterpret_cast<const float*>(ysrc) + i));
  } else {
    xb = static_cast<uint32_t>(rows * pr.ldx * 4);  // rows % 4 == 0 here: a multiple of 16
    yb = static_cast<uint32_t>(rows * 4);
  }
  fence_proxy_async_smem();
  mbar_arrive_expect_tx_s(bar, xb + yb);
  if (xb) {
    if (pr.l2_hint)
      bulk_g2s_hint_s(sb, src, xb, bar, policy);
    else
      bulk_g2s_s(sb, src, xb, bar);
  }
  if (yb) bulk_g2s_s(sb + pr.y_off_bytes, ysrc, yb, bar);
}

__device__ __forceinline__ void cta_ring_init(CtaRing& ring, const SmemLayout& sm, const KArgs& a, int group) {
  ring.base_s = smem_u32(sm.ring + static_cast<size_t>(group) * a.S * a.stage_floats);
  ring.bars_s = smem_u32(sm.bars + group * kMaxStages);
  ring.cnt_s = smem_u32(sm.bars + kMaxWarps * kMaxStages) + group * kMaxStages * 4;
  ring.stage = 0;
  ring.parity = 0;
  ring.cpass = 0;
  ring.par0 = 0;
  ring.q = 0;
  ring.q_total = 0;
  ring.npass = 0;
  ring.nk = 0;
}

// Shared-memory setup of ring mode 1: zero the stages (stale / padded columns must be finite), per group S barriers
// with one arrival each and S counters.
__device__ __forceinline__ void smem_setup_cta(const SmemLayout& L, const KArgs& a, int ngroups) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nring = ngroups * a.S * a.stage_floats;
  for (int i = tid; i < nring; i += nthr) L.ring[i] = 0.0f;
  for (int i = tid; i < a.wpad; i += nthr) L.theta_s[i] = 0.0f;
  if (tid < kMaxWarps * kMaxStages) {
    mbar_init(L.bars + tid, 1);
    reinterpret_cast<uint32_t*>(L.bars + kMaxWarps * kMaxStages)[tid] = 0u;
  }
  fence_mbar_init();
  fence_proxy_async_smem();
  __syncthreads();
}

// Fills the ring at the start of a launch (the group's first thread arms; every thread sets its cursors).
__device__ __forceinline__ void cta_ring_prologue(const PlanRegs& pr, const WarpTiles& ct, CtaRing& ring, long long n_passes,
                                                  uint64_t policy, bool issuer) {
  ring.q_total = n_passes * ct.nt;
  int ip = 0, ik = 0;
  for (int s = 0; s < pr.S; ++s) {
    if (s < ring.q_total && issuer) cta_ring_issue(pr, ct, ring, s, ring.par0 + ip, ik, policy);
    if (ct.nt > 0 && ++ik == ct.nt) {
      ik = 0;
      ++ip;
    }
  }
  ring.npass = ip;
  ring.nk = ik;
}

__device__ __forceinline__ uint32_t atom_add_acq_rel_smem(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// One pass of this CTA over its rows (ring mode 1). Same contract as stream_pass: on return cta_acc[0..P] holds the
// CTA's float64 sums, reduced in a fixed order; ends with a __syncthreads().
//
// NOT inlined on purpose: inside the persistent kernel the chain state (cursors, float64 scalars, plan constants) is
// live across the pass, and with everything inlined the register allocator answered with rematerialisation inside
// the tile loop (262 instructions per 32 rows instead of ~140, ncu r02). As a separate function the loop gets a clean
// register file; the caller's state is spilled once per pass around the call, which costs nothing next to a pass.
struct CtaPassCtx {
  uint32_t ring_s, bars_s, cnt_s, theta_s;
  uint32_t y_off_bytes;
  uint32_t stage_bytes;
  uint32_t row_bytes;
  int RT, S, J, nt, rows_last, backward, wpg;
  int family, y_dtype;
  float lik_scale, bias;
  int want_lp;
  // ring cursor (in/out)
  int stage;
  uint32_t parity;
  long long q, q_total;
  int npass, nk, par0;
  // development timeline
  long long* tl_wait;  // where warp w adds its tile-wait cycles (slot w), or nullptr
};

#ifndef EDHMC_PASS_INLINE
#define EDHMC_PASS_INLINE __forceinline__
#endif
template <int G, int V, int K, int NW, int FAM, int LPM>
__device__ EDHMC_PASS_INLINE void stream_pass_cta_tiles(CtaPassCtx* __restrict__ cx, const PlanRegs* __restrict__ prp,
                                                  const WarpTiles* __restrict__ ctp, const CtaRing* __restrict__ ringp,
                                                  uint64_t policy, typename Acc<V>::type* __restrict__ gout, float* gbout,
                                                  double* lpout) {
  constexpr int RPS = 32 / G;
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;
  constexpr bool WREG = (3 * KV + 40 <= reg_cap(NW));
  using acc_t = typename Acc<V>::type;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = LaneMap<G, true>::lg(lane), grp = LaneMap<G, true>::grp(lane);
  // scalars of the loop, copied to registers once
  const uint32_t ring_s = cx->ring_s, bars_s = cx->bars_s, cnt_s = cx->cnt_s;
  const uint32_t stage_bytes = cx->stage_bytes, y_off_bytes = cx->y_off_bytes, row_bytes = cx->row_bytes;
  const int RT = cx->RT, S = cx->S, J = cx->J, nt = cx->nt, rows_last = cx->rows_last, wpg = cx->wpg;
  const bool backward = cx->backward != 0;
  const int family = cx->family, y_dtype = cx->y_dtype;
  const float lik_scale = cx->lik_scale, bias = cx->bias;
  const bool want_lp = cx->want_lp != 0;
  int stage = cx->stage;
  uint32_t parity = cx->parity;
  long long q = cx->q;
  const long long q_total = cx->q_total;
  int npass = cx->npass, nk = cx->nk;
  const int par0 = cx->par0;
  const float* theta_gen = static_cast<const float*>(__cvta_shared_to_generic(cx->theta_s));

  acc_t g[NA];
  acc_t w[WREG ? NA : 1];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    if constexpr (V == 1)
      g[i] = 0.0f;
    else
      g[i] = make_float2(0.0f, 0.0f);
  }
  const uint32_t theta_lane_s = cx->theta_s + lg * V * 4;
  if constexpr (WREG) ChunkLoader<V, K, G * V>::load(theta_lane_s, w);
  float gb = 0.0f;
  double lp = 0.0;

  const int wrow0 = (warp % wpg) * J * RPS;  // this warp's first row inside a stage of its group
  const uint32_t lane_off = (wrow0 + grp) * row_bytes + lg * V * 4;
  const bool acct = cx->tl_wait != nullptr && lane == 0 && warp < 8;
  long long wait_cyc = 0;
  if (cx->tl_wait != nullptr && threadIdx.x == 0) cx->tl_wait[8] = clock64();  // slot 16: tile loop starts
  for (int kt = 0; kt < nt; ++kt) {
    const int kk = backward ? (nt - 1 - kt) : kt;
    const uint32_t sb = ring_s + stage * stage_bytes;
    const long long tw0 = acct ? clock64() : 0;
    mbar_wait_s(bars_s + stage * 8, parity);
    if (acct) wait_cyc += clock64() - tw0;
    const uint32_t ys = sb + y_off_bytes;
    uint32_t xaddr = sb + lane_off;
    if (kk != nt - 1 || rows_last == RT) {
      // full tile: every row group of every warp is complete
      for (int j = 0; j < J; ++j, xaddr += RPS * row_bytes)
        row_group<G, V, K, WREG, true, false, 4, FAM, LPM, true>(xaddr, ys, wrow0 + j * RPS + grp, RT, lg, w, theta_lane_s, theta_gen, bias,
                                                 family, lik_scale, y_dtype, want_lp, g, gb, lp);
    } else if (wrow0 < rows_last) {  // partial last tile; warp-uniform: upper warps may have no rows at all
      for (int j = 0; j < J; ++j, xaddr += RPS * row_bytes)
        row_group<G, V, K, WREG, true, true, 4, FAM, LPM, true>(xaddr, ys, wrow0 + j * RPS + grp, rows_last, lg, w, theta_lane_s, theta_gen,
                                                bias, family, lik_scale, y_dtype, want_lp, g, gb, lp);
    }
    __syncwarp();
    if (lane == 0) {
      const uint32_t cnt = cnt_s + stage * 4;
      if (atom_add_acq_rel_smem(cnt, 1u) == static_cast<uint32_t>(wpg - 1)) {  // last warp of the group re-arms the stage
        sts_u32(cnt, 0u);
        if (q + S < q_total) cta_ring_issue(*prp, *ctp, *ringp, stage, par0 + npass, nk, policy);
      }
    }
    ++q;
    if (++nk == nt) {
      nk = 0;
      ++npass;
    }
    if (++stage == S) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (acct) cx->tl_wait[warp] = wait_cyc;
  if (cx->tl_wait != nullptr && threadIdx.x == 0) cx->tl_wait[9] = clock64();  // slot 17: tile loop done (warp 0)
  // cursor back to the caller (every thread holds the same values; its own copy of the context is per thread)
  cx->stage = stage;
  cx->parity = parity;
  cx->q = q;
  cx->npass = npass;
  cx->nk = nk;
#pragma unroll
  for (int i = 0; i < NA; ++i) gout[i] = g[i];
  *gbout = gb;
  *lpout = lp;
}


// Dispatches once per pass on (family, log-likelihood wanted) so that the tile loop itself is branch-free.
template <int G, int V, int K, int NW>
__device__ __forceinline__ void stream_pass_cta_loop(CtaPassCtx* __restrict__ cx, const PlanRegs* __restrict__ prp,
                                                     const WarpTiles* __restrict__ ctp, const CtaRing* __restrict__ ringp,
                                                     uint64_t policy, typename Acc<V>::type* __restrict__ gout, float* gbout,
                                                     double* lpout) {
  const int fam = cx->family;
  if (cx->want_lp) {
    if (fam == 0)
      stream_pass_cta_tiles<G, V, K, NW, 0, 1>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
    else if (fam == 1)
      stream_pass_cta_tiles<G, V, K, NW, 1, 1>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
    else
      stream_pass_cta_tiles<G, V, K, NW, 2, 1>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
  } else {
    if (fam == 0)
      stream_pass_cta_tiles<G, V, K, NW, 0, 0>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
    else if (fam == 1)
      stream_pass_cta_tiles<G, V, K, NW, 1, 0>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
    else
      stream_pass_cta_tiles<G, V, K, NW, 2, 0>(cx, prp, ctp, ringp, policy, gout, gbout, lpout);
  }
}

template <int G, int V, int K, int NW>
__device__ __forceinline__ void stream_pass_cta(const KArgs& a, const PlanRegs& pr, const WarpTiles& ct, CtaRing& ring,
                                                const SmemLayout& sm, float bias, uint64_t policy, bool want_lp) {
  constexpr int KV = K * V;
  constexpr int NA = (V == 1) ? KV : KV / 2;
  static_assert(NW * G * K * V <= kXwFloats, "ring mode 1 is for narrow rows (one-shot cross-warp reduction)");
  using acc_t = typename Acc<V>::type;
  CtaPassCtx cx;
  cx.ring_s = ring.base_s;
  cx.bars_s = ring.bars_s;
  cx.cnt_s = ring.cnt_s;
  cx.theta_s = smem_u32(sm.theta_s);
  cx.y_off_bytes = pr.y_off_bytes;
  cx.stage_bytes = pr.stage_bytes;
  cx.row_bytes = pr.ldx * 4;
  cx.RT = pr.RT;
  cx.S = pr.S;
  cx.J = a.J;
  cx.wpg = a.wpg;
  cx.nt = ct.nt;
  cx.rows_last = ct.rows_last;
  cx.backward = (pr.zigzag && (ring.cpass & 1)) ? 1 : 0;
  cx.family = a.family;
  cx.y_dtype = a.y_dtype;
  cx.lik_scale = a.lik_scale;
  cx.bias = bias;
  cx.want_lp = want_lp ? 1 : 0;
  cx.stage = ring.stage;
  cx.parity = ring.parity;
  cx.q = ring.q;
  cx.q_total = ring.q_total;
  cx.npass = ring.npass;
  cx.nk = ring.nk;
  cx.par0 = ring.par0;
  const long long tl_pass = ring.cpass - ring.par0;  // pass index within the launch
  cx.tl_wait = (a.timeline != nullptr && a.mode == 0 && tl_pass < a.tl_cap)
                   ? a.timeline + (static_cast<size_t>(tl_pass) * gridDim.x + blockIdx.x) * kTlRec + 8
                   : nullptr;
  acc_t g[NA];
  float gb;
  double lp;
  stream_pass_cta_loop<G, V, K, NW>(&cx, &pr, &ct, &ring, policy, g, &gb, &lp);
  ring.stage = cx.stage;
  ring.parity = cx.parity;
  ring.q = cx.q;
  ring.npass = cx.npass;
  ring.nk = cx.nk;
  ++ring.cpass;
  pass_reduce<G, V, K, NW, true>(a, pr, ct, ring, sm, g, gb, lp, false, false, 0u, policy, cx.tl_wait);
}

}  // namespace edhmc
