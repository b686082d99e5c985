// chain.cuh — the HMC transition (hmc.py:61-130) and leapfrog integrator (hmc.py:195-210) on the device.
//
// k_hmc<G,V,K> is the one kernel that touches X. It has two launch modes:
//   mode 0 (persistent plan, cooperative launch): runs ALL transitions of an edhmc_run. Every CTA carries
//     a redundant copy of the O(P) chain state in shared memory and performs the same integrator / accept
//     arithmetic on the same all-CTA totals, so the only grid-wide communication per leapfrog step is the
//     exchange of the per-CTA partial sums: on one GPU with a narrow model as fence-free flag-in-data entries
//     (partials -> sums of 16-CTA groups -> every CTA; "flat" protocol, a.leader == 2), otherwise one grid
//     barrier + (row shards) the in-kernel all-reduce over peer memory. The data pass between two exchanges
//     is one of three layouts chosen by the host (RM): a TMA ring per warp, a TMA ring per CTA, or — narrow
//     rows — re-laid tiles read with LDG, with as many rows as fit parked in shared and tensor memory for
//     the whole launch (stream_ldg.cuh).
//   mode 1 (stepwise plan): one data pass at a given theta; the last CTA to finish folds the partials
//     into a.sums. Single-CTA chain kernels (below) and, when rows are sharded over GPUs, an NCCL
//     all-reduce of a.sums run between passes.
//
// Schedule: one gradient evaluation per leapfrog step. The reference evaluates L+1 gradients and 2
// extra forward passes per transition (hmc.py:199,206,104-105); the values it recomputes — gradient
// and log joint of the transition's start state — are cached here from the previous transition.
#pragma once
#include "stream_cta.cuh"
#include "stream_ldg.cuh"

namespace edhmc {

// leapfrog pieces, float32 with separate multiply and add exactly as the TF graph has them
// (hmc.py:203-204,208: r + 0.5*step_size*grad ; z + step_size*r).
__device__ __forceinline__ float kick(float r, float half_eps, float g) { return __fadd_rn(r, __fmul_rn(half_eps, g)); }
__device__ __forceinline__ float drift(float z, float eps, float r) { return __fadd_rn(z, __fmul_rn(eps, r)); }

struct ChainRegs {
  double logp_cur, logp_new, k_old, log_u;
  long long it, n_acc;
  int s, nonfinite, in_init, pad;
};

struct AcceptResult {
  double ratio, log_u;
  bool accept;
};
// hmc.py:100-109: ratio = K(r0) - K(rL) + logp(zL) - logp(z0) accumulated in that order; accept = log(u) < ratio.
__device__ __forceinline__ AcceptResult mh_accept(double k_old, double k_new, double logp_new, double logp_old,
                                                  double log_u) {
  AcceptResult a;
  a.ratio = ((k_old - k_new) + logp_new) - logp_old;
  a.log_u = log_u;
  a.accept = a.log_u < a.ratio;
  return a;
}

__device__ __forceinline__ void write_trace(const KArgs& a, long long it, double logp_old, double logp_new, double k_old,
                                            double k_new, const AcceptResult& ar) {
  if (a.trace_scalars) {
    double* ts = a.trace_scalars + it * 8;
    ts[0] = logp_old;
    ts[1] = logp_new;
    ts[2] = k_old;
    ts[3] = k_new;
    ts[4] = ar.ratio;
    ts[5] = ar.log_u;
    ts[6] = ar.accept ? 1.0 : 0.0;
    ts[7] = 0.0;
  }
}

// One-shot all-reduce of the shard totals cta_acc[0..P] over the ranks of an NVLink domain, inside the
// persistent kernel. CTA 0 stores this rank's totals into entry `rank` of EVERY rank's inbox (peer-mapped
// memory, plain stores over NVLink) and then releases one flag per destination; every CTA of every rank waits
// for the nranks flags of its own inbox and adds the entries in rank order — so all CTAs on all GPUs hold
// bit-identical sums, and the chain needs no further broadcast. Two slots (parity of the sequence number)
// suffice: a rank can run at most one pass ahead of the slowest one, because its next totals depend on
// everyone's current ones. Returns false if a wait timed out (a peer died): the kernel then unwinds.
struct PeerArgs {  // passed by value: taking the address of the kernel's KArgs would force a local copy of it
  unsigned char* const* inbox;
  int* abort_flag;
  long long spin_limit;
  int nranks, rank, ncol;
};
static __device__ __noinline__ bool peer_allreduce(const PeerArgs a, unsigned long long seq, double* cta_acc, int* s_flag) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int ncol = a.ncol, R = a.nranks;
  const int slot = static_cast<int>(seq & 1ull);
  if (blockIdx.x == 0) {
    for (int idx = tid; idx < R * ncol; idx += nthr) {
      const int dst = idx / ncol, c = idx - dst * ncol;
      double* data = reinterpret_cast<double*>(a.inbox[dst] + kInboxDataOff);
      data[(static_cast<size_t>(slot) * kMaxRanks + a.rank) * kInboxStride + c] = cta_acc[c];
    }
    __syncthreads();
    if (tid < R) {
      __threadfence_system();  // cumulative over the CTA's stores (bar.sync above)
      st_release_sys_u64(reinterpret_cast<unsigned long long*>(a.inbox[tid]) + slot * kMaxRanks + a.rank, seq);
    }
  }
  const unsigned char* own = a.inbox[a.rank];
  if (tid == 0) *s_flag = 0;
  __syncthreads();
  if (tid < R) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(own) + slot * kMaxRanks + tid;
    const long long t_start = clock64();
    unsigned int n = 0;
    while (ld_acquire_sys_u64(f) < seq) {
      if ((++n & 255u) == 0u) {
        if (clock64() - t_start > a.spin_limit) atomicExch(a.abort_flag, 1);
        if (ld_volatile_s32(a.abort_flag)) {
          *s_flag = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
  if (*s_flag) return false;
  const double* data = reinterpret_cast<const double*>(own + kInboxDataOff) + static_cast<size_t>(slot) * kMaxRanks * kInboxStride;
  for (int c = tid; c < ncol; c += nthr) {
    double sum = 0.0;
    for (int rr = 0; rr < R; ++rr) sum += __ldcg(data + static_cast<size_t>(rr) * kInboxStride + c);
    cta_acc[c] = sum;
  }
  __syncthreads();
  return true;
}

// Canonical summation order of a wide column (P+1 > kWideCols) over the per-CTA partials: lane l adds the CTAs
// l, l+32, ... in ascending order, then an xor butterfly (the same tree in every lane). Both plans use it, so
// they stay bit-identical.
__device__ __forceinline__ double wide_column_sum(const double* part, int ncta, int ncol, int c, int lane) {
  double s = 0.0;
  for (int cta = lane; cta < ncta; cta += 32) s += __ldcg(part + static_cast<size_t>(cta) * ncol + c);
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
  return s;
}
// Stepwise plan: the last CTA folds all columns in that order, four columns per warp at a time for load parallelism.
__device__ __forceinline__ void reduce_partials_wide(const double* part, int ncta, int ncol, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = warp * 4; c < ncol; c += nwarp * 4) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int cta = lane; cta < ncta; cta += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c + i < ncol) s[i] += __ldcg(part + static_cast<size_t>(cta) * ncol + c + i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int off = 16; off; off >>= 1) s[i] += __shfl_xor_sync(kFull, s[i], off);
      if (lane == 0 && c + i < ncol) out[c + i] = s[i];
    }
  }
  __syncthreads();
}

// Two-level variant for wide models (P+1 > kWideCols), used with one rank as well: letting every CTA read every
// CTA's partials costs ncta^2 * (P+1) float64 loads from L2 per leapfrog step (175 MB at D=1000 on 148 SMs).
// Instead the columns are cut into kWideSlices slices; after the grid barrier the CTA that owns a slice sums
// that slice over all CTAs' partials (one warp per column, fixed order), stores the slice totals into entry
// `rank` of every rank's inbox and bumps that inbox's slice counter (release, system scope; remote atomics over
// NVLink for peers). A CTA continues once its own inbox has all slices of all ranks, and adds the entries in
// rank order. Traffic per step: ncta * (P+1) loads for the slices + nranks * (P+1) per CTA for the totals.
struct WideArgs {
  const double* part;  // [ncta][ncol] partials of this pass
  unsigned char* const* inbox;
  int* abort_flag;
  long long spin_limit;
  int nranks, rank, ncol, ncta;
};
static __device__ __noinline__ bool wide_allreduce(const WideArgs a, unsigned long long seq, double* cta_acc, int* s_flag) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int R = a.nranks;
  const int slot = static_cast<int>(seq & 1ull);
  const int w = (a.ncol + kWideSlices - 1) / kWideSlices;
  const int n_slices = (a.ncol + w - 1) / w;
  const size_t entry = (static_cast<size_t>(slot) * kMaxRanks + a.rank) * kInboxStride;
  int mine = 0;
  for (int sl = blockIdx.x; sl < n_slices; sl += a.ncta) {
    const int c0 = sl * w, c1 = min(a.ncol, c0 + w);
    for (int c = c0 + warp; c < c1; c += nwarp) {
      const double s = wide_column_sum(a.part, a.ncta, a.ncol, c, lane);
      if (lane < R) reinterpret_cast<double*>(a.inbox[lane] + kInboxDataOff)[entry + c] = s;
    }
    ++mine;
  }
  __syncthreads();
  if (mine && tid < R) {
    __threadfence_system();  // cumulative over the CTA's stores (bar.sync above)
    red_release_sys_add_u64(reinterpret_cast<unsigned long long*>(a.inbox[tid] + kInboxCountOff) + slot * kMaxRanks + a.rank,
                            static_cast<unsigned long long>(mine));
  }
  const unsigned char* own = a.inbox[a.rank];
  if (tid == 0) *s_flag = 0;
  __syncthreads();
  if (tid < R) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(own + kInboxCountOff) + slot * kMaxRanks + tid;
    const unsigned long long target = static_cast<unsigned long long>(n_slices) * ((seq + static_cast<unsigned long long>(slot)) >> 1);
    const long long t_start = clock64();
    unsigned int n = 0;
    while (ld_acquire_sys_u64(f) < target) {
      if ((++n & 255u) == 0u) {
        if (clock64() - t_start > a.spin_limit) atomicExch(a.abort_flag, 1);
        if (ld_volatile_s32(a.abort_flag)) {
          *s_flag = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
  if (*s_flag) return false;
  const double* data = reinterpret_cast<const double*>(own + kInboxDataOff) + static_cast<size_t>(slot) * kMaxRanks * kInboxStride;
  for (int c = tid; c < a.ncol; c += nthr) {
    double sum = 0.0;
    for (int rr = 0; rr < R; ++rr) sum += __ldcg(data + static_cast<size_t>(rr) * kInboxStride + c);
    cta_acc[c] = sum;
  }
  __syncthreads();
  return true;
}

// Leader protocol (one GPU, P+1 <= kWideCols): the serial section between two data passes runs on CTA 0 alone.
//   worker CTA, end of pass p:   stores its float64 sums as flag-in-data entries {lo32, seq, hi32, seq} (no fence, no
//                                counter) and goes straight to polling the position of pass p+1;
//   group leader (every kLlGroup-th CTA): polls the entries of its group (its own included) until every one carries
//                                seq(p), adds them in ascending order and publishes the group sum the same way (one CTA
//                                gathering all 148 x (P+1) entries alone is bound by its 64 B/clk L2 port: 2,000+ cycles);
//   leader (CTA 0):              polls the group sums, adds them in the order of reduce_partials, runs the integrator /
//                                accept step and stores the next position as {float, seq(p+1)} words, which the workers'
//                                lanes are polling.
// Against "grid barrier + every CTA reads every CTA's partials + redundant integrator" this removes the release fence
// and the counter round trip of the barrier, 147 of the 148 reads of the partials (which contend for the same L2
// lines) and one L2 round trip between "totals known" and "next pass starts". Both buffers are double-buffered on the
// parity of the pass; reuse is safe because entry p+2 of a CTA is only written after it has seen position p+2, which
// the leader computes after it has consumed every entry of pass p+1 (and so of pass p).
// The summation order is reduce_partials' (tree16 over the CTAs of a group, tree16 over the groups), so the persistent
// and the stepwise plan stay bit-identical.
// Level 1 (group leaders, CTA index % kLlGroup == 0): thread c tree-sums column c of its group's members (tree16, the
// order of reduce_partials) and publishes the group sum as a flag-in-data entry.
static __device__ __forceinline__ void ll_group_sum(const uint4* part, uint4* grp_out, int ncta, int ncol, unsigned int seq) {
  const int g0 = blockIdx.x;  // first member
  const int nm = min(kLlGroup, ncta - g0);
  for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
    uint4 e[kLlGroup];
#pragma unroll
    for (int m = 0; m < kLlGroup; ++m)
      e[m] = m < nm ? ld_relaxed_v4(part + static_cast<size_t>(g0 + m) * ncol + c) : make_uint4(0u, seq, 0u, seq);
    bool again;
    do {
      again = false;
#pragma unroll
      for (int m = 0; m < kLlGroup; ++m) {
        if (e[m].y != seq || e[m].w != seq) {
          e[m] = ld_relaxed_v4(part + static_cast<size_t>(g0 + m) * ncol + c);
          again = true;
        }
      }
    } while (again);
    double v[16];
#pragma unroll
    for (int m = 0; m < kLlGroup; ++m) v[m] = __hiloint2double(static_cast<int>(e[m].z), static_cast<int>(e[m].x));
    const double s = tree16(v);
    st_relaxed_v4(grp_out + c, make_uint4(static_cast<unsigned int>(__double2loint(s)), seq,
                                          static_cast<unsigned int>(__double2hiint(s)), seq));
  }
}

// Level 2 (CTA 0): thread c polls the group sums of column c (all loads in flight together), tree-sums them and owns
// cta_acc[c] — the thread that goes on to compute latent c's gradient, so only the log-likelihood column needs a barrier.
static __device__ __forceinline__ void ll_reduce_groups(const uint4* grp, int ngroups, int P, unsigned int seq, double* cta_acc,
                                                     long long* dbg) {
  const int ncol = P + 1;
  for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
    if (dbg && threadIdx.x == 0) dbg[20] = clock64();
    uint4 e[kLlMaxGroups];
#pragma unroll
    for (int g = 0; g < kLlMaxGroups; ++g)
      e[g] = g < ngroups ? ld_relaxed_v4(grp + static_cast<size_t>(g) * ncol + c) : make_uint4(0u, seq, 0u, seq);
    bool again;
    do {
      again = false;
#pragma unroll
      for (int g = 0; g < kLlMaxGroups; ++g) {
        if (e[g].y != seq || e[g].w != seq) {
          e[g] = ld_relaxed_v4(grp + static_cast<size_t>(g) * ncol + c);
          again = true;
        }
      }
    } while (again);
    if (dbg && threadIdx.x == 0) dbg[21] = clock64();
    double v[16];
#pragma unroll
    for (int g = 0; g < kLlMaxGroups; ++g) v[g] = __hiloint2double(static_cast<int>(e[g].z), static_cast<int>(e[g].x));
    cta_acc[c] = tree16(v);
  }
  __syncthreads();
}

// Development timeline (edhmc_set_timeline): thread 0 of every CTA stamps clock64 / globaltimer at fixed points of
// the first tl_cap passes of a persistent launch. Record = kTlRec int64: {pass start, tiles done + CTA reduced, partials
// published, grid barrier passed, totals ready (incl. peer exchange), integrator done, globaltimer at pass start,
// globaltimer at barrier passed, cycles warps 0..7 spent waiting for their tiles to land}.
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define EDHMC_TL(slot, val)                                                                                   \
  do {                                                                                                        \
    if (tl_on && pass < a.tl_cap) a.timeline[(static_cast<size_t>(pass) * gridDim.x + blockIdx.x) * kTlRec + (slot)] = (val); \
  } while (0)

template <int G, int V, int K, int NW, int RM = 0>
__global__ void __launch_bounds__(NW * 32, 1) k_hmc(const KArgs a) {
  constexpr int kThreads = NW * 32;
  constexpr int CT = kChainThreads;  // O(P) chain work is owned by the first 256 threads: same reduction order for every NW
  const bool single = (a.mode == 1);
  if (single && a.gate && !a.sc->need_init) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_last;
  __shared__ uint32_t s_tmem;  // ring mode 2: base address of the tensor memory allocation
  __shared__ int s_init0;  // leader protocol: 1 if this launch starts with the initial evaluation pass
  // The chain scalars live in shared memory BETWEEN the serial sections: every thread reloads them after a data pass
  // and thread 0 stores them back before the next one, so none of them occupies a register across the tile loop (with
  // them live the allocator rematerialised address arithmetic inside the loop: 262 instructions per 32 rows, ncu r02).
  // Two copies, selected by the parity of the pass: thread 0 may already store the state for pass p+1 while a slow
  // warp still reads the state of pass p (a mid-trajectory serial section has no barrier between the two).
  __shared__ ChainRegs s_cs2[2];
  const int ngroups = RM == 2 ? 1 : (RM == 1 ? NW / a.wpg : NW);  // ring mode 1: warp groups, each with a row range and a ring
  const SmemLayout sm = carve_smem(smem_raw, a, ngroups);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, D = a.D;
  const int ncta = gridDim.x;
  const int ppad = static_cast<int>(align_up(static_cast<size_t>(P) * 4, 16) / 4);
  float* z = sm.state;
  float* r = z + ppad;
  float* g = r + ppad;
  float* zc = g + ppad;
  float* gc = zc + ppad;

  uint32_t tm_base = kNoTmem;
  if constexpr (RM == 2) {
    for (int i = tid; i < a.wpad; i += kThreads) sm.theta_s[i] = 0.0f;  // no ring, no barriers: theta only
    const bool tm_use = a.tm_on != 0 && a.mode == 0;
    if (tm_use) {  // tensor memory for resident tiles (stream_ldg.cuh); released after the pass loop on every exit path
      if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
      tc_fence_before();
    }
    __syncthreads();
    if (tm_use) {
      tc_fence_after();
      tm_base = s_tmem;
    }
  } else if constexpr (RM == 1) {
    smem_setup_cta(sm, a, ngroups);
  } else {
    smem_setup(sm, a, NW);
  }
  const PlanRegs pr = plan_regs(a);
  const uint64_t policy = a.l2_hint == 2 ? l2_policy_pin_fraction(a.l2_frac) : (a.l2_hint ? l2_policy_evict_last() : 0ull);
  // ring mode 0: every warp owns a row range and a private ring; ring mode 1: the CTA owns a row range and one ring
  const int group = RM == 1 ? warp / a.wpg : warp;
  const WarpTiles wt = RM == 2 ? WarpTiles{} : warp_tiles(a, blockIdx.x * ngroups + group, ncta * ngroups);
  using RingT = typename std::conditional<RM >= 1, CtaRing, Ring>::type;
  RingT ring;
  if constexpr (RM == 2)
    ring.cpass = 0;  // ring mode 2 has no ring: only the pass parity (zig-zag direction) lives here
  else if constexpr (RM == 1)
    cta_ring_init(ring, sm, a, group);
  else
    ring_init(ring, sm.ring + warp * a.S * a.stage_floats, sm.bars + warp * kMaxStages);
  const bool tl_on = a.timeline != nullptr && tid == 0 && !single;

  // prior of this thread's first latent, in registers: keeps two global loads and two float64 divisions off the
  // serial section between two data passes
  float pl0 = 0.0f, ps0 = 1.0f;
  double piv0 = 1.0;
  int pk0 = 0;
  if (tid < P && tid < CT) {
    pl0 = a.prior_loc[tid];
    ps0 = a.prior_scale[tid];
    piv0 = prior_inv_var(ps0);
    pk0 = a.prior_kind ? a.prior_kind[tid] : 0;
  }

  // ---- chain registers (uniform across threads and CTAs) ----
  double logp_cur = 0.0, logp_new = 0.0, k_old = 0.0, log_u = 0.0;
  long long it = 0, n_acc = 0;
  int s = 0, nonfinite = 0;
  bool in_init = false;
  long long n_passes = 1;

  // Metropolis–Hastings accept + Empirical write for transition `it` (hmc.py:100-126).
  auto finish_transition = [&]() {
    double ks = 0.0;
    for (int c = tid; c < P && tid < CT; c += CT) ks += static_cast<double>(__fmul_rn(r[c], r[c]));
    const double k_new = 0.5 * block_sum_f64(ks, sm.red);
    const AcceptResult ar = mh_accept(k_old, k_new, logp_new, logp_cur, log_u);
    if (!isfinite(logp_new)) nonfinite = 1;
    const long long t = a.t0 + it;
    if (blockIdx.x == 0) {
      if (tid == 0) write_trace(a, it, logp_cur, logp_new, k_old, k_new, ar);
      if (a.trace_pos)
        for (int c = tid; c < P && tid < CT; c += CT) a.trace_pos[it * P + c] = z[c];
    }
    if (ar.accept) {
      for (int c = tid; c < P && tid < CT; c += CT) {
        zc[c] = z[c];
        gc[c] = g[c];
      }
      logp_cur = logp_new;
      ++n_acc;
    }
    if (blockIdx.x == 0)
      for (int c = tid; c < P && tid < CT; c += CT) a.params[t * a.ldp + c] = zc[c];
  };

  // Starts transition `it`: momentum draw, kinetic energy, first half kick + drift, theta for the pass.
  // Transitions that need no data pass (n_steps == 0) are completed on the spot.
  auto start_next = [&]() {
    while (it < a.n_iter) {
      const long long t = a.t0 + it;
      double ks = 0.0;
      for (int c = tid; c < P && tid < CT; c += CT) {
        const float rv = a.r0 ? a.r0[it * P + c] : philox_normal(a.seed, t, c);  // hmc.py:88-91
        r[c] = rv;
        z[c] = zc[c];
        g[c] = gc[c];
        ks += static_cast<double>(__fmul_rn(rv, rv));
      }
      k_old = 0.5 * block_sum_f64(ks, sm.red);
      log_u = static_cast<double>(logf(a.u ? a.u[it] : philox_uniform(a.seed, t)));  // hmc.py:108-109
      s = 0;
      logp_new = logp_cur;
      if (a.L > 0) {
        for (int c = tid; c < P && tid < CT; c += CT) {
          const float rn = kick(r[c], a.half_eps, g[c]);
          r[c] = rn;
          const float zn = drift(z[c], a.eps, rn);
          z[c] = zn;
          if (c < D) sm.theta_s[c] = zn;
        }
        return;
      }
      finish_transition();
      ++it;
    }
  };

  if (single) {
    for (int c = tid; c < D && tid < CT; c += CT) sm.theta_s[c] = a.theta_in[c];
  } else {
    // current state: row max(t0-1,0) of the Empirical store (hmc.py:81-85) + cached (grad, logp)
    const long long t_prev = a.t0 > 0 ? a.t0 - 1 : 0;
    bool mismatch = false;
    for (int c = tid; c < P && tid < CT; c += CT) {
      const float v = a.params[t_prev * a.ldp + c];
      zc[c] = v;
      gc[c] = a.gcur[c];
      if (__float_as_uint(v) != __float_as_uint(a.zcur[c])) mismatch = true;
    }
    const int valid = a.sc->valid;
    logp_cur = a.sc->logp_cur;
    in_init = __syncthreads_or((mismatch || !valid) ? 1 : 0) != 0;
    n_passes = (in_init ? 1 : 0) + a.n_iter * a.L;
  }
  // Zig-zag direction = parity of the pass's global leapfrog-step index t*L + s (the initial evaluation
  // counts as step -1), so the accumulation order does not depend on how transitions are chunked into
  // launches.
  {
    const int par0 = single ? (a.par0 & 1) : static_cast<int>((a.t0 * a.L - (in_init ? 1 : 0)) & 1);
    ring.cpass = par0;
    if constexpr (RM == 2) ldg_load_resident<K, NW>(a, sm, tm_base);
    if constexpr (RM == 1)
      ring.par0 = par0;
    else if constexpr (RM == 0)
      ring.ipass = par0;
  }
  if constexpr (RM == 1)
    cta_ring_prologue(pr, wt, ring, n_passes, policy, tid == group * a.wpg * 32);
  else if constexpr (RM == 0)
    ring_prologue(pr, wt, ring, n_passes, lane, policy);
  if (!single) {
    if (in_init) {
      for (int c = tid; c < D && tid < CT; c += CT) sm.theta_s[c] = zc[c];
    } else {
      start_next();
    }
  }
  auto store_chain = [&](long long for_pass) {
    if (tid == 0) {
      ChainRegs& cs = s_cs2[for_pass & 1];
      cs.logp_cur = logp_cur;
      cs.logp_new = logp_new;
      cs.k_old = k_old;
      cs.log_u = log_u;
      cs.it = it;
      cs.n_acc = n_acc;
      cs.s = s;
      cs.nonfinite = nonfinite;
      cs.in_init = in_init ? 1 : 0;
    }
  };
  auto load_chain = [&](long long of_pass) {
    const ChainRegs& cs = s_cs2[of_pass & 1];
    logp_cur = cs.logp_cur;
    logp_new = cs.logp_new;
    k_old = cs.k_old;
    log_u = cs.log_u;
    it = cs.it;
    n_acc = cs.n_acc;
    s = cs.s;
    nonfinite = cs.nonfinite;
    in_init = cs.in_init != 0;
  };
  store_chain(0);
  __syncthreads();

  int bufsel = 0;
  unsigned long long epoch = 0;
  constexpr bool kMayWide = G * K * V + 2 > kWideCols;  // P + 1 <= G*K*V + 2: narrow kernels carry no wide path
  // (with row shards the choice must not depend on this rank's grid size: every rank has to speak the same protocol)
  const bool wide = kMayWide && !single && P + 1 > kWideCols && (ncta > 1 || a.nranks > 1);
  const bool guarded = a.nranks > 1 || wide;  // waits that can time out: poll the abort flag
  const unsigned long long seq0 = (!single && guarded) ? *a.comm_seq : 0ull;
  bool aborted = false;
  // leader protocol (see ll_reduce_partials): the host selects it for one GPU and P+1 <= kWideCols
  // (a.leader is set only for the persistent plan on one GPU with a narrow model and more than one CTA). Nothing of it
  // is kept in registers across the data pass: the flag comes from the constant bank, in_init0 from shared memory.
  // a.leader == 2 ("flat"): the partials and group sums travel the same way, but EVERY CTA polls the group sums and runs
  // the O(P) integrator itself on its own copy of the chain state (as the grid-barrier protocol does), so the next position
  // needs no broadcast: one L2 round trip (leader -> workers) less per leapfrog step. Only CTA 0 writes global state.
#define EDHMC_LEAD (a.leader != 0)
#define EDHMC_WORKER (a.leader == 1 && blockIdx.x != 0)
  if (tid == 0) s_init0 = in_init ? 1 : 0;
  __syncthreads();
  for (long long pass = 0; pass < n_passes; ++pass) {
    if (EDHMC_WORKER && pass > 0) {
      const unsigned int ll_seq = a.ll_seq0 + static_cast<unsigned int>(pass) + 1u;
      // the position of this pass, from the leader: every lane polls its own {float, seq} word
      const uint2* src = a.ll_theta + (static_cast<size_t>(pass & 1) * kLlCopies + (blockIdx.x % kLlCopies)) * P;
      for (int c = tid; c < P; c += kThreads) {
        uint2 v;
        do {
          v = ld_relaxed_v2(src + c);
        } while (v.y != ll_seq);
        const float zn = __uint_as_float(v.x);
        z[c] = zn;
        if (c < D) sm.theta_s[c] = zn;
      }
      EDHMC_TL(25, clock64());
      __syncthreads();
    }
    // which pass of a trajectory this is follows from its index alone (every transition has exactly L passes)
    const bool in_init_p = EDHMC_LEAD ? (s_init0 != 0 && pass == 0) : (s_cs2[pass & 1].in_init != 0);
    const int s_p = EDHMC_LEAD ? (static_cast<int>(pass) - s_init0) % (a.L > 0 ? a.L : 1) : s_cs2[pass & 1].s;
    const float* pos = single ? a.theta_in : (in_init_p ? zc : z);
    const float bias = a.has_bias ? pos[D] : 0.0f;
    // the log likelihood is only consumed at the ends of a trajectory (initial evaluation, last leapfrog step)
    const bool want_lp = single ? (a.single_lp != 0) : (in_init_p || s_p + 1 >= a.L);
    EDHMC_TL(0, clock64());
    EDHMC_TL(6, global_timer_ns());
    if constexpr (RM == 2) {
      stream_pass_ldg<K, NW>(a, sm, bias, want_lp, a.zigzag && (ring.cpass & 1), tm_base);
      ++ring.cpass;
    } else if constexpr (RM == 1) {
      stream_pass_cta<G, V, K, NW>(a, pr, wt, ring, sm, bias, policy, want_lp);
    } else {
      stream_pass<G, V, K, NW>(a, pr, wt, ring, sm, bias, policy, want_lp);
    }
    EDHMC_TL(1, clock64());

    if (single) {
      // last-arriving CTA folds the partials (threadFenceReduction pattern), fixed summation order
      double* mine = a.partials + static_cast<size_t>(blockIdx.x) * (P + 1);
      for (int c = tid; c <= P; c += kThreads) mine[c] = sm.cta_acc[c];
      __syncthreads();
      if (tid == 0) {
        __threadfence();  // cumulative: orders the whole CTA's partial writes (bar.sync above) before the ticket
        const unsigned int ticket = atomicAdd(a.ticket, 1u);
        s_last = (ticket == static_cast<unsigned int>(ncta - 1));
        __threadfence();
      }
      __syncthreads();
      if (s_last) {
        if (kMayWide && P + 1 > kWideCols)
          reduce_partials_wide(a.partials, ncta, P + 1, sm.cta_acc);
        else
          reduce_partials<2>(a.partials, ncta, P, sm.cta_acc, reinterpret_cast<double*>(sm.xw));
        for (int c = tid; c <= P; c += kThreads) a.sums[c] = sm.cta_acc[c];
        if (tid == 0) *a.ticket = 0u;
      }
      return;
    }

    if (EDHMC_LEAD) {
      const unsigned int ll_seq = a.ll_seq0 + static_cast<unsigned int>(pass) + 1u;
      uint4* mine = a.ll_part + (static_cast<size_t>(pass & 1) * ncta + blockIdx.x) * (P + 1);
      for (int c = tid; c <= P; c += kThreads) {
        const double v = sm.cta_acc[c];
        st_relaxed_v4(mine + c, make_uint4(static_cast<unsigned int>(__double2loint(v)), ll_seq,
                                           static_cast<unsigned int>(__double2hiint(v)), ll_seq));
      }
      EDHMC_TL(2, clock64());
      const int ngroups = (ncta + kLlGroup - 1) / kLlGroup;
      uint4* grp = a.ll_group + static_cast<size_t>(pass & 1) * ngroups * (P + 1);
      if (blockIdx.x % kLlGroup == 0)
        ll_group_sum(a.ll_part + static_cast<size_t>(pass & 1) * ncta * (P + 1), grp + static_cast<size_t>(blockIdx.x / kLlGroup) * (P + 1),
                     ncta, P + 1, ll_seq);
      if (EDHMC_WORKER) continue;  // next: poll the position of pass + 1 (top of the loop)
      ll_reduce_groups(grp, ngroups, P, ll_seq, sm.cta_acc,
                       (tl_on && pass < a.tl_cap) ? a.timeline + (static_cast<size_t>(pass) * gridDim.x + blockIdx.x) * kTlRec
                                                   : nullptr);
      EDHMC_TL(3, clock64());
      EDHMC_TL(7, global_timer_ns());
    } else if (ncta > 1 || wide) {
      double* mine = a.partials + (static_cast<size_t>(bufsel) * ncta + blockIdx.x) * (P + 1);
      for (int c = tid; c <= P; c += kThreads) mine[c] = sm.cta_acc[c];
      __syncthreads();
      EDHMC_TL(2, clock64());
      if (tid == 0) {
        // release (cumulative over the CTA's writes ordered by the bar.sync above) / acquire on one counter
        red_release_add_u64(a.bar, 1ull);
        const unsigned long long target = (epoch + 1) * static_cast<unsigned long long>(ncta);
        s_last = 0;
        if (guarded) {
          // a CTA that gave up on a peer never arrives here again: poll the abort flag while waiting
          unsigned int n = 0;
          while (ld_acquire_u64(a.bar) < target) {
            if ((++n & 1023u) == 0u && ld_volatile_s32(a.abort_flag)) {
              s_last = 1;
              break;
            }
          }
        } else {
          while (ld_acquire_u64(a.bar) < target) {
          }
        }
      }
      __syncthreads();
      if (s_last) {
        aborted = true;
        break;
      }
      EDHMC_TL(3, clock64());
      EDHMC_TL(7, global_timer_ns());
      if (!wide) reduce_partials<2>(a.partials + static_cast<size_t>(bufsel) * ncta * (P + 1), ncta, P, sm.cta_acc, reinterpret_cast<double*>(sm.xw));
      bufsel ^= 1;
      ++epoch;
    }
    if (wide) {
      WideArgs wa;
      wa.part = a.partials + static_cast<size_t>(bufsel ^ 1) * ncta * (P + 1);
      wa.inbox = a.peer_inbox;
      wa.abort_flag = a.abort_flag;
      wa.spin_limit = a.spin_limit;
      wa.nranks = a.nranks;
      wa.rank = a.nranks > 1 ? a.rank : 0;
      wa.ncol = P + 1;
      wa.ncta = ncta;
      if (!wide_allreduce(wa, seq0 + static_cast<unsigned long long>(pass) + 1ull, sm.cta_acc, &s_last)) {
        aborted = true;
        break;
      }
    } else if (a.nranks > 1) {
      PeerArgs pa;
      pa.inbox = a.peer_inbox;
      pa.abort_flag = a.abort_flag;
      pa.spin_limit = a.spin_limit;
      pa.nranks = a.nranks;
      pa.rank = a.rank;
      pa.ncol = P + 1;
      if (!peer_allreduce(pa, seq0 + static_cast<unsigned long long>(pass) + 1ull, sm.cta_acc, &s_last)) {
        aborted = true;
        break;
      }
    }

    EDHMC_TL(4, clock64());
    load_chain(pass);
    // gradient and log joint at `pos`: likelihood totals + Normal prior (hmc.py:183-190)
    float* gout = in_init ? gc : g;
    double pl = 0.0;
    for (int c = tid; c < P && tid < CT; c += CT) {
      float loc = pl0, sc = ps0;
      double iv = piv0;
      int kind = pk0;
      if (c != tid) {  // models with more than kChainThreads latents: the further columns come from global memory
        loc = a.prior_loc[c];
        sc = a.prior_scale[c];
        iv = prior_inv_var(sc);
        kind = a.prior_kind ? a.prior_kind[c] : 0;
      }
      gout[c] = static_cast<float>(sm.cta_acc[c] + prior_grad_kind(kind, pos[c], loc, sc, iv));
      if (want_lp) pl += prior_logp_kind(kind, pos[c], loc, sc);
    }
    if (want_lp) {
      const double lik = sm.cta_acc[P];
      logp_new = (block_sum_f64(pl, sm.red) - a.prior_const) + lik;
    }
    EDHMC_TL(22, clock64());

    if (in_init) {
      logp_cur = logp_new;
      in_init = false;
      start_next();
    } else {
      ++s;
      if (s == a.L) {
        for (int c = tid; c < P && tid < CT; c += CT) r[c] = kick(r[c], a.half_eps, g[c]);
        finish_transition();
        ++it;
        start_next();
      } else {
        // end of step s (hmc.py:207-208) and start of step s+1 (hmc.py:201-204): two separate half kicks
        for (int c = tid; c < P && tid < CT; c += CT) {
          const float rn = kick(kick(r[c], a.half_eps, g[c]), a.half_eps, g[c]);
          r[c] = rn;
          const float zn = drift(z[c], a.eps, rn);
          z[c] = zn;
          if (c < D) sm.theta_s[c] = zn;
        }
      }
    }
    EDHMC_TL(23, clock64());
    if (a.leader == 1 && pass + 1 < n_passes) {
      // the position of pass + 1 to the workers; each thread sends the latents it has just written itself
      const unsigned int seq_next = a.ll_seq0 + static_cast<unsigned int>(pass) + 2u;
      uint2* dst = a.ll_theta + static_cast<size_t>((pass + 1) & 1) * kLlCopies * P;
      for (int c = tid; c < P && tid < CT; c += CT) {
        const uint2 w = make_uint2(__float_as_uint(z[c]), seq_next);
#pragma unroll
        for (int k = 0; k < kLlCopies; ++k) st_relaxed_v2(dst + static_cast<size_t>(k) * P + c, w);
      }
    }
    EDHMC_TL(24, clock64());
    store_chain(pass + 1);
    __syncthreads();
    EDHMC_TL(5, clock64());
  }

  if constexpr (RM == 2) {
    if (tm_base != kNoTmem) {  // uniform over the CTA; also reached when a peer wait timed out
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tm_base, 512);
      }
    }
  }
  if (aborted) return;
  load_chain(n_passes);
  if (!single && blockIdx.x == 0) {
    if (guarded && tid == 0) *a.comm_seq = seq0 + static_cast<unsigned long long>(n_passes);
    for (int c = tid; c < P && tid < CT; c += CT) {
      a.zcur[c] = zc[c];
      a.gcur[c] = gc[c];
    }
    if (tid == 0) {
      a.sc->logp_cur = logp_cur;
      a.sc->valid = 1;
      a.sc->n_accept += n_acc;
      if (nonfinite) a.sc->nonfinite = 1;
    }
  }
}

}  // namespace edhmc
