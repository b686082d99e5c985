// edhmc.cu — host side of libedhmc.so: execution planner, kernel dispatch, the C ABI of include/edhmc.h,
// and NCCL (loaded lazily with dlopen so the library has no link-time NCCL dependency).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/edhmc.h"
#include "chain_small.cuh"
#include "chains.cuh"
#include "chains_wide.cuh"
#include "f64.cuh"
#include "criticism.cuh"

using namespace edhmc;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(EDHMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ------------------------------------------------------------------------------------------------
struct NcclUid {
  char internal[128];
};
typedef int (*fn_ncclGetUniqueId)(NcclUid*);
typedef int (*fn_ncclCommInitRank)(void**, int, NcclUid, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void*);
typedef const char* (*fn_ncclGetErrorString)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_ncclGetUniqueId GetUniqueId = nullptr;
  fn_ncclCommInitRank CommInitRank = nullptr;
  fn_ncclAllReduce AllReduce = nullptr;
  fn_ncclCommDestroy CommDestroy = nullptr;
  fn_ncclGetErrorString GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // reuse the copy torch already mapped
    if (lib) break;
  }
  if (!lib)
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
  if (!lib) return fail(EDHMC_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(EDHMC_ERR_COMM, "libnccl is missing required symbols");
  g_nccl.lib = lib;
  return 0;
}
#define NCCL_TRY(expr)                                                                                   \
  do {                                                                                                   \
    int r__ = (expr);                                                                                    \
    if (r__ != 0)                                                                                        \
      return fail(EDHMC_ERR_COMM, "%s failed: %s", #expr,                                                \
                  g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");                    \
  } while (0)
static const int kNcclFloat64 = 8, kNcclSum = 0;

// ------------------------------------------------------------------------------------------------
// Process-wide communicator state, one per device: the NCCL communicator and the peer-mapped inboxes of the in-kernel
// all-reduce are created by the FIRST handle of a (device, nranks, rank) and reused by every later one. ed.HMC builds a
// handle per inference object; without the cache each construction paid ncclCommInitRank + cudaIpc export / attach +
// three all-gathers (2.4 s of a 3.2 s call at 8 GPUs, round-1 verdict). The exchange sequence number lives with the
// inboxes, so handles that share them continue one sequence; sharded runs of one process must not overlap in time.
// ------------------------------------------------------------------------------------------------
struct SharedComm {
  void* comm = nullptr;
  int nranks = 0, rank = -1;
  unsigned char* inbox = nullptr;            // this rank's inbox (own cudaMalloc: cudaIpc exports whole allocations)
  unsigned char** d_peer_ptrs = nullptr;     // device table [kMaxRanks]
  unsigned long long* d_seq = nullptr;       // passes exchanged so far (identical on every rank)
  void* mapped[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool peers_ready = false;
};
static SharedComm g_shared[64];

static void shared_release(SharedComm& sc) {
  if (sc.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(sc.comm);
  for (int r = 0; r < kMaxRanks; ++r)
    if (sc.mapped[r]) cudaIpcCloseMemHandle(sc.mapped[r]);
  if (sc.inbox) cudaFree(sc.inbox);
  if (sc.d_peer_ptrs) cudaFree(sc.d_peer_ptrs);
  if (sc.d_seq) cudaFree(sc.d_seq);
  sc = SharedComm();
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct Plan {
  int G = 0, V = 0, K = 0, Kact = 0, NW = 0, J = 0, RT = 0, S = 0, stage_floats = 0, y_off = 0, wpad = 0, tl = 0, tm = 0;
  int grid = 0;
  int WPG = 0;  // ring mode 1: warps per group
  int RM = 0;  // ring mode: 0 = one ring per warp (stream.cuh), 1 = one ring per CTA (stream_cta.cuh, narrow rows)
  int n_res = 0;  // ring mode 2: 32-row tiles per CTA kept resident in shared memory during a persistent launch
  int n_tm = 0;   // ring mode 2: 32-row tiles per CTA parked in tensor memory during a persistent launch
  size_t smem = 0;
  const void* fn = nullptr;
};

#include "inst_table.inc"

struct edhmc_handle {
  edhmc_cfg cfg;
  int P = 0;
  int num_sms = 0;
  int smem_optin = 0;
  bool bound = false;
  const float* X = nullptr;
  const void* y = nullptr;
  int y_dtype = 0;
  int* y_owned = nullptr;  // int32 copy when the caller's y is uint8
  Plan plan;        // the plan of full-data passes (HMC, logp_grad, full-batch SGLD / SGHMC)
  Plan plan_rows;   // a plan that reads the caller's row-major X (ring mode 0 / 1): row-window passes (mini-batches)
  float2* d_xt = nullptr;  // ring mode 2: X re-laid into column-pair-major 32-row tiles (stream_ldg.cuh), owned
  float* d_yt = nullptr;
  long long n_tiles = 0;
  int zigzag = 1;
  int l2_hint = 0;
  float l2_frac = 1.0f;
  int interleave = 0;
  // device buffers
  float* d_prior_loc = nullptr;
  float* d_prior_scale = nullptr;
  int* d_prior_kind = nullptr;  // nullptr: every latent has a Normal prior
  std::vector<float> prior_loc_h, prior_scale_h;
  double prior_const = 0.0;
  ChainScalars* d_sc = nullptr;
  float* d_zcur = nullptr;
  float* d_gcur = nullptr;
  float* d_z = nullptr;
  float* d_r = nullptr;
  float* d_g = nullptr;
  double* d_partials = nullptr;
  uint4* d_ll_part = nullptr;   // leader protocol (chain.cuh): flag-in-data partials [2][num_sms][P+1]
  uint4* d_ll_group = nullptr;  // leader protocol: sums of groups of kLlGroup CTAs [2][ceil(num_sms / kLlGroup)][P+1]
  uint2* d_ll_theta = nullptr;  // leader protocol: flag-in-data position [2][P]
  unsigned int ll_seq = 1;      // sequence numbers handed out so far
  int leader = 2;               // 2: flag-in-data partials + group sums, every CTA integrates (default); 1: CTA 0 integrates and
                                // broadcasts the position; 0: grid barrier, every CTA reads every partial (EDHMC_LEADER, same-box A/B)
  size_t partials_cap = 0;
  unsigned long long* d_bar = nullptr;
  unsigned int* d_ticket = nullptr;
  double* d_sums = nullptr;
  unsigned long long* d_bad = nullptr;
  uint64_t seed = 0x243F6A8885A308D3ull;
  double* trace_scalars = nullptr;
  float* trace_pos = nullptr;
  // nccl
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  // peer inboxes of the in-kernel all-reduce (persistent plan over row shards)
  unsigned char* d_inbox = nullptr;
  unsigned char** d_peer_ptrs = nullptr;
  void* peer_mapped[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long* d_comm_seq = nullptr;
  unsigned char** own_peer_ptrs = nullptr;    // the handle's own table / counter (single-GPU wide reduction); d_peer_ptrs and
  unsigned long long* own_comm_seq = nullptr;  // d_comm_seq point at g_shared[device] once peers are attached
  int* d_abort = nullptr;
  bool peers_ready = false;
  bool shared_comm = false;  // comm / peer tables belong to g_shared[device]
  long long spin_limit = 0;
  // many chains
  int C = 0;     // the caller's chains
  int Cpad = 0;  // rounded up to whole 128-chain tiles: what the pass kernels run and the handle's state arrays hold
  float* mc_theta_pad = nullptr;  // [Cpad][P] staging of a caller's theta when C is not a multiple of 128
  int mc_nrg = 0, mc_use_tc = 0, mc_Dp = 0;
  float *mc_z = nullptr, *mc_r = nullptr, *mc_g = nullptr, *mc_zcur = nullptr, *mc_gcur = nullptr;
  double *mc_part_g = nullptr, *mc_logp = nullptr, *mc_kold = nullptr, *mc_logu = nullptr, *mc_part_lp = nullptr;
  long long* mc_nacc = nullptr;
  int* mc_flags = nullptr;  // [0] valid, [1] need_init
  double* mc_trace = nullptr;
  long long* mc_dbg = nullptr;
  // wide models / row shards: two-GEMM path (chains_wide.cu)
  bool mc_wide = false, mcw_pretiled = false;
  McwArgs mcw;  // tiling + buffers; the run fields are filled per call
  float* mc_xt = nullptr;  // pre-tiled operand copy of X (tensor-core pass v3)
  float* mc_yt = nullptr;
  bool mc_pretiled = false;
  // one slab for the small per-handle buffers: a handle costs 2 cudaMalloc / cudaFree instead of ~20 (each cudaFree is
  // a device synchronisation; ed.HMC builds and drops a handle per inference object)
  unsigned char* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  // stats
  long long passes_last = 0, launches_last = 0;
  int plan_in_use = 0;
  // float64 models (f64.cuh)
  bool f64 = false;
  const double* X64 = nullptr;
  double *d64_zcur = nullptr, *d64_gcur = nullptr, *d64_z = nullptr, *d64_r = nullptr, *d64_g = nullptr, *d64_partials = nullptr;
  double *trace_scalars64 = nullptr, *trace_pos64 = nullptr;
  bool no_cta_ring = false;  // y is not 16-byte aligned: the CTA-wide ring (bulk copies of y) cannot be used
  long long* timeline = nullptr;
  int tl_cap = 0;
};

// Shared-memory bank-conflict degree of the row loads for G lanes per row, vectors of V floats, row stride
// ldx floats: lanes of one LDS phase (32/V lanes) hit bank groups of V words; returns the worst multiplicity.
// warps per CTA of the float64 pass: as many (<= 8) as fit their [P+1] double slices in shared memory
static int pass64_warps(int P) {
  const size_t per_warp = static_cast<size_t>(P + 1) * sizeof(double);
  size_t nw = (200u << 10) / per_warp;
  return nw >= 8 ? 8 : static_cast<int>(nw);
}

static cudaError_t h_alloc(edhmc_handle* h, void** p, size_t bytes) {
  const size_t need = (bytes + 255) / 256 * 256;
  if (h->arena && h->arena_used + need <= h->arena_cap) {
    *p = h->arena + h->arena_used;
    h->arena_used += need;
    return cudaSuccess;
  }
  return cudaMalloc(p, bytes);
}
static void h_free(edhmc_handle* h, void* p) {
  if (!p) return;
  const unsigned char* q = static_cast<const unsigned char*>(p);
  if (h->arena && q >= h->arena && q < h->arena + h->arena_cap) return;  // carved from the slab
  cudaFree(p);
}


// Shared-memory bank-conflict degree of the row loads for G lanes per row, vectors of V floats, row stride
// ldx floats: lanes of one LDS phase (32/V lanes) hit bank groups of V words; returns the worst multiplicity.
static int conflict_degree(long long ldx, int V, int G) {
  const int phase = 32 / V;  // lanes served together by one shared-memory wavefront
  int cnt[32] = {0};
  int worst = 0;
  for (int lane = 0; lane < phase; ++lane) {
    const long long word = (lane / G) * ldx + static_cast<long long>(lane % G) * V;
    const int slot = static_cast<int>((word / V) % phase);
    if (++cnt[slot] > worst) worst = cnt[slot];
  }
  return worst;
}

// Ring mode 1 (stream_cta.cuh) for narrow rows: a row fits one lane at <= 64 floats, ONE lane per row (spreading a row
// over more lanes to raise the warp count measured slower, see gen_inst.py), warps from the register footprint; a
// stage holds WPG * 32 * J rows (~1/3 of the shared memory per ring), one bulk copy of X and one of y. Returns false
// if the shape is not eligible (the caller then plans ring mode 0).
static bool make_plan_cta(edhmc_handle* h, Plan& out) {
  const edhmc_cfg& c = h->cfg;
  Plan p;
  p.RM = 1;
  const int D = c.n_features;
  const long long ldx = c.ldx;
  p.V = (ldx % 4 == 0) ? 4 : (ldx % 2 == 0 ? 2 : 1);
  const int kmax = 64 / p.V;
  const int chunks = (D + p.V - 1) / p.V;
  if (chunks > kmax || ldx > 4096) return false;
  // one lane per row needs a row stride that spreads the lanes of a wavefront over the banks (any odd multiple of the
  // vector width does; 64-, 128-, 256-byte strides do not): those shapes keep ring mode 0, which spreads a row over
  // more lanes instead
  if (conflict_degree(ldx, p.V, 1) >= 2 && !getenv("EDHMC_FORCE_G")) return false;
  static const int tiers4[] = {1, 2, 3, 4, 5, 6, 7, 8, 12, 16}, tiers2[] = {1, 2, 4, 6, 8, 10, 12, 14, 16, 24, 27, 32},
                   tiers1[] = {1, 2, 4, 8, 16, 28, 32, 64};
  const int* tiers = p.V == 4 ? tiers4 : (p.V == 2 ? tiers2 : tiers1);
  const int ntier = p.V == 4 ? 10 : (p.V == 2 ? 12 : 8);
  auto tier_for = [&](int kact) {
    for (int i = 0; i < ntier; ++i)
      if (tiers[i] >= kact) return tiers[i];
    return 0;
  };
  int force_g = 0, force_nw = 0, force_j = 0, force_wpg = 0;
  if (const char* e = getenv("EDHMC_FORCE_G")) force_g = atoi(e);
  if (const char* e = getenv("EDHMC_FORCE_NW")) force_nw = atoi(e);
  if (const char* e = getenv("EDHMC_FORCE_J")) force_j = atoi(e);
  if (const char* e = getenv("EDHMC_FORCE_WPG")) force_wpg = atoi(e);
  const int gs[3] = {1, 2, 4};
  for (int g : gs) {
    if (force_g > 0 && g != force_g) continue;
    const int kact = (chunks + g - 1) / g;
    const int k = tier_for(kact);
    if (!k) continue;
    p.G = g;
    p.Kact = kact;
    p.K = k;
    break;  // one lane per row unless EDHMC_FORCE_G says otherwise
  }
  if (!p.G) return false;
  p.NW = force_nw > 0 ? force_nw : warps_for(p.K * p.V);
  p.fn = lookup_kernel(p.G, p.V, p.K, p.NW, 1);
  if (!p.fn) return false;
  p.wpad = p.G * p.K * p.V;
  const int RPS = 32 / p.G;
  p.WPG = (force_wpg > 0 && p.NW % force_wpg == 0) ? force_wpg : p.NW;
  const int ngrp = p.NW / p.WPG;
  const long long r1 = static_cast<long long>(p.WPG) * RPS;  // rows per stage of a group and J
  if ((r1 & 3) != 0) return false;  // tiles must start on 4-row boundaries
  const size_t budget = static_cast<size_t>(h->smem_optin) - 1024;
  size_t offs[8];
  const size_t fixed = smem_layout_bytes(ngrp, 0, 0, h->P, p.wpad, offs);
  if (fixed + 8192 > budget) return false;
  auto stage_floats_for = [&](long long J, int* y_off) {
    const long long R = r1 * J;
    long long xf = (R - 1) * ldx + p.wpad;
    if (xf < R * ldx) xf = R * ldx;
    xf = (xf + 3) / 4 * 4;
    *y_off = static_cast<int>(xf);
    return static_cast<long long>((xf + R + 31) / 32 * 32);
  };
  long long target = static_cast<long long>((budget - fixed) / 3);
  if (target > 65536) target = 65536;
  target /= ngrp;
  long long J = force_j > 0 ? force_j : (target - 512) / (r1 * (ldx * 4 + 4));
  if (J < 1) J = 1;
  int y_off = 0;
  long long sf = stage_floats_for(J, &y_off);
  int S = kMaxStages;
  for (; S >= 1; --S)
    if (smem_layout_bytes(ngrp, S, static_cast<int>(sf), h->P, p.wpad, offs) <= budget) break;
  if (S < 2) return false;
  p.J = static_cast<int>(J);
  p.RT = static_cast<int>(r1 * J);
  p.tl = static_cast<int>(p.RT * ldx);
  p.tm = 0;
  p.y_off = y_off;
  p.stage_floats = static_cast<int>(sf);
  p.S = S;
  p.smem = smem_layout_bytes(ngrp, S, p.stage_floats, h->P, p.wpad, offs);
  if (cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p.fn, p.NW * 32, p.smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return false;
  }
  long long want = (c.n_rows + 2ll * p.RT * ngrp - 1) / (2ll * p.RT * ngrp);  // >= two tiles per group before another SM is used
  if (want < 1) want = 1;
  p.grid = static_cast<int>(want < h->num_sms ? want : h->num_sms);
  out = p;
  return true;
}

// Chooses the streaming geometry (restated in csrc/gen_inst.py, which instantiates every reachable kernel):
// V = widest vector the row stride allows; G = 1 lane per row while the row fits 64 floats per lane (measured:
// large tiles amortise the per-tile bookkeeping best), else the fewest lanes whose slice is <= 32 floats;
// K = smallest compiled tier >= chunks per lane; warps per CTA from the register footprint (warps_for);
// tiles sized so that >= 3 ring stages per warp fit in shared memory.
static int make_plan_rows(edhmc_handle* h) {
  const edhmc_cfg& c = h->cfg;
  {
    int ring = 1;  // EDHMC_RING=0 keeps the per-warp rings for every shape (A/B runs)
    if (const char* e = getenv("EDHMC_RING")) ring = atoi(e) == 0 ? 0 : 1;
    Plan pc;
    if (ring == 1 && !h->no_cta_ring && !h->interleave && make_plan_cta(h, pc)) {
      h->plan = pc;
      return 0;
    }
  }
  Plan p;
  const int D = c.n_features;
  const long long ldx = c.ldx;
  p.V = (ldx % 4 == 0) ? 4 : (ldx % 2 == 0 ? 2 : 1);
  const int kmax = 64 / p.V, lim = 32 / p.V;
  const int chunks = (D + p.V - 1) / p.V;
  const int gs[5] = {2, 4, 8, 16, 32};
  int force_g = 0, force_nw = 0;
  if (const char* e = getenv("EDHMC_FORCE_G")) force_g = atoi(e);
  if (const char* e = getenv("EDHMC_FORCE_NW")) force_nw = atoi(e);
  p.G = 0;
  if (force_g > 0) {
    if ((chunks + force_g - 1) / force_g <= kmax) p.G = force_g;
  } else if (chunks <= kmax) {
    // one lane per row: the largest tiles, no shuffles, no redundant link-function work — unless the row
    // stride makes those loads collide on the same banks (e.g. 256-byte rows: 8-way); then spread a row
    // over the fewest lanes that bring the conflict degree down to <= 2.
    p.G = 1;
    if (conflict_degree(ldx, p.V, 1) >= 4) {
      const int alt[3] = {2, 4, 8};
      for (int g : alt)
        if (conflict_degree(ldx, p.V, g) <= 2) {
          p.G = g;
          break;
        }
    }
  } else {
    for (int g : gs)
      if ((chunks + g - 1) / g <= lim) {
        p.G = g;
        break;
      }
    if (!p.G && (chunks + 31) / 32 <= kmax) p.G = 32;
  }
  if (!p.G) return fail(EDHMC_ERR_INVALID, "n_features=%d exceeds the supported maximum of %d", D, kMaxFeatures);
  p.Kact = (chunks + p.G - 1) / p.G;
  static const int tiers4[] = {1, 2, 3, 4, 5, 6, 7, 8, 12, 16}, tiers2[] = {1, 2, 4, 6, 8, 10, 12, 14, 16, 24, 27, 32},
                   tiers1[] = {1, 2, 4, 8, 16, 28, 32, 64};
  const int* tiers = p.V == 4 ? tiers4 : (p.V == 2 ? tiers2 : tiers1);
  const int ntier = p.V == 4 ? 10 : (p.V == 2 ? 12 : 8);
  p.K = 0;
  for (int i = 0; i < ntier; ++i)
    if (tiers[i] >= p.Kact) {
      p.K = tiers[i];
      break;
    }
  if (!p.K) return fail(EDHMC_ERR_INVALID, "internal: no tier for %d chunks", p.Kact);
  p.NW = force_nw > 0 ? force_nw : warps_for(p.K * p.V);
  p.wpad = p.G * p.K * p.V;
  const int RPS = 32 / p.G;
  const long long row_bytes = ldx * 4;
  const size_t budget = static_cast<size_t>(h->smem_optin) - 1024;
  size_t offs[8];
  size_t fixed = smem_layout_bytes(p.NW, 0, 0, h->P, p.wpad, offs);
  if (fixed + 4096 > budget) return fail(EDHMC_ERR_INVALID, "chain state does not fit shared memory");
  // Tile size. Every tile costs ~150 instructions of ring bookkeeping per warp, so small tiles make the pass
  // issue-bound instead of HBM-bound (D=1000, one 4 KB row per tile, 12 warps: 88 % of HBM peak; two rows per tile,
  // 8 warps, double-buffered: 104 %). Default: three stages per warp with the default warp count; when that leaves
  // tiles under 6 KB, trade warps for tile size (8 warps, three stages, or two if that is what doubles the tile).
  // Not for odd row strides (V = 1: bound by 32-bit shared loads) and not for lanes that hold fewer than 20 floats of a
  // row (D = 8, 16, 32: the per-row link-function work dominates there): both want the warps (measured, profiles/README).
  auto rows_for = [&](int nw, int stages, long long cap) -> long long {
    const size_t fx = smem_layout_bytes(nw, 0, 0, h->P, p.wpad, offs);
    long long target = static_cast<long long>((budget - fx) / (static_cast<size_t>(nw) * stages)) - 192;
    if (target > cap) target = cap;
    long long j = target / (RPS * row_bytes);
    return j < 1 ? 1 : (j > 8 ? 8 : j);
  };
  long long J = rows_for(p.NW, 3, 7168);
  int bigtile = 1;
  if (const char* e = getenv("EDHMC_BIGTILE")) bigtile = atoi(e);
  if (bigtile && force_nw == 0 && p.NW != 8 && p.V >= 2 && p.K * p.V >= 20 && J * RPS * row_bytes < 6000 && lookup_kernel(p.G, p.V, p.K, 8)) {
    const long long j3 = rows_for(8, 3, 8192 + 64), j2 = rows_for(8, 2, 12288);
    long long jb = j3;
    if (j3 * RPS * row_bytes < 6000 && j2 > j3) jb = j2;
    if (jb * RPS * row_bytes >= J * RPS * row_bytes * 3 / 2) {
      p.NW = 8;
      J = jb;
      fixed = smem_layout_bytes(p.NW, 0, 0, h->P, p.wpad, offs);
    }
  }
  if (const char* e = getenv("EDHMC_FORCE_J")) J = atoi(e) > 0 ? atoi(e) : J;  // development: rows-per-tile multiplier
  p.fn = lookup_kernel(p.G, p.V, p.K, p.NW);
  if (!p.fn) return fail(EDHMC_ERR_INVALID, "no kernel compiled for G=%d V=%d K=%d NW=%d", p.G, p.V, p.K, p.NW);
  p.J = static_cast<int>(J);
  p.RT = RPS * p.J;
  p.tl = static_cast<int>(p.RT * ldx);
  p.tm = p.tl & 3;
  // x region: up to 3 floats of alignment skew + the tile + the over-read of the padded chunks / 16-byte round-up
  long long x_floats = 3 + static_cast<long long>(p.RT - 1) * ldx + p.wpad;
  const long long copy_floats = 3 + static_cast<long long>(p.RT) * ldx + 4;
  if (x_floats < copy_floats) x_floats = copy_floats;
  x_floats = (x_floats + 3) / 4 * 4;
  p.y_off = static_cast<int>(x_floats);
  long long sf = x_floats + p.RT;
  sf = (sf + 31) / 32 * 32;
  p.stage_floats = static_cast<int>(sf);

  int S = kMaxStages;
  for (; S >= 1; --S)
    if (smem_layout_bytes(p.NW, S, p.stage_floats, h->P, p.wpad, offs) <= budget) break;
  if (S < 2) return fail(EDHMC_ERR_INVALID, "row of %lld bytes does not fit the shared-memory ring", row_bytes);
  p.S = S;
  p.smem = smem_layout_bytes(p.NW, p.S, p.stage_floats, h->P, p.wpad, offs);
  cudaError_t e = cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem));
  if (e != cudaSuccess) return fail(EDHMC_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p.fn, p.NW * 32, p.smem);
  if (e != cudaSuccess) return fail(EDHMC_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e));
  if (per_sm < 1) return fail(EDHMC_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", p.smem);
  // grid: one CTA per SM, fewer when there are not enough rows to give every warp two tiles
  long long want = (c.n_rows + static_cast<long long>(p.NW) * 2 * p.RT - 1) / (static_cast<long long>(p.NW) * 2 * p.RT);
  if (want < 1) want = 1;
  p.grid = static_cast<int>(want < h->num_sms ? want : h->num_sms);
  h->plan = p;
  return 0;
}

// Ring mode 2 (stream_ldg.cuh): narrow rows (D <= 64) read from a bind-time re-lay of X with coalesced LDG.64, one lane per
// row, theta in shared memory, 12 warps per CTA, no shared-memory staging.
static bool make_plan_ldg(edhmc_handle* h, Plan& out) {
  const edhmc_cfg& c = h->cfg;
  const int D = c.n_features;
  if (D > 64) return false;
  static const int tiers2[] = {1, 2, 4, 6, 8, 10, 12, 14, 16, 24, 27, 32};
  const int chunks = (D + 1) / 2;
  int K = 0;
  for (int t : tiers2)
    if (t >= chunks) {
      K = t;
      break;
    }
  if (!K) return false;
  Plan p;
  p.RM = 2;
  p.G = 1;
  p.V = 2;
  p.K = K;
  p.Kact = chunks;
  p.NW = 12;
  p.WPG = 12;
  p.J = 1;
  p.RT = 32;
  p.S = 0;
  p.stage_floats = 0;
  p.wpad = (2 * K + 3) / 4 * 4;  // theta is read with LDS.128
  p.fn = lookup_kernel(1, 2, K, p.NW, 2);
  if (!p.fn) return false;
  const long long n_tiles = (c.n_rows + 31) / 32;
  long long want = (n_tiles + 2ll * p.NW - 1) / (2ll * p.NW);  // >= two tiles per warp before another SM is used
  if (want < 1) want = 1;
  p.grid = static_cast<int>(want < h->num_sms ? want : h->num_sms);
  size_t offs[8];
  // Shared memory that is not needed for staging keeps the first tiles of the CTA's range resident for the whole
  // persistent launch (stream_ldg.cuh, ldg_load_resident): EDHMC_RESIDENT=0 switches it off (A/B runs).
  {
    const size_t fixed = smem_layout_bytes(1, 0, 0, h->P, p.wpad, offs);
    const size_t tile_bytes = static_cast<size_t>(chunks) * 256 + 128;
    const long long per = (n_tiles + p.grid - 1) / p.grid;
    // tensor memory (one tile per warp and round only: K >= 14): 4 lane quarters x (512 / columns per row) slots
    long long tm = K >= 14 ? 4ll * (512 / ((2 * K + 7) / 8 * 8)) : 0;
    if (tm > per) tm = per;
    size_t budget = static_cast<size_t>(h->smem_optin) > fixed + 2048 ? static_cast<size_t>(h->smem_optin) - fixed - 2048 : 0;
    budget = budget > static_cast<size_t>(tm) * 128 ? budget - static_cast<size_t>(tm) * 128 : 0;  // y of the TMEM tiles
    long long r = static_cast<long long>(budget / tile_bytes);
    if (r > per - tm) r = per - tm;
    // worth it only when a sizeable part of X stays on chip (same-box A/B, profiles/README round 2: L2-resident shapes gain
    // 5-13 %, HBM-bound ones with < 7 % of their tiles resident lose 0-13 % to the per-tile branch)
    if ((r + tm) * 10 < per) r = tm = 0;
    if (const char* e = getenv("EDHMC_RESIDENT")) {
      const long long lim = atoll(e);
      if (lim >= 0 && r > lim) r = lim;
    }
    if (const char* e = getenv("EDHMC_TMEM")) {
      const long long lim = atoll(e);
      if (lim >= 0 && tm > lim) tm = lim;
    }
    p.n_res = static_cast<int>(r);
    p.n_tm = static_cast<int>(tm);
    if (p.n_res > 0 || p.n_tm > 0) {
      p.S = 1;
      p.stage_floats = static_cast<int>(r * (tile_bytes / 4) + tm * 32);
    }
  }
  p.smem = smem_layout_bytes(1, p.S, p.stage_floats, h->P, p.wpad, offs);
  if (cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p.fn, p.NW * 32, p.smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return false;
  }
  out = p;
  return true;
}

// Which narrow shapes take ring mode 2 (tools/shape_sweep_ldg.sh, same-box A/B, profiles/README round 2): every D <= 48 and
// D = 63, 64 — there it runs at 98-109 % of the HBM copy peak against 58-94 % for the shared-memory rings (power-of-two
// strides conflict in shared memory, rows of few columns leave the rings latency-bound), and L2-resident sizes gain
// 1.3-1.9x. For 49 <= D <= 62 the CTA-wide ring is as fast or faster when X streams from HBM; ring mode 2 is taken there
// when at least 30 % of the CTA's tiles stay resident in shared / tensor memory (see below).
// EDHMC_RING: unset = that rule; 2 = ring mode 2 wherever eligible; 1 / 0 = row-major plans only (A/B runs).
static int make_plan(edhmc_handle* h) {
  if (int rc = make_plan_rows(h)) return rc;
  h->plan_rows = h->plan;
  int ring = -1;
  if (const char* e = getenv("EDHMC_RING")) ring = atoi(e);
  const int D = h->cfg.n_features;
  Plan pl;
  if (ring != 2 && ring >= 0) return 0;
  if (h->interleave || !make_plan_ldg(h, pl)) return 0;
  bool want = ring == 2 || D <= 48 || D >= 63;
  if (!want) {
    // 49 <= D <= 62: the CTA-wide ring streams faster from the L2 / HBM, ring mode 2 wins once a good part of X stays in
    // shared and tensor memory for the whole launch (cfg 2: 65 of 123 tiles per SM, 14.6 against 16.1 us per step)
    const long long n_tiles = (h->cfg.n_rows + 31) / 32;
    const long long per = (n_tiles + pl.grid - 1) / pl.grid;
    want = 10ll * (pl.n_res + pl.n_tm) >= 3 * per;  // 1M x 54 (31 % resident): 108 vs 104 % of the HBM copy peak; 1M x 60 (27 %): 95 vs 102 %
  }
  if (want) h->plan = pl;
  return 0;
}

static void fill_args(edhmc_handle* h, KArgs& a, const Plan* pp = nullptr) {
  memset(&a, 0, sizeof(a));
  const edhmc_cfg& c = h->cfg;
  a.X = h->X;
  a.y = h->y;
  a.n_rows = c.n_rows;
  a.ldx = c.ldx;
  a.D = c.n_features;
  a.P = h->P;
  a.has_bias = c.has_bias;
  a.family = c.family;
  a.y_dtype = h->y_dtype;
  a.lik_scale = c.lik_scale;
  a.prior_loc = h->d_prior_loc;
  a.prior_scale = h->d_prior_scale;
  a.prior_kind = h->d_prior_kind;
  a.prior_const = h->prior_const;
  const Plan& p = pp ? *pp : h->plan;
  a.Xt = h->d_xt;
  a.Yt = h->d_yt;
  a.n_tiles = h->n_tiles;
  a.n_res = p.RM == 2 ? p.n_res : 0;
  a.n_tm = p.RM == 2 ? p.n_tm : 0;
  a.tm_on = a.n_tm > 0 ? 1 : 0;
  a.Kact = p.Kact;
  a.J = p.J;
  a.RT = p.RT;
  a.S = p.S;
  a.stage_floats = p.stage_floats;
  a.y_off = p.y_off;
  a.wpad = p.wpad;
  a.zigzag = h->zigzag;
  a.l2_hint = h->l2_hint;
  a.l2_frac = h->l2_frac;
  a.ldx_i = static_cast<int>(c.ldx);
  a.tl = p.tl;
  a.tm = p.tm;
  a.interleave = p.RM == 1 ? 0 : h->interleave;
  a.wpg = p.WPG > 0 ? p.WPG : 1;
  a.timeline = h->timeline;
  a.tl_cap = h->tl_cap;
  a.partials = h->d_partials;
  a.bar = h->d_bar;
  a.ticket = h->d_ticket;
  a.sums = h->d_sums;
  a.sc = h->d_sc;
  a.zcur = h->d_zcur;
  a.gcur = h->d_gcur;
  a.z = h->d_z;
  a.r = h->d_r;
  a.seed = h->seed;
  a.trace_scalars = h->trace_scalars;
  a.trace_pos = h->trace_pos;
  a.nranks = 1;  // the in-kernel exchange is switched on by edhmc_run for the persistent plan only
  a.rank = h->rank;
  a.peer_inbox = h->d_peer_ptrs;
  a.comm_seq = h->d_comm_seq;
  a.abort_flag = h->d_abort;
  a.spin_limit = h->spin_limit;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int edhmc_version(void) { return EDHMC_VERSION_MAJOR * 1000 + EDHMC_VERSION_MINOR; }

const char* edhmc_last_error(void) { return g_last_error.c_str(); }

int edhmc_create(edhmc_t** out, const edhmc_cfg* cfg) {
  if (!out || !cfg) return fail(EDHMC_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_rows < 0 || cfg->n_rows_global < cfg->n_rows) return fail(EDHMC_ERR_INVALID, "bad n_rows / n_rows_global");
  if (cfg->n_features < 1 || cfg->n_features > kMaxFeatures)
    return fail(EDHMC_ERR_INVALID, "n_features must be in [1, %d], got %d", kMaxFeatures, cfg->n_features);
  if (cfg->ldx < cfg->n_features) return fail(EDHMC_ERR_INVALID, "ldx (%lld) < n_features (%d)", (long long)cfg->ldx, cfg->n_features);
  if (cfg->ldx > (1 << 20)) return fail(EDHMC_ERR_INVALID, "ldx too large");
  if (cfg->family < 0 || cfg->family > 2) return fail(EDHMC_ERR_INVALID, "unknown family %d", cfg->family);
  if (cfg->dtype != EDHMC_F32 && cfg->dtype != EDHMC_F64) return fail(EDHMC_ERR_INVALID, "unknown dtype %d", cfg->dtype);
  if (cfg->y_dtype < 0 || cfg->y_dtype > 3) return fail(EDHMC_ERR_INVALID, "unknown y_dtype %d", cfg->y_dtype);
  if (cfg->dtype == EDHMC_F64) {
    if (cfg->y_dtype != EDHMC_Y_I32 && cfg->y_dtype != EDHMC_Y_F64) return fail(EDHMC_ERR_INVALID, "float64 models take int32 or float64 y");
    if (cfg->n_chains > 1) return fail(EDHMC_ERR_INVALID, "vectorised chains are float32 only");
    if (pass64_warps(cfg->n_features + (cfg->has_bias ? 1 : 0)) < 1)
      return fail(EDHMC_ERR_INVALID, "float64 models support up to 25,000 latent dimensions");
  } else if (cfg->y_dtype == EDHMC_Y_F64) {
    return fail(EDHMC_ERR_INVALID, "float64 y needs a float64 model");
  }
  if (cfg->family == EDHMC_NORMAL_IDENTITY && !(cfg->lik_scale > 0.0f))
    return fail(EDHMC_ERR_INVALID, "lik_scale must be > 0 for the Normal family");
  if (!cfg->prior_loc_host || !cfg->prior_scale_host) return fail(EDHMC_ERR_INVALID, "prior arrays are required");
  for (int i = 0; i < 3; ++i)
    if (cfg->reserved[i] != 0) return fail(EDHMC_ERR_INVALID, "reserved fields must be zero");
  const int P = cfg->n_features + (cfg->has_bias ? 1 : 0);
  double pc = 0.0;
  for (int i = 0; i < P; ++i) {
    if (!(cfg->prior_scale_host[i] > 0.0f) || !isfinite(cfg->prior_scale_host[i]) || !isfinite(cfg->prior_loc_host[i]))
      return fail(EDHMC_ERR_INVALID, "prior_scale[%d] must be finite and > 0, prior_loc finite", i);
    pc += 0.5 * log(2.0 * M_PI) + log(static_cast<double>(cfg->prior_scale_host[i]));
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(EDHMC_ERR_INVALID, "device %d out of range (%d CUDA devices visible)", cfg->device, ndev);
  CUDA_TRY(cudaSetDevice(cfg->device));

  edhmc_handle* h = new edhmc_handle();
  h->cfg = *cfg;
  h->cfg.prior_loc_host = nullptr;
  h->cfg.prior_scale_host = nullptr;
  h->P = P;
  h->prior_const = pc;
  h->prior_loc_h.assign(cfg->prior_loc_host, cfg->prior_loc_host + P);
  h->prior_scale_h.assign(cfg->prior_scale_host, cfg->prior_scale_host + P);
  if (const char* e = getenv("EDHMC_ZIGZAG")) h->zigzag = atoi(e);
  if (const char* e = getenv("EDHMC_L2_HINT")) h->l2_hint = atoi(e);
  if (const char* e = getenv("EDHMC_L2_FRAC")) h->l2_frac = static_cast<float>(atof(e));
  if (const char* e = getenv("EDHMC_INTERLEAVE")) h->interleave = atoi(e);
  cudaError_t e1 = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  cudaError_t e2 = cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    delete h;
    return fail(EDHMC_ERR_CUDA, "cudaDeviceGetAttribute failed");
  }
  h->f64 = cfg->dtype == EDHMC_F64;
  int rc = h->f64 ? 0 : make_plan(h);
  if (rc) {
    delete h;
    return rc;
  }
#define ALLOC(ptr, bytes)                                                                    \
  do {                                                                                       \
    cudaError_t e__ = h_alloc(h, reinterpret_cast<void**>(&(ptr)), (bytes));                 \
    if (e__ != cudaSuccess) {                                                                \
      edhmc_destroy(h);                                                                      \
      return fail(EDHMC_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e__)); \
    }                                                                                        \
  } while (0)
  const size_t pb = static_cast<size_t>(P) * sizeof(float);
  {
    const size_t cap = static_cast<size_t>(2) * h->num_sms * (P + 1) * sizeof(double) + 40 * (static_cast<size_t>(P) + 64) * sizeof(double) + 65536;
    if (cudaMalloc(reinterpret_cast<void**>(&h->arena), cap) == cudaSuccess) {
      h->arena_cap = cap;
    } else {
      cudaGetLastError();
      h->arena = nullptr;  // fall back to one allocation per buffer
    }
  }
  ALLOC(h->d_prior_loc, pb);
  ALLOC(h->d_prior_scale, pb);
  ALLOC(h->d_sc, sizeof(ChainScalars));
  ALLOC(h->d_zcur, pb);
  ALLOC(h->d_gcur, pb);
  ALLOC(h->d_z, pb);
  ALLOC(h->d_r, pb);
  ALLOC(h->d_g, pb);
  h->partials_cap = static_cast<size_t>(2) * h->num_sms * (P + 1);
  ALLOC(h->d_partials, h->partials_cap * sizeof(double));
  ALLOC(h->d_bar, sizeof(unsigned long long));
  if (P + 1 <= kWideCols) {
    const size_t pbytes = static_cast<size_t>(2) * h->num_sms * (P + 1) * sizeof(uint4);
    ALLOC(h->d_ll_part, pbytes);
    const size_t gbytes = static_cast<size_t>(2) * ((h->num_sms + kLlGroup - 1) / kLlGroup) * (P + 1) * sizeof(uint4);
    ALLOC(h->d_ll_group, gbytes);
    ALLOC(h->d_ll_theta, static_cast<size_t>(2) * kLlCopies * P * sizeof(uint2));
    cudaMemset(h->d_ll_group, 0, gbytes);
    cudaMemset(h->d_ll_part, 0, pbytes);
    cudaMemset(h->d_ll_theta, 0, static_cast<size_t>(2) * kLlCopies * P * sizeof(uint2));
  }
  if (const char* e = getenv("EDHMC_LEADER")) h->leader = atoi(e);
  ALLOC(h->d_ticket, sizeof(unsigned int));
  ALLOC(h->d_sums, static_cast<size_t>(P + 1) * sizeof(double));
  ALLOC(h->d_bad, sizeof(unsigned long long));
  if (h->f64) {
    const size_t pd = static_cast<size_t>(P) * sizeof(double);
    ALLOC(h->d64_zcur, pd);
    ALLOC(h->d64_gcur, pd);
    ALLOC(h->d64_z, pd);
    ALLOC(h->d64_r, pd);
    ALLOC(h->d64_g, pd);
    ALLOC(h->d64_partials, static_cast<size_t>(h->num_sms) * 4 * (P + 1) * sizeof(double));
    if (cudaFuncSetAttribute(k64_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10) != cudaSuccess) {
      edhmc_destroy(h);
      return fail(EDHMC_ERR_CUDA, "cudaFuncSetAttribute(k64_pass) failed");
    }
    cudaMemset(h->d64_zcur, 0, pd);
    cudaMemset(h->d64_gcur, 0, pd);
  }
  if (cfg->n_chains > 1) {
    h->C = cfg->n_chains;
    h->Cpad = (h->C + kMcChainsPerCta - 1) / kMcChainsPerCta * kMcChainsPerCta;
    h->mc_Dp = (P + 7) / 8 * 8;  // a bias latent is one more column (of ones) of the pre-tiled operand
    const long long ntiles = (cfg->n_rows + kMcTileRows - 1) / kMcTileRows;
    long long nrg = h->num_sms / (h->Cpad / kMcChainsPerCta);
    if (nrg < 1) nrg = 1;
    if (nrg > ntiles) nrg = ntiles > 0 ? ntiles : 1;
    h->mc_nrg = static_cast<int>(nrg);
    // pass implementation: 3 = pre-tiled pipelined tcgen05 (default), 0 = CUDA cores (cross-check)
    h->mc_use_tc = 3;
    h->mc_wide = P > kMcMaxD;
    if (const char* e = getenv("EDHMC_MC_IMPL")) {
      if (strcmp(e, "wide") == 0) h->mc_wide = true;
      if (strcmp(e, "simple") == 0) h->mc_use_tc = 0;
      else if (strcmp(e, "tc") == 0) h->mc_use_tc = 3;
    }
    memset(&h->mcw, 0, sizeof(h->mcw));
    if (h->mc_wide) {
      McwArgs& w = h->mcw;
      w.n_rows = cfg->n_rows;
      w.D = P;
      w.Dx = cfg->n_features;
      w.C = h->Cpad;
      w.Cu = h->C;
      mcw_plan(w, h->num_sms);
      const McwSizes z = mcw_sizes(w);
      ALLOC(w.xk, z.xk);
      ALLOC(w.xt, z.xt);
      ALLOC(w.yt, z.yt);
      ALLOC(w.wt, z.wt);
      ALLOC(w.rp, z.rp);
      ALLOC(w.part_g64, z.part_g64);
      ALLOC(w.part_lp, z.part_lp);
      ALLOC(w.gsum, z.gsum);
      h->mc_use_tc = 0;
      h->mc_nrg = 1;
      cudaError_t e = mcw_prepare();
      if (e != cudaSuccess) {
        edhmc_destroy(h);
        return fail(EDHMC_ERR_CUDA, "wide many-chain setup failed: %s", cudaGetErrorString(e));
      }
    } else if (h->mc_use_tc == 3) {
      size_t ytb = 0;
      const size_t xtb = mc_pretile_bytes(cfg->n_rows, h->mc_Dp, &ytb);
      ALLOC(h->mc_xt, xtb + 4096);
      ALLOC(h->mc_yt, ytb + 256);
    }
    const size_t cd = static_cast<size_t>(h->Cpad) * P * sizeof(float);
    ALLOC(h->mc_z, cd);
    ALLOC(h->mc_r, cd);
    ALLOC(h->mc_g, cd);
    ALLOC(h->mc_zcur, cd);
    ALLOC(h->mc_gcur, cd);
    ALLOC(h->mc_logp, h->Cpad * sizeof(double));
    ALLOC(h->mc_kold, h->Cpad * sizeof(double));
    ALLOC(h->mc_logu, h->Cpad * sizeof(double));
    ALLOC(h->mc_nacc, h->Cpad * sizeof(long long));
    ALLOC(h->mc_flags, 4 * sizeof(int));
    ALLOC(h->mc_part_g, static_cast<size_t>(h->mc_nrg) * h->Cpad * h->mc_Dp * sizeof(double));
    ALLOC(h->mc_part_lp, static_cast<size_t>(h->mc_nrg) * h->Cpad * sizeof(double));
    cudaMemset(h->mc_zcur, 0, cd);
    cudaMemset(h->mc_gcur, 0, cd);
    cudaMemset(h->mc_logp, 0, h->Cpad * sizeof(double));
    cudaMemset(h->mc_nacc, 0, h->Cpad * sizeof(long long));
    cudaMemset(h->mc_z, 0, cd);  // the padding chains (C not a multiple of 128) stay at theta = 0 for good
    cudaMemset(h->mc_r, 0, cd);
    cudaMemset(h->mc_g, 0, cd);
    if (h->Cpad != h->C) {
      ALLOC(h->mc_theta_pad, cd);
      cudaMemset(h->mc_theta_pad, 0, cd);
    }
    cudaMemset(h->mc_flags, 0, 4 * sizeof(int));
    if (h->mc_use_tc) {
      cudaError_t e = mc_prepare_tc();
      if (e != cudaSuccess) {
        edhmc_destroy(h);
        return fail(EDHMC_ERR_CUDA, "tensor-core pass setup failed: %s", cudaGetErrorString(e));
      }
    }
  }
#undef ALLOC
  cudaMemcpy(h->d_prior_loc, cfg->prior_loc_host, pb, cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_prior_scale, cfg->prior_scale_host, pb, cudaMemcpyHostToDevice);
  {
    // inbox of the in-kernel reductions (second level of the wide single-GPU reduce; peer exchange when sharded)
    unsigned char* own[kMaxRanks] = {nullptr};
    // the inbox is its own allocation: cudaIpcGetMemHandle exports whole allocations
    if (cudaMalloc(&h->d_inbox, kInboxBytes) != cudaSuccess ||
        h_alloc(h, reinterpret_cast<void**>(&h->d_peer_ptrs), sizeof(own)) != cudaSuccess ||
        h_alloc(h, reinterpret_cast<void**>(&h->d_comm_seq), sizeof(unsigned long long)) != cudaSuccess ||
        h_alloc(h, reinterpret_cast<void**>(&h->d_abort), sizeof(int)) != cudaSuccess) {
      edhmc_destroy(h);
      return fail(EDHMC_ERR_NOMEM, "cudaMalloc failed (inbox)");
    }
    own[0] = h->d_inbox;
    h->own_peer_ptrs = h->d_peer_ptrs;
    h->own_comm_seq = h->d_comm_seq;
    cudaMemset(h->d_inbox, 0, kInboxBytes);
    cudaMemcpy(h->d_peer_ptrs, own, sizeof(own), cudaMemcpyHostToDevice);
    cudaMemset(h->d_comm_seq, 0, sizeof(unsigned long long));
    cudaMemset(h->d_abort, 0, sizeof(int));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->cfg.device);
    long long ms = 20000;
    if (const char* e = getenv("EDHMC_PEER_TIMEOUT_MS")) ms = atoll(e);
    h->spin_limit = ms * static_cast<long long>(khz);
  }
  cudaMemset(h->d_sc, 0, sizeof(ChainScalars));
  cudaMemset(h->d_zcur, 0, pb);
  cudaMemset(h->d_gcur, 0, pb);
  cudaMemset(h->d_bar, 0, sizeof(unsigned long long));
  cudaMemset(h->d_ticket, 0, sizeof(unsigned int));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    edhmc_destroy(h);
    return fail(EDHMC_ERR_CUDA, "initialisation failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

int edhmc_destroy(edhmc_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->comm && !h->shared_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < kMaxRanks; ++r)
    if (h->peer_mapped[r]) cudaIpcCloseMemHandle(h->peer_mapped[r]);
  h_free(h, h->d_inbox);
  h_free(h, h->own_peer_ptrs);
  h_free(h, h->own_comm_seq);
  h_free(h, h->d_abort);
  h_free(h, h->d_prior_loc);
  h_free(h, h->d_prior_scale);
  if (h->d_prior_kind) cudaFree(h->d_prior_kind);
  h_free(h, h->d_sc);
  h_free(h, h->d_zcur);
  h_free(h, h->d_gcur);
  h_free(h, h->d_z);
  h_free(h, h->d_r);
  h_free(h, h->d_g);
  h_free(h, h->d_partials);
  h_free(h, h->d_bar);
  h_free(h, h->d_ll_part);
  h_free(h, h->d_ll_theta);
  if (h->d_xt) cudaFree(h->d_xt);
  if (h->d_yt) cudaFree(h->d_yt);
  h_free(h, h->d_ll_group);
  h_free(h, h->d_ticket);
  h_free(h, h->d_sums);
  h_free(h, h->d_bad);
  h_free(h, h->d64_zcur);
  h_free(h, h->d64_gcur);
  h_free(h, h->d64_z);
  h_free(h, h->d64_r);
  h_free(h, h->d64_g);
  h_free(h, h->d64_partials);
  h_free(h, h->y_owned);
  h_free(h, h->mc_theta_pad);
  h_free(h, h->mc_z);
  h_free(h, h->mc_r);
  h_free(h, h->mc_g);
  h_free(h, h->mc_zcur);
  h_free(h, h->mc_gcur);
  h_free(h, h->mc_logp);
  h_free(h, h->mc_kold);
  h_free(h, h->mc_logu);
  h_free(h, h->mc_nacc);
  h_free(h, h->mc_flags);
  h_free(h, h->mc_part_g);
  h_free(h, h->mc_part_lp);
  h_free(h, h->mc_xt);
  h_free(h, h->mc_yt);
  h_free(h, h->mcw.xk);
  h_free(h, h->mcw.xt);
  h_free(h, h->mcw.yt);
  h_free(h, h->mcw.wt);
  h_free(h, h->mcw.rp);
  h_free(h, h->mcw.part_g64);
  h_free(h, h->mcw.part_lp);
  h_free(h, h->mcw.gsum);
  cudaFree(h->arena);
  delete h;
  return 0;
}

int edhmc_bind_data(edhmc_t* h, const float* X, const void* y, int check_finite, void* stream_) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->cfg.n_rows > 0 && (!X || !y)) return fail(EDHMC_ERR_INVALID, "X and y are required");
  if (h->f64) return fail(EDHMC_ERR_INVALID, "float64 model: use edhmc_bind_data_f64");
  if (reinterpret_cast<uintptr_t>(X) % 16 != 0) return fail(EDHMC_ERR_INVALID, "X must be 16-byte aligned");
  if (h->cfg.y_dtype != EDHMC_Y_U8 && reinterpret_cast<uintptr_t>(y) % 4 != 0)
    return fail(EDHMC_ERR_INVALID, "y must be 4-byte aligned");
  h->X = X;
  h->y = y;
  h->y_dtype = h->cfg.y_dtype;
  if (h->plan_rows.RM == 1 && h->cfg.y_dtype != EDHMC_Y_U8 && reinterpret_cast<uintptr_t>(y) % 16 != 0) {
    h->no_cta_ring = true;  // the CTA-wide ring bulk-copies y: fall back to the per-warp rings
    if (int rc = make_plan(h)) return rc;
  }
  if (h->cfg.y_dtype == EDHMC_Y_U8 && h->cfg.n_rows > 0) {
    if (!h->y_owned) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->y_owned), static_cast<size_t>(h->cfg.n_rows) * 4));
    k_u8_to_i32<<<h->num_sms * 4, 256, 0, stream>>>(reinterpret_cast<const unsigned char*>(y), h->y_owned, h->cfg.n_rows);
    CUDA_TRY(cudaGetLastError());
    h->y = h->y_owned;
    h->y_dtype = EDHMC_Y_I32;
  }
  if (h->plan.RM == 2) {
    // ring mode 2 reads its own column-pair-major copy of X / y (stream_ldg.cuh): re-lay once per bind
    const long long n_tiles = (h->cfg.n_rows + 31) / 32;
    if (n_tiles != h->n_tiles || (!h->d_xt && n_tiles > 0)) {
      if (h->d_xt) cudaFree(h->d_xt);
      if (h->d_yt) cudaFree(h->d_yt);
      h->d_xt = nullptr;
      h->d_yt = nullptr;
      h->n_tiles = n_tiles;
      if (n_tiles > 0) {
        const size_t xb = static_cast<size_t>(n_tiles) * h->plan.Kact * 32 * sizeof(float2);
        if (cudaMalloc(reinterpret_cast<void**>(&h->d_xt), xb) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&h->d_yt), static_cast<size_t>(n_tiles) * 32 * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          if (h->d_xt) cudaFree(h->d_xt);
          h->d_xt = nullptr;
          h->n_tiles = 0;
          h->plan = h->plan_rows;  // not enough memory for the copy: stream the caller's X through the shared-memory ring
        }
      }
    }
    if (h->plan.RM == 2 && n_tiles > 0) {
      k_relay_tiles<<<h->num_sms * 8, 256, 0, stream>>>(X, h->cfg.n_rows, h->cfg.ldx, h->cfg.n_features, h->y, h->y_dtype,
                                                       h->plan.Kact, n_tiles, h->d_xt, h->d_yt);
      CUDA_TRY(cudaGetLastError());
    }
  }
  h->bound = true;
  // data changed: the cached log joint / gradient no longer apply
  CUDA_TRY(cudaMemsetAsync(&h->d_sc->valid, 0, sizeof(int), stream));
  if (h->mc_flags) CUDA_TRY(cudaMemsetAsync(h->mc_flags, 0, 4 * sizeof(int), stream));
  h->mc_pretiled = false;
  h->mcw_pretiled = false;
  if (check_finite && h->cfg.n_rows > 0) {
    CUDA_TRY(cudaMemsetAsync(h->d_bad, 0, sizeof(unsigned long long), stream));
    k_check_finite<<<h->num_sms * 8, 256, 0, stream>>>(X, h->cfg.n_rows, h->cfg.ldx, h->cfg.n_features, h->y,
                                                        h->y_dtype, h->d_bad);
    CUDA_TRY(cudaGetLastError());
    unsigned long long bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, h->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (bad) {
      h->bound = false;
      return fail(EDHMC_ERR_NONFINITE, "Tensor had NaN or Inf values (%llu non-finite entries in X/y)", bad);
    }
  }
  return 0;
}

static int launch_pass(edhmc_handle* h, const KArgs& a, const float* theta, int gate, long long gsi, cudaStream_t stream,
                       int want_lp = 1, const Plan* pp = nullptr) {
  const Plan& pl = pp ? *pp : h->plan;
  KArgs aa = a;
  aa.mode = 1;
  aa.n_res = 0;  // one pass per launch: nothing to keep resident
  aa.tm_on = 0;  // (n_tm stays: the tile order is the same in both plans)
  aa.gate = gate;
  aa.single_lp = want_lp;
  aa.par0 = static_cast<int>(gsi & 1);
  aa.theta_in = theta;
  void* params[] = {&aa};
  CUDA_TRY(cudaLaunchKernel(pl.fn, dim3(pl.grid), dim3(pl.NW * 32), params, pl.smem, stream));
  ++h->launches_last;
  return 0;
}

static int allreduce_sums(edhmc_handle* h, cudaStream_t stream) {
  if (h->nranks > 1) {
    NCCL_TRY(g_nccl.AllReduce(h->d_sums, h->d_sums, static_cast<size_t>(h->P + 1), kNcclFloat64, kNcclSum, h->comm, stream));
    ++h->launches_last;
  }
  return 0;
}

// ---- float64 models (f64.cuh): the stepwise schedule in double ----
static void fill_args64(edhmc_handle* h, K64Args& a) {
  memset(&a, 0, sizeof(a));
  const edhmc_cfg& c = h->cfg;
  a.X = h->X64;
  a.y = h->y;
  a.n_rows = c.n_rows;
  a.ldx = c.ldx;
  a.D = c.n_features;
  a.P = h->P;
  a.has_bias = c.has_bias;
  a.family = c.family;
  a.y_is_f64 = c.y_dtype == EDHMC_Y_F64 ? 1 : 0;
  a.lik_scale = c.lik_scale;
  a.prior_loc = h->d_prior_loc;
  a.prior_scale = h->d_prior_scale;
  a.prior_kind = h->d_prior_kind;
  a.prior_const = h->prior_const;
  a.sums = h->d_sums;
  a.sc = h->d_sc;
  a.zcur = h->d64_zcur;
  a.gcur = h->d64_gcur;
  a.z = h->d64_z;
  a.r = h->d64_r;
  a.g = h->d64_g;
  a.seed = h->seed;
  a.trace_scalars = h->trace_scalars64;
  a.trace_pos = h->trace_pos64;
}

static int pass64(edhmc_handle* h, const K64Args& a, const double* theta, int gate, cudaStream_t stream) {
  const int nw = pass64_warps(a.P);
  long long want = (a.n_rows + 8 * nw - 1) / (8 * nw);
  if (want < 1) want = 1;
  const int grid = static_cast<int>(want < h->num_sms * 4 ? want : h->num_sms * 4);
  const size_t smem = static_cast<size_t>(nw) * (a.P + 1) * sizeof(double);
  k64_pass<<<grid, nw * 32, smem, stream>>>(a, theta, h->d64_partials, gate);
  k64_reduce<<<(a.P + 1 + 255) / 256, 256, 0, stream>>>(h->d64_partials, grid, a.P + 1, a.sums, a.sc, gate);
  CUDA_TRY(cudaGetLastError());
  h->launches_last += 2;
  return allreduce_sums(h, stream);
}

static int run64(edhmc_handle* h, double* params, int64_t ldp, int64_t t0, int64_t n_iter, double eps, int32_t n_steps,
                 const double* r0, const double* u, cudaStream_t stream) {
  K64Args a;
  fill_args64(h, a);
  a.params = params;
  a.ldp = ldp;
  a.t0 = t0;
  a.n_iter = n_iter;
  a.eps = eps;
  a.L = n_steps;
  a.r0 = r0;
  a.u = u;
  h->launches_last = 0;
  h->passes_last = n_iter * n_steps;
  h->plan_in_use = EDHMC_PLAN_STEPWISE;
  k64_check<<<1, kChainThreads, 0, stream>>>(a);
  if (int rc = pass64(h, a, h->d64_zcur, 1, stream)) return rc;
  k64_init_finish<<<1, kChainThreads, 0, stream>>>(a);
  for (int64_t it = 0; it < n_iter; ++it) {
    k64_begin<<<1, kChainThreads, 0, stream>>>(a, it);
    for (int s = 0; s < n_steps; ++s) {
      if (int rc = pass64(h, a, h->d64_z, 0, stream)) return rc;
      k64_leap<<<1, kChainThreads, 0, stream>>>(a, it, s);
    }
    h->launches_last += 1 + n_steps;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int edhmc_bind_data_f64(edhmc_t* h, const double* X, const void* y, int check_finite, void* stream_) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  if (!h->f64) return fail(EDHMC_ERR_INVALID, "float32 model: use edhmc_bind_data");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->cfg.n_rows > 0 && (!X || !y)) return fail(EDHMC_ERR_INVALID, "X and y are required");
  if (reinterpret_cast<uintptr_t>(X) % 8 != 0) return fail(EDHMC_ERR_INVALID, "X must be 8-byte aligned");
  h->X64 = X;
  h->y = y;
  h->bound = true;
  CUDA_TRY(cudaMemsetAsync(&h->d_sc->valid, 0, sizeof(int), stream));
  if (check_finite && h->cfg.n_rows > 0) {
    CUDA_TRY(cudaMemsetAsync(h->d_bad, 0, sizeof(unsigned long long), stream));
    k64_check_finite<<<h->num_sms * 4, 256, 0, stream>>>(h->X64, h->cfg.n_rows, h->cfg.ldx, h->cfg.n_features, y,
                                                        h->cfg.y_dtype == EDHMC_Y_F64 ? 1 : 0, h->d_bad);
    CUDA_TRY(cudaGetLastError());
    unsigned long long bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, h->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (bad) {
      h->bound = false;
      return fail(EDHMC_ERR_NONFINITE, "Tensor had NaN or Inf values (%llu non-finite entries in X/y)", bad);
    }
  }
  return 0;
}

int edhmc_logp_grad_f64(edhmc_t* h, const double* theta, double* logp, double* grad, void* stream_) {
  if (!h || !theta || !logp || !grad) return fail(EDHMC_ERR_INVALID, "null argument");
  if (!h->f64) return fail(EDHMC_ERR_INVALID, "float32 model: use edhmc_logp_grad");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data_f64 has not been called");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  K64Args a64;
  fill_args64(h, a64);
  h->launches_last = 0;
  h->passes_last = 1;
  if (int rc = pass64(h, a64, theta, 0, stream)) return rc;
  k64_logp_grad_finish<<<1, kChainThreads, 0, stream>>>(a64, theta, logp, grad);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int edhmc_run_f64(edhmc_t* h, double* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter, double step_size,
                  int32_t n_steps, const double* r0, const double* u, void* stream_) {
  if (!h || !params) return fail(EDHMC_ERR_INVALID, "null argument");
  if (!h->f64) return fail(EDHMC_ERR_INVALID, "float32 model: use edhmc_run");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data_f64 has not been called");
  if (ldp < h->P) return fail(EDHMC_ERR_INVALID, "ldp (%lld) < number of latent dimensions (%d)", (long long)ldp, h->P);
  if (n_steps < 0 || n_iter < 0 || t0 < 0) return fail(EDHMC_ERR_INVALID, "negative n_steps / n_iter / t0");
  if (t0 + n_iter > T)
    return fail(EDHMC_ERR_RANGE, "indices[0] = %lld is not in [0, %lld)", (long long)(t0 + n_iter - 1), (long long)T);
  if (n_iter == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (int rc = run64(h, params, ldp, t0, n_iter, step_size, n_steps, r0, u, stream)) return rc;
  if (h->cfg.debug) {
    ChainScalars sc;
    CUDA_TRY(cudaMemcpyAsync(&sc, h->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (sc.nonfinite) return fail(EDHMC_ERR_NONFINITE, "log joint density had NaN or Inf values");
  }
  return 0;
}

int edhmc_logp_grad(edhmc_t* h, const float* theta, double* logp, float* grad, void* stream_) {
  if (!h || !theta || !logp || !grad) return fail(EDHMC_ERR_INVALID, "null argument");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data has not been called");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->f64) return fail(EDHMC_ERR_INVALID, "float64 model: use edhmc_logp_grad_f64");
  h->launches_last = 0;
  h->passes_last = 1;
  KArgs a;
  fill_args(h, a);
  int rc = launch_pass(h, a, theta, 0, 0, stream);
  if (rc) return rc;
  rc = allreduce_sums(h, stream);
  if (rc) return rc;
  k_logp_grad_finish<<<1, kChainThreads, 0, stream>>>(a, theta, logp, grad);
  CUDA_TRY(cudaGetLastError());
  ++h->launches_last;
  return 0;
}

int edhmc_run(edhmc_t* h, float* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter, float step_size,
              int32_t n_steps, const float* r0, const float* u, void* stream_) {
  if (!h || !params) return fail(EDHMC_ERR_INVALID, "null argument");
  if (h->f64) return fail(EDHMC_ERR_INVALID, "float64 model: use edhmc_run_f64");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data has not been called");
  if (ldp < h->P) return fail(EDHMC_ERR_INVALID, "ldp (%lld) < number of latent dimensions (%d)", (long long)ldp, h->P);
  if (n_steps < 0 || n_iter < 0 || t0 < 0) return fail(EDHMC_ERR_INVALID, "negative n_steps / n_iter / t0");
  if (t0 + n_iter > T)
    return fail(EDHMC_ERR_RANGE, "indices[0] = %lld is not in [0, %lld)", (long long)(t0 + n_iter - 1), (long long)T);
  if (n_iter == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  KArgs a;
  fill_args(h, a);
  a.params = params;
  a.ldp = ldp;
  a.t0 = t0;
  a.n_iter = n_iter;
  a.eps = step_size;
  a.half_eps = 0.5f * step_size;
  a.L = n_steps;
  a.r0 = r0;
  a.u = u;
  h->launches_last = 0;
  h->passes_last = n_iter * n_steps;

  bool persistent = h->nranks == 1 || h->peers_ready;
  if (h->cfg.plan == EDHMC_PLAN_STEPWISE) persistent = false;
  if (h->cfg.plan == EDHMC_PLAN_PERSISTENT && !persistent)
    return fail(EDHMC_ERR_INVALID, "persistent plan over row shards needs edhmc_peer_attach");
  if (persistent && h->nranks > 1) a.nranks = h->nranks;
  h->plan_in_use = persistent ? EDHMC_PLAN_PERSISTENT : EDHMC_PLAN_STEPWISE;

  if (persistent) {
    CUDA_TRY(cudaMemsetAsync(h->d_bar, 0, sizeof(unsigned long long), stream));
    void* kp[] = {&a};
    a.mode = 0;
    if (h->leader && h->d_ll_part && a.nranks == 1 && h->plan.grid > 1 && h->plan.grid <= h->num_sms &&
        h->plan.grid <= kLlGroup * kLlMaxGroups && n_iter * n_steps < (1ll << 30)) {
      a.leader = h->leader >= 2 ? 2 : 1;
      a.ll_part = h->d_ll_part;
      a.ll_theta = h->d_ll_theta;
      a.ll_group = h->d_ll_group;
      a.ll_seq0 = h->ll_seq;
      h->ll_seq += static_cast<unsigned int>(n_iter * n_steps + 2);  // one number per pass of this launch, never reused
    }
    CUDA_TRY(cudaLaunchCooperativeKernel(h->plan.fn, dim3(h->plan.grid), dim3(h->plan.NW * 32), kp, h->plan.smem, stream));
    ++h->launches_last;
  } else {
    int rc;
    k_chain_check<<<1, kChainThreads, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    if ((rc = launch_pass(h, a, h->d_zcur, 1, t0 * n_steps - 1, stream))) return rc;
    if ((rc = allreduce_sums(h, stream))) return rc;
    k_chain_init_finish<<<1, kChainThreads, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    h->launches_last += 2;
    for (int64_t it = 0; it < n_iter; ++it) {
      k_chain_begin<<<1, kChainThreads, 0, stream>>>(a, it, h->d_g);
      CUDA_TRY(cudaGetLastError());
      ++h->launches_last;
      for (int s = 0; s < n_steps; ++s) {
        // the log joint is consumed only after the last leapfrog step (hmc.py:104-105)
        if ((rc = launch_pass(h, a, h->d_z, 0, (t0 + it) * n_steps + s, stream, s == n_steps - 1 ? 1 : 0))) return rc;
        if ((rc = allreduce_sums(h, stream))) return rc;
        k_chain_leap<<<1, kChainThreads, 0, stream>>>(a, it, s, h->d_g);
        CUDA_TRY(cudaGetLastError());
        ++h->launches_last;
      }
    }
  }
  if (h->cfg.debug) {
    ChainScalars sc;
    CUDA_TRY(cudaMemcpyAsync(&sc, h->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (sc.nonfinite) return fail(EDHMC_ERR_NONFINITE, "log joint density had NaN or Inf values");
  }
  return 0;
}

int edhmc_set_trace(edhmc_t* h, double* trace_scalars, float* trace_pos) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->trace_scalars = trace_scalars;
  h->trace_pos = trace_pos;
  return 0;
}

int edhmc_set_trace_f64(edhmc_t* h, double* trace_scalars, double* trace_pos) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->trace_scalars64 = trace_scalars;
  h->trace_pos64 = trace_pos;
  return 0;
}

int edhmc_set_prior_kinds(edhmc_t* h, const int32_t* kinds_host) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (!kinds_host) {
    if (h->d_prior_kind) cudaFree(h->d_prior_kind);
    h->d_prior_kind = nullptr;
  } else {
    if (h->C > 1) return fail(EDHMC_ERR_INVALID, "vectorised chains support Normal priors only");
    double pc = 0.0;
    for (int i = 0; i < h->P; ++i) {
      const double p0 = h->prior_loc_h[i], p1 = h->prior_scale_h[i];
      if (kinds_host[i] == 0) {
        pc += 0.5 * log(2.0 * M_PI) + log(p1);
      } else if (kinds_host[i] == 1) {
        if (!(p0 > 0.0)) return fail(EDHMC_ERR_INVALID, "Beta prior of latent %d needs a > 0 (prior_loc), got %g", i, p0);
        pc += lgamma(p0) + lgamma(p1) - lgamma(p0 + p1);  // lbeta(a, b)
      } else {
        return fail(EDHMC_ERR_INVALID, "unknown prior kind %d for latent %d", kinds_host[i], i);
      }
    }
    if (!h->d_prior_kind) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->d_prior_kind), sizeof(int) * h->P));
    CUDA_TRY(cudaMemcpy(h->d_prior_kind, kinds_host, sizeof(int) * h->P, cudaMemcpyHostToDevice));
    h->prior_const = pc;
  }
  CUDA_TRY(cudaMemset(&h->d_sc->valid, 0, sizeof(int)));  // the cached log joint no longer applies
  return 0;
}

int edhmc_set_timeline(edhmc_t* h, long long* buf, int32_t n_passes) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->timeline = buf;
  h->tl_cap = buf ? n_passes : 0;
  return 0;
}

int edhmc_read_state(edhmc_t* h, int64_t* n_accept_host, double* logp_host, void* stream_) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  ChainScalars sc;
  int aborted = 0;
  CUDA_TRY(cudaMemcpyAsync(&sc, h->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, stream));
  if (h->d_abort) CUDA_TRY(cudaMemcpyAsync(&aborted, h->d_abort, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (aborted) return fail(EDHMC_ERR_COMM, "timed out waiting for the shard totals of a peer rank");
  if (n_accept_host) *n_accept_host = sc.n_accept;
  if (logp_host) *logp_host = sc.logp_cur;
  return 0;
}

int edhmc_reset(edhmc_t* h, void* stream_) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaMemsetAsync(h->d_sc, 0, sizeof(ChainScalars), stream));
  if (h->C > 1) {
    CUDA_TRY(cudaMemsetAsync(h->mc_nacc, 0, h->Cpad * sizeof(long long), stream));
    CUDA_TRY(cudaMemsetAsync(h->mc_flags, 0, 4 * sizeof(int), stream));
  }
  return 0;
}

int edhmc_seed(edhmc_t* h, uint64_t seed) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->seed = seed;
  return 0;
}

int edhmc_comm_unique_id(void* id128_host) {
  if (!id128_host) return fail(EDHMC_ERR_INVALID, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  NCCL_TRY(g_nccl.GetUniqueId(reinterpret_cast<NcclUid*>(id128_host)));
  return 0;
}

int edhmc_comm_cached(int32_t device, int32_t nranks, int32_t rank) {
  if (device < 0 || device >= 64) return 0;
  const SharedComm& sc = g_shared[device];
  if (!sc.comm || sc.nranks != nranks || sc.rank != rank) return 0;
  return sc.peers_ready ? 2 : 1;
}

int edhmc_comm_release(int32_t device) {
  if (device < 0 || device >= 64) return fail(EDHMC_ERR_INVALID, "device out of range");
  cudaSetDevice(device);
  shared_release(g_shared[device]);
  return 0;
}

int edhmc_comm_init(edhmc_t* h, const void* id128_host, int32_t nranks, int32_t rank) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(EDHMC_ERR_INVALID, "bad nranks/rank");
  if (h->cfg.device >= 64) return fail(EDHMC_ERR_INVALID, "device ordinal too large");
  int rc = nccl_load();
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  SharedComm& sc = g_shared[h->cfg.device];
  if (!id128_host) {  // reuse the communicator an earlier handle of this process created
    if (!sc.comm || sc.nranks != nranks || sc.rank != rank)
      return fail(EDHMC_ERR_STATE, "no cached communicator for device %d with nranks=%d rank=%d", h->cfg.device, nranks, rank);
  } else {
    if (sc.comm) shared_release(sc);  // a different world: start over
    NcclUid id;
    memcpy(&id, id128_host, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&sc.comm, nranks, id, rank));
    sc.nranks = nranks;
    sc.rank = rank;
  }
  h->comm = sc.comm;
  h->shared_comm = true;
  h->nranks = nranks;
  h->rank = rank;
  return 0;
}

static int shared_inbox(SharedComm& sc) {
  if (sc.inbox) return 0;
  if (cudaMalloc(&sc.inbox, kInboxBytes) != cudaSuccess || cudaMalloc(&sc.d_peer_ptrs, sizeof(unsigned char*) * kMaxRanks) != cudaSuccess ||
      cudaMalloc(&sc.d_seq, sizeof(unsigned long long)) != cudaSuccess)
    return fail(EDHMC_ERR_NOMEM, "cudaMalloc failed (shared inbox)");
  cudaMemset(sc.inbox, 0, kInboxBytes);
  cudaMemset(sc.d_seq, 0, sizeof(unsigned long long));
  return 0;
}

int edhmc_peer_export(edhmc_t* h, void* handle64_host) {
  if (!h || !handle64_host) return fail(EDHMC_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (h->cfg.device >= 64) return fail(EDHMC_ERR_INVALID, "device ordinal too large");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  SharedComm& sc = g_shared[h->cfg.device];
  if (int rc = shared_inbox(sc)) return rc;
  cudaIpcMemHandle_t ipc;
  CUDA_TRY(cudaIpcGetMemHandle(&ipc, sc.inbox));
  memcpy(handle64_host, &ipc, sizeof(ipc));
  return 0;
}

int edhmc_peer_attach(edhmc_t* h, const void* handles_host, int32_t nranks, int32_t rank) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null argument");
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
    return fail(EDHMC_ERR_INVALID, "peer exchange supports 1..%d ranks, got nranks=%d rank=%d", kMaxRanks, nranks, rank);
  if (h->comm && (nranks != h->nranks || rank != h->rank)) return fail(EDHMC_ERR_INVALID, "nranks/rank differ from edhmc_comm_init");
  if (h->cfg.device >= 64) return fail(EDHMC_ERR_INVALID, "device ordinal too large");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  SharedComm& sc = g_shared[h->cfg.device];
  if (!handles_host) {  // reuse the mapping an earlier handle of this process made
    if (!sc.peers_ready || sc.nranks != nranks || sc.rank != rank)
      return fail(EDHMC_ERR_STATE, "no cached peer mapping for device %d with nranks=%d rank=%d", h->cfg.device, nranks, rank);
  } else {
    if (int rc = shared_inbox(sc)) return rc;
    for (int r = 0; r < kMaxRanks; ++r)
      if (sc.mapped[r]) {
        cudaIpcCloseMemHandle(sc.mapped[r]);
        sc.mapped[r] = nullptr;
      }
    unsigned char* ptrs[kMaxRanks] = {nullptr};
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) {
        ptrs[r] = sc.inbox;
        continue;
      }
      cudaIpcMemHandle_t ipc;
      memcpy(&ipc, static_cast<const unsigned char*>(handles_host) + 64 * r, sizeof(ipc));
      void* mapped = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&mapped, ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(EDHMC_ERR_COMM, "cannot map the inbox of rank %d (cudaIpcOpenMemHandle: %s)", r, cudaGetErrorString(e));
      }
      sc.mapped[r] = mapped;
      ptrs[r] = static_cast<unsigned char*>(mapped);
    }
    CUDA_TRY(cudaMemcpy(sc.d_peer_ptrs, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice));
    sc.nranks = nranks;
    sc.rank = rank;
    sc.peers_ready = true;
  }
  h->d_peer_ptrs = sc.d_peer_ptrs;
  h->d_comm_seq = sc.d_seq;
  h->nranks = nranks;
  h->rank = rank;
  h->peers_ready = true;
  return 0;
}

int edhmc_peer_detach(edhmc_t* h) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->peers_ready = false;
  h->d_peer_ptrs = h->own_peer_ptrs;
  h->d_comm_seq = h->own_comm_seq;
  if (h->cfg.device < 64) g_shared[h->cfg.device].peers_ready = false;
  return 0;
}

int edhmc_sgmcmc_run(edhmc_t* h, int32_t kind, float* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter,
                     float step_size, float friction, float lik_factor, const float* prior_factor, float* velocity,
                     const float* noise, int64_t batch_rows, void* stream_) {
  if (!h || !params) return fail(EDHMC_ERR_INVALID, "null argument");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data has not been called");
  if (kind != 0 && kind != 1) return fail(EDHMC_ERR_INVALID, "kind must be 0 (SGLD) or 1 (SGHMC)");
  if (h->f64) return fail(EDHMC_ERR_INVALID, "SGLD / SGHMC are float32 only");
  if (kind == 1 && !velocity) return fail(EDHMC_ERR_INVALID, "SGHMC needs a velocity buffer");
  if (ldp < h->P) return fail(EDHMC_ERR_INVALID, "ldp (%lld) < number of latent dimensions (%d)", (long long)ldp, h->P);
  if (n_iter < 0 || t0 < 0) return fail(EDHMC_ERR_INVALID, "negative n_iter / t0");
  if (t0 + n_iter > T)
    return fail(EDHMC_ERR_RANGE, "indices[0] = %lld is not in [0, %lld)", (long long)(t0 + n_iter - 1), (long long)T);
  if (batch_rows < 0 || batch_rows % 4 != 0 || batch_rows > h->cfg.n_rows)
    return fail(EDHMC_ERR_INVALID, "batch_rows must be a multiple of 4 in [0, n_rows]");
  if (n_iter == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  KArgs a;
  fill_args(h, a);
  a.params = params;
  a.ldp = ldp;
  a.t0 = t0;
  a.n_iter = n_iter;
  SgArgs g;
  g.kind = kind;
  g.step_size = step_size;
  g.friction = friction;
  g.lik_factor = lik_factor;
  g.prior_factor = prior_factor;
  g.velocity = velocity;
  g.noise = noise;
  h->launches_last = 0;
  h->passes_last = n_iter;
  h->plan_in_use = EDHMC_PLAN_STEPWISE;
  const int64_t nb = batch_rows > 0 ? h->cfg.n_rows / batch_rows : 1;
  // mini-batches are row windows of the caller's X: they take the row-major plan (ring mode 2 reads whole 32-row tiles of
  // its own copy)
  const Plan* pp = batch_rows > 0 ? &h->plan_rows : &h->plan;
  KArgs arows;
  if (batch_rows > 0) {
    fill_args(h, arows, pp);
    arows.params = a.params;
    arows.ldp = a.ldp;
    arows.t0 = a.t0;
    arows.n_iter = a.n_iter;
  }
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t t = t0 + it;
    const int64_t t_prev = t > 0 ? t - 1 : 0;
    KArgs ab = batch_rows > 0 ? arows : a;
    if (batch_rows > 0) {  // mini-batch: a contiguous slice of the bound rows
      const int64_t lo = (t % nb) * batch_rows;
      ab.X = a.X + lo * a.ldx;
      ab.y = reinterpret_cast<const char*>(a.y) + lo * 4;
      ab.n_rows = batch_rows;
    }
    int rc;
    if ((rc = launch_pass(h, ab, params + t_prev * ldp, 0, t, stream, 1, pp))) return rc;
    if ((rc = allreduce_sums(h, stream))) return rc;
    k_sg_update<<<1, kChainThreads, 0, stream>>>(a, g, it);
    CUDA_TRY(cudaGetLastError());
    ++h->launches_last;
  }
  return 0;
}

// ---- vectorised chains -----------------------------------------------------------------------------
static void fill_mc_args(edhmc_handle* h, McArgs& a) {
  memset(&a, 0, sizeof(a));
  const edhmc_cfg& c = h->cfg;
  a.X = h->X;
  a.y = h->y;
  a.n_rows = c.n_rows;
  a.ldx = c.ldx;
  a.D = h->P;
  a.Dx = c.n_features;
  a.Dp = h->mc_Dp;
  a.family = c.family;
  a.y_dtype = h->y_dtype;
  a.lik_scale = c.lik_scale;
  a.prior_loc = h->d_prior_loc;
  a.prior_scale = h->d_prior_scale;
  a.prior_const = h->prior_const;
  a.C = h->Cpad;
  a.Cu = h->C;
  a.n_rowgroups = h->mc_nrg;
  if (const char* e = getenv("EDHMC_MC_SEG_MODE")) a.seg_mode = atoi(e);
  if (const char* e = getenv("EDHMC_MC_SEG")) a.seg_tiles = atoi(e);  // development: TMEM accumulation length (tiles)
  a.want_logp = 1;
  a.z = h->mc_z;
  a.r = h->mc_r;
  a.g = h->mc_g;
  a.zcur = h->mc_zcur;
  a.gcur = h->mc_gcur;
  a.logp_cur = h->mc_logp;
  a.k_old = h->mc_kold;
  a.log_u = h->mc_logu;
  a.n_accept = h->mc_nacc;
  a.valid = h->mc_flags;
  a.need_init = h->mc_flags + 1;
  a.part_g = h->mc_part_g;
  a.part_lp = h->mc_part_lp;
  a.seed = h->seed;
  a.trace = h->mc_trace;
  a.dbg = h->mc_dbg;
  a.dbg_lp = getenv("EDHMC_MC_DBG_LP") ? atoi(getenv("EDHMC_MC_DBG_LP")) : 0;
  a.xt = h->mc_xt;
  a.yt = h->mc_yt;
}

static void fill_mcw_args(edhmc_handle* h, McwArgs& a) {
  a = h->mcw;  // tiling + buffers (C = Cpad, Cu = the caller's chains)
  const edhmc_cfg& c = h->cfg;
  a.X = h->X;
  a.y = h->y;
  a.ldx = c.ldx;
  a.family = c.family;
  a.y_dtype = h->y_dtype;
  a.lik_scale = c.lik_scale;
  a.prior_loc = h->d_prior_loc;
  a.prior_scale = h->d_prior_scale;
  a.prior_const = h->prior_const;
  a.want_logp = 1;
  a.z = h->mc_z;
  a.r = h->mc_r;
  a.g = h->mc_g;
  a.zcur = h->mc_zcur;
  a.gcur = h->mc_gcur;
  a.logp_cur = h->mc_logp;
  a.k_old = h->mc_kold;
  a.log_u = h->mc_logu;
  a.n_accept = h->mc_nacc;
  a.valid = h->mc_flags;
  a.need_init = h->mc_flags + 1;
  a.seed = h->seed;
  a.trace = h->mc_trace;
}

// One evaluation of the per-chain [grad, logp] likelihood sums at theta, summed over the row shards.
static int mcw_pass(edhmc_handle* h, const McwArgs& a, const float* theta, int gate, cudaStream_t stream) {
  if (!h->mcw_pretiled) {
    CUDA_TRY(mcw_launch_pretile(a, stream));
    h->mcw_pretiled = true;
    h->launches_last += 2;
  }
  CUDA_TRY(mcw_launch_pass(a, theta, gate, stream));
  h->launches_last += 4;
  if (h->nranks > 1) {
    NCCL_TRY(g_nccl.AllReduce(a.gsum, a.gsum, static_cast<size_t>(a.C) * (a.D + 1), kNcclFloat64, kNcclSum, h->comm, stream));
    ++h->launches_last;
  }
  return 0;
}

static int run_chains_wide(edhmc_handle* h, float* params, int64_t t0, int64_t n_iter, float step_size, int32_t n_steps,
                           const float* r0, const float* u, cudaStream_t stream) {
  McwArgs a;
  fill_mcw_args(h, a);
  a.params = params;
  a.t0 = t0;
  a.n_iter = n_iter;
  a.eps = step_size;
  a.half_eps = 0.5f * step_size;
  a.L = n_steps;
  a.r0 = r0;
  a.u = u;
  h->launches_last = 0;
  h->passes_last = n_iter * n_steps;
  CUDA_TRY(mcw_launch_check(a, stream));
  if (int rc = mcw_pass(h, a, h->mc_zcur, 1, stream)) return rc;
  CUDA_TRY(mcw_launch_init_finish(a, stream));
  h->launches_last += 2;
  for (int64_t it = 0; it < n_iter; ++it) {
    CUDA_TRY(mcw_launch_begin(a, it, stream));
    ++h->launches_last;
    for (int s = 0; s < n_steps; ++s) {
      McwArgs as = a;
      as.want_logp = (s == n_steps - 1) ? 1 : 0;
      if (int rc = mcw_pass(h, as, h->mc_z, 0, stream)) return rc;
      CUDA_TRY(mcw_launch_leap(a, it, s, stream));
      ++h->launches_last;
    }
  }
  return 0;
}

// Re-lays X into the tensor-core operand layout once per bound data set.
static int mc_ensure_pretiled(edhmc_handle* h, const McArgs& a, cudaStream_t stream) {
  if (h->mc_use_tc == 3 && !h->mc_pretiled) {
    CUDA_TRY(mc_launch_pretile(a, h->mc_xt, h->mc_yt, stream));
    h->mc_pretiled = true;
  }
  return 0;
}

int edhmc_run_chains(edhmc_t* h, float* params, int64_t T, int64_t t0, int64_t n_iter, float step_size, int32_t n_steps,
                     const float* r0, const float* u, void* stream_) {
  if (!h || !params) return fail(EDHMC_ERR_INVALID, "null argument");
  if (h->C <= 1) return fail(EDHMC_ERR_STATE, "handle was not created with n_chains > 1");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data has not been called");
  if (n_steps < 0 || n_iter < 0 || t0 < 0) return fail(EDHMC_ERR_INVALID, "negative n_steps / n_iter / t0");
  if (t0 + n_iter > T)
    return fail(EDHMC_ERR_RANGE, "indices[0] = %lld is not in [0, %lld)", (long long)(t0 + n_iter - 1), (long long)T);
  if (n_iter == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->mc_wide) return run_chains_wide(h, params, t0, n_iter, step_size, n_steps, r0, u, stream);
  if (h->nranks > 1) return fail(EDHMC_ERR_INVALID, "row-sharded vectorised chains use the wide path: set EDHMC_MC_IMPL=wide");
  McArgs a;
  fill_mc_args(h, a);
  a.params = params;
  a.t0 = t0;
  a.n_iter = n_iter;
  a.eps = step_size;
  a.half_eps = 0.5f * step_size;
  a.L = n_steps;
  a.r0 = r0;
  a.u = u;
  h->launches_last = 0;
  h->passes_last = n_iter * n_steps;
  if (int rc = mc_ensure_pretiled(h, a, stream)) return rc;
  CUDA_TRY(mc_launch_check(a, stream));
  CUDA_TRY(mc_launch_pass(a, h->mc_zcur, h->mc_use_tc, 1, stream));
  CUDA_TRY(mc_launch_init_finish(a, stream));
  h->launches_last += 3;
  for (int64_t it = 0; it < n_iter; ++it) {
    CUDA_TRY(mc_launch_begin(a, it, stream));
    ++h->launches_last;
    for (int s = 0; s < n_steps; ++s) {
      McArgs as = a;
      as.want_logp = (s == n_steps - 1) ? 1 : 0;  // the log joint is only needed at the end of the trajectory
      CUDA_TRY(mc_launch_pass(as, h->mc_z, h->mc_use_tc, 0, stream));
      CUDA_TRY(mc_launch_leap(a, it, s, stream));
      h->launches_last += 2;
    }
  }
  return 0;
}

int edhmc_logp_grad_chains(edhmc_t* h, const float* theta, double* logp, float* grad, void* stream_) {
  if (!h || !theta || !logp || !grad) return fail(EDHMC_ERR_INVALID, "null argument");
  if (h->C <= 1) return fail(EDHMC_ERR_STATE, "handle was not created with n_chains > 1");
  if (!h->bound) return fail(EDHMC_ERR_STATE, "edhmc_bind_data has not been called");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const float* theta_pass = theta;  // what the pass kernels read: whole 128-chain tiles
  if (h->mc_theta_pad) {
    CUDA_TRY(cudaMemcpyAsync(h->mc_theta_pad, theta, static_cast<size_t>(h->C) * h->P * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    theta_pass = h->mc_theta_pad;
  }
  if (h->mc_wide) {
    McwArgs w;
    fill_mcw_args(h, w);
    h->launches_last = 0;
    h->passes_last = 1;
    if (int rc = mcw_pass(h, w, theta_pass, 0, stream)) return rc;
    CUDA_TRY(mcw_launch_logp_grad_finish(w, theta, logp, grad, stream));
    ++h->launches_last;
    return 0;
  }
  McArgs a;
  fill_mc_args(h, a);
  if (int rc = mc_ensure_pretiled(h, a, stream)) return rc;
  CUDA_TRY(mc_launch_pass(a, theta_pass, h->mc_use_tc, 0, stream));
  CUDA_TRY(mc_launch_logp_grad_finish(a, theta, logp, grad, stream));
  h->launches_last = 2;
  h->passes_last = 1;
  return 0;
}

int edhmc_read_chain_state(edhmc_t* h, int64_t* n_accept_host, double* logp_host, void* stream_) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  if (h->C <= 1) return fail(EDHMC_ERR_STATE, "handle was not created with n_chains > 1");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (n_accept_host) CUDA_TRY(cudaMemcpyAsync(n_accept_host, h->mc_nacc, h->C * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  if (logp_host) CUDA_TRY(cudaMemcpyAsync(logp_host, h->mc_logp, h->C * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return 0;
}

int edhmc_set_chain_debug(edhmc_t* h, long long* buf) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->mc_dbg = buf;
  return 0;
}

int edhmc_set_chain_trace(edhmc_t* h, double* trace) {
  if (!h) return fail(EDHMC_ERR_INVALID, "null handle");
  h->mc_trace = trace;
  return 0;
}

int edhmc_chains_plan_probe(int64_t n_rows, int32_t n_features, int32_t n_chains, int32_t num_sms, int64_t* out8) {
  if (!out8) return fail(EDHMC_ERR_INVALID, "null argument");
  if (n_rows < 1 || n_features < 1 || n_features > kMaxFeatures || num_sms < 1)
    return fail(EDHMC_ERR_INVALID, "bad n_rows / n_features / num_sms");
  if (n_chains < 2) return fail(EDHMC_ERR_INVALID, "n_chains must be >= 2, got %d", n_chains);
  McwArgs w;
  memset(&w, 0, sizeof(w));
  w.n_rows = n_rows;
  w.D = n_features;
  w.Dx = n_features;
  w.C = (n_chains + kMcChainsPerCta - 1) / kMcChainsPerCta * kMcChainsPerCta;
  w.Cu = n_chains;
  mcw_plan(w, num_sms);
  const int64_t v[8] = {w.nct, w.Kp1, w.nrt, w.nft, w.NB2, w.g1, w.splits, w.Dp2};
  for (int i = 0; i < 8; ++i) out8[i] = v[i];
  return 0;
}

int edhmc_plan_info(edhmc_t* h, int64_t* out, int32_t cap) {
  if (!h || !out) return fail(EDHMC_ERR_INVALID, "null argument");
  const int64_t v[13] = {h->plan.grid,
                         h->plan.NW,
                         h->plan.S,
                         h->plan.RT,
                         h->plan.G,
                         h->plan.V,
                         static_cast<int64_t>(h->plan.smem),
                         h->plan_in_use,
                         h->passes_last,
                         h->launches_last,
                         h->plan.RM,
                         h->plan.RM == 2 ? h->plan.n_res : 0,
                         h->plan.RM == 2 ? h->plan.n_tm : 0};
  int n = cap < 13 ? cap : 13;
  for (int i = 0; i < n; ++i) out[i] = v[i];
  return n;
}

int edhmc_predictive(const float* X, int64_t n_rows, int64_t ldx, int32_t n_features, const void* y, int32_t y_dtype,
                     int32_t family, float lik_scale, const float* params, int64_t ldp, const int32_t* idx_w,
                     const int32_t* idx_b, int32_t bias_col, int32_t n_draws, float* mean_out, double* loglik_out,
                     int32_t device, void* stream) {
  if (!X || !params || !idx_w || !mean_out) return fail(EDHMC_ERR_INVALID, "null argument");
  if (n_rows < 0 || n_features < 1 || ldx < n_features || n_draws < 1 || ldp < n_features)
    return fail(EDHMC_ERR_INVALID, "bad shape (n_rows %lld, n_features %d, ldx %lld, n_draws %d, ldp %lld)", (long long)n_rows,
                n_features, (long long)ldx, n_draws, (long long)ldp);
  if (family < 0 || family > 2) return fail(EDHMC_ERR_INVALID, "unknown family %d", family);
  if (y_dtype < 0 || y_dtype > 2) return fail(EDHMC_ERR_INVALID, "unknown y_dtype %d", y_dtype);
  if (loglik_out && !y) return fail(EDHMC_ERR_INVALID, "the log-likelihood needs y");
  if (idx_b && (bias_col < 0 || bias_col >= ldp)) return fail(EDHMC_ERR_INVALID, "bias_col out of range");
  if (family == EDHMC_NORMAL_IDENTITY && !(lik_scale > 0.0f)) return fail(EDHMC_ERR_INVALID, "lik_scale must be > 0");
  if (n_rows == 0) return 0;
  CUDA_TRY(cudaSetDevice(device));
  const long long blocks = (n_rows + kPredRows - 1) / kPredRows;
  k_predictive<<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X, n_rows, ldx, n_features, y, y_dtype, family, lik_scale, params, ldp, idx_w, idx_b, bias_col, n_draws, mean_out,
      loglik_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
