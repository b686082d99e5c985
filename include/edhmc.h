/*
 * edhmc.h — C ABI of libedhmc.so: Edward's HMC hot path for GLM-style models on B200 (sm_100a).
 *
 * This is the drop-in boundary. The reference (blei-lab/edward 1.3.5) has no native code: its hot
 * path is a TensorFlow graph built by edward/inferences/hmc.py and executed by tf.Session.run. The
 * entry points below are what a ctypes binding on the reference side would call instead of
 * `sess.run(self.train)`; each one cites the reference interface it replaces.
 *
 * Conventions
 *   - C linkage, plain pointers and sizes. No torch / C++ types.
 *   - Every function returns 0 on success and a negative edhmc_status on failure;
 *     edhmc_last_error() gives a thread-local, human-readable message for the last failure.
 *   - Pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - The caller owns every buffer it passes; the library owns only the handle and its scratch.
 *   - A handle is not thread-safe. The library starts no host threads. All device work is queued on
 *     the caller-supplied stream (`void* stream` is a cudaStream_t; NULL = legacy default stream).
 *   - Floating point is float32 (the reference's dtype on this path); reductions are carried in
 *     float64 on the device.
 *
 * Model covered (anything else is rejected by the host front-end with NotImplementedError):
 *   latent  theta = [w[0..D), b?]         with independent Normal(loc, scale) priors
 *   eta_n   = sum_d X[n,d] * w[d] (+ b)
 *   y_n     ~ family(eta_n)               Bernoulli-logit (north star), Normal-identity, Poisson-log
 */
#ifndef EDHMC_H_
#define EDHMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDHMC_VERSION_MAJOR 0
#define EDHMC_VERSION_MINOR 1

typedef enum edhmc_status {
  EDHMC_OK = 0,
  EDHMC_ERR_INVALID = -1,    /* bad argument / unsupported configuration (reference: TypeError/ValueError) */
  EDHMC_ERR_CUDA = -2,       /* CUDA runtime error */
  EDHMC_ERR_NONFINITE = -3,  /* NaN/Inf in X or y (reference: ed.dot raises InvalidArgumentError,
                                edward/util/tensorflow.py:27-36) or in log-joint/gradient when debug */
  EDHMC_ERR_RANGE = -4,      /* update past the last Empirical row (reference: tf.scatter_update
                                out-of-range, edward/inferences/hmc.py:125) */
  EDHMC_ERR_STATE = -5,      /* call sequence error (e.g. run before bind_data) */
  EDHMC_ERR_COMM = -6,       /* NCCL failure */
  EDHMC_ERR_NOMEM = -7
} edhmc_status;

/* Likelihood families: log p(y_n | eta_n). */
typedef enum edhmc_family {
  EDHMC_BERNOULLI_LOGIT = 0, /* Bernoulli(logits=eta).log_prob(y): models/random_variables.py:13-25 →
                                TF bernoulli._log_prob → -sigmoid_cross_entropy_with_logits */
  EDHMC_NORMAL_IDENTITY = 1, /* Normal(loc=eta, scale=s).log_prob(y)   (tests/inferences/hmc_test.py:14-91) */
  EDHMC_POISSON_LOG = 2      /* Poisson(rate=exp(eta)).log_prob(y) */
} edhmc_family;

typedef enum edhmc_ydtype {
  EDHMC_Y_I32 = 0, /* reference dtype of Bernoulli data (inference.py:88-95 casts to key.dtype=int32) */
  EDHMC_Y_F32 = 1,
  EDHMC_Y_U8 = 2,
  EDHMC_Y_F64 = 3  /* float64 models only */
} edhmc_ydtype;

/* Floating-point type of the model: X, theta, params, gradients, momentum / uniform draws. The hot path and every
 * BASELINE configuration are float32. EDHMC_F64 serves the reference's float64 models (tests/inferences/hmc_test.py:93-97)
 * through a compact device path (csrc/f64.cuh) with its own typed entry points: edhmc_bind_data_f64, edhmc_logp_grad_f64,
 * edhmc_run_f64, edhmc_set_trace_f64 (y: EDHMC_Y_I32 or EDHMC_Y_F64). One chain, HMC only; the float32 entry points
 * refuse a float64 handle and vice versa. */
typedef enum edhmc_dtype {
  EDHMC_F32 = 0,
  EDHMC_F64 = 1
} edhmc_dtype;

/* Execution plan selector (all are CUDA paths; there is no CPU path). */
typedef enum edhmc_plan {
  EDHMC_PLAN_AUTO = 0,
  EDHMC_PLAN_PERSISTENT = 1, /* one cooperative launch runs every transition of edhmc_run */
  EDHMC_PLAN_STEPWISE = 2    /* one launch per data pass (+ NCCL all-reduce when sharded) */
} edhmc_plan;

typedef struct edhmc_cfg {
  int64_t n_rows;          /* rows of X / y held by THIS rank (its shard) */
  int64_t n_rows_global;   /* rows over all ranks (== n_rows when nranks == 1) */
  int32_t n_features;      /* D: columns of X */
  int64_t ldx;             /* row stride of X in floats (>= D) */
  int32_t has_bias;        /* 1: a scalar latent b is added to every linear predictor */
  int32_t family;          /* edhmc_family */
  int32_t y_dtype;         /* edhmc_ydtype */
  float   lik_scale;       /* Normal-identity only: likelihood scale s (> 0) */
  const float* prior_loc_host;   /* [P] P = D + has_bias; order [w..., b] */
  const float* prior_scale_host; /* [P] > 0 */
  int32_t device;          /* CUDA device ordinal */
  int32_t plan;            /* edhmc_plan */
  int32_t debug;           /* 1: check log-joint/gradient for NaN/Inf after every run (inference.py:279-281) */
  int32_t n_chains;        /* 0 or 1: one chain (the reference). C > 1: C vectorised chains (extension; any C — the kernels run whole
                              128-chain tiles, the padding chains are internal;
                              n_features + bias <= 64 on the one-kernel path, wider models on the two-GEMM path)
                              served by edhmc_run_chains / edhmc_logp_grad_chains */
  int32_t dtype;           /* edhmc_dtype */
  int32_t reserved[3];     /* must be zero */
} edhmc_cfg;

typedef struct edhmc_handle edhmc_t;

/* Library version as major*1000+minor. */
int edhmc_version(void);

/* Thread-local message for the most recent failure on this thread ("" if none). */
const char* edhmc_last_error(void);

/* Builds the sampler state for one model. Replaces the graph construction done once by
 * HMC.initialize → MonteCarlo.initialize → HMC.build_update (hmc.py:45-130, monte_carlo.py:95-109). */
int edhmc_create(edhmc_t** out, const edhmc_cfg* cfg);
int edhmc_destroy(edhmc_t* h);

/* Binds the observed data (borrowed, must stay valid while the handle uses it).
 * X: [n_rows, ldx] float32 row-major, 16-byte aligned. y: [n_rows] of cfg.y_dtype.
 * check_finite != 0 scans X and y once and fails with EDHMC_ERR_NONFINITE — this replaces the
 * per-evaluation CheckNumerics of ed.dot (util/tensorflow.py:33-36). Synchronises the stream when
 * check_finite != 0. */
int edhmc_bind_data(edhmc_t* h, const float* X, const void* y, int check_finite, void* stream);

/* log p(y, theta) and its gradient at theta, exactly what one `tf.gradients(log_joint(z), z)` of
 * hmc.py:199/206 evaluates via HMC._log_joint (hmc.py:161-192).
 * theta: [P] float32. logp: [1] float64 (device). grad: [P] float32 (device).
 * With row shards the result is the all-reduced global value on every rank. */
int edhmc_logp_grad(edhmc_t* h, const float* theta, double* logp, float* grad, void* stream);

/* Runs n_iter HMC transitions t = t0 .. t0+n_iter-1 with no host round-trip (n_iter == 1 is one
 * MonteCarlo.update, monte_carlo.py:111-150; n_iter == T is Inference.run's loop, inference.py:145-147).
 * Transition t reads row max(t-1,0) of `params` and writes row t (hmc.py:81-85,121-126).
 *   params    [T, ldp] float32, in place; first P entries of each row are theta = [w..., b]
 *   step_size, n_steps: leapfrog step size and count (hmc.py:45)
 *   r0        [n_iter, P] float32 momentum draws, or NULL → device Philox N(0,1)      (hmc.py:88-91)
 *   u         [n_iter]    float32 accept uniforms in (0,1), or NULL → device Philox   (hmc.py:108)
 * Accept rule: log(u) < K(r0) - K(rL) + logp(zL) - logp(z0), strict (hmc.py:100-109).
 * Fails with EDHMC_ERR_RANGE if t0 + n_iter > T. */
int edhmc_run(edhmc_t* h, float* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter,
              float step_size, int32_t n_steps, const float* r0, const float* u, void* stream);

/* Priors of constrained latents (reference: auto_transform, inference.py:223-264, hmc.py:132-159,
 * util/random_variables.py:856-917). kinds_host[P]: 0 = Normal(loc, scale) (default), 1 = a latent with support (0,1)
 * and a Beta(a = prior_loc, b = prior_scale) prior, sampled in the unconstrained space u = logit(z) with the
 * log-det-Jacobian of the sigmoid folded into the log joint: log p(u) = a log sigmoid(u) + b log sigmoid(-u) - lbeta(a,b).
 * The Empirical rows hold u, as the reference's do; the front-end maps them back with the bijector. A Bernoulli(probs=z)
 * likelihood of such a latent is the Bernoulli-logit family with logits u. NULL restores all-Normal. Single chain only. */
int edhmc_set_prior_kinds(edhmc_t* h, const int32_t* kinds_host);

/* Optional per-transition trace for parity tests. Both may be NULL (default).
 *   trace_scalars [n_iter, 8] float64: {logp_old, logp_new, K_old, K_new, ratio, log_u, accept, reserved}
 *   trace_pos     [n_iter, P] float32: proposed position z_L of each transition. */
int edhmc_set_trace(edhmc_t* h, double* trace_scalars, float* trace_pos);

/* float64 models (cfg.dtype == EDHMC_F64; the reference runs its HMC tests in both dtypes, tests/inferences/hmc_test.py:93-97,
 * with every tensor of hmc.py:61-130 a float64). Same semantics as the entry points above with X, theta, grad, params, the
 * injected draws, the step size and trace_pos in double. X: 8-byte aligned. */
int edhmc_bind_data_f64(edhmc_t* h, const double* X, const void* y, int check_finite, void* stream);
int edhmc_logp_grad_f64(edhmc_t* h, const double* theta, double* logp, double* grad, void* stream);
int edhmc_run_f64(edhmc_t* h, double* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter, double step_size,
                  int32_t n_steps, const double* r0, const double* u, void* stream);
int edhmc_set_trace_f64(edhmc_t* h, double* trace_scalars, double* trace_pos);

/* Counters owned by the handle: number of accepted proposals since the last reset
 * (MonteCarlo.n_accept, monte_carlo.py:100) and the cached log-joint of the current state.
 * Synchronises the stream. Any output pointer may be NULL. */
int edhmc_read_state(edhmc_t* h, int64_t* n_accept_host, double* logp_host, void* stream);

/* `sess.run(inference.reset)`: zero n_accept (monte_carlo.py:104) and drop the cached state. */
int edhmc_reset(edhmc_t* h, void* stream);

/* Seeds the device Philox4x32-10 streams used when r0 / u are NULL (ed.set_seed, util/graphs.py:59-73). */
int edhmc_seed(edhmc_t* h, uint64_t seed);

/* Row-sharded multi-GPU (extension; the reference is single-device). One handle per rank/GPU.
 * edhmc_comm_unique_id fills a 128-byte ncclUniqueId on rank 0; the caller broadcasts it (e.g. with
 * torch.distributed) and every rank calls edhmc_comm_init. Afterwards edhmc_logp_grad / edhmc_run
 * all-reduce (sum, float64) the per-shard [grad, logp] once per data pass over NVLink. */
int edhmc_comm_unique_id(void* id128_host);
int edhmc_comm_init(edhmc_t* h, const void* id128_host, int32_t nranks, int32_t rank);
/* The communicator (and the peer mapping below) belongs to the PROCESS, one per device: the first handle creates it,
 * later handles of the same (device, nranks, rank) attach to it by passing id128_host == NULL (edhmc_peer_attach:
 * handles_host == NULL) — ed.HMC builds one handle per inference object and must not pay ncclCommInitRank + cudaIpc
 * mapping each time. edhmc_comm_cached returns 0 (nothing cached), 1 (communicator) or 2 (communicator + peer
 * mapping); every rank must take the same branch. Handles that share the peer inboxes continue one exchange sequence:
 * sharded runs of one process must not overlap in time. edhmc_comm_release frees the cached state of a device. */
int edhmc_comm_cached(int32_t device, int32_t nranks, int32_t rank);
int edhmc_comm_release(int32_t device);

/* In-kernel all-reduce over peer memory (NVLink / NVSwitch), used by edhmc_run's persistent plan when rows are
 * sharded: edhmc_peer_export allocates this rank's inbox and returns its 64-byte cudaIpcMemHandle; the caller
 * all-gathers the handles (rank order) and every rank calls edhmc_peer_attach with the nranks*64-byte table.
 * Afterwards edhmc_run executes ALL transitions of a call in one cooperative launch per rank: after each data
 * pass CTA 0 stores the shard's [grad, logp] totals straight into every rank's inbox and every CTA sums the
 * nranks entries in rank order (bit-identical on all ranks) — no NCCL call and no kernel boundary per leapfrog
 * step. nranks <= 8 (one NVLink domain), one process per GPU. edhmc_cfg.plan = EDHMC_PLAN_STEPWISE keeps the
 * per-pass launch + ncclAllReduce path. A rank that waits longer than EDHMC_PEER_TIMEOUT_MS (default 20000) for a
 * peer gives up; edhmc_read_state then returns EDHMC_ERR_COMM.
 * Replaces: nothing in the reference (single device); SURVEY §8(e). */
int edhmc_peer_export(edhmc_t* h, void* handle64_host);
int edhmc_peer_attach(edhmc_t* h, const void* handles_host, int32_t nranks, int32_t rank);
int edhmc_peer_detach(edhmc_t* h); /* back to the ncclAllReduce plan (all ranks must agree) */

/* ---- Stochastic-gradient MCMC on the same log-joint gradient --------------------------------
 * n_iter iterations t = t0.. of SGLD (kind 0, edward/inferences/sgld.py:52-87) or SGHMC (kind 1, sghmc.py:58-96):
 * one gradient evaluation at row max(t-1,0) of `params`, then
 *   SGLD   lr = step_size/(t+1)^0.55;  row t = old + 0.5*lr*grad + sqrt(lr)*noise
 *   SGHMC  lr = 0.01*step_size;  row t = old + v;  v = (1-0.5*friction)*v + lr*grad + sqrt(lr*friction)*noise
 * grad = lik_factor * grad_lik + prior_factor[c] * grad_prior  (the `scale` argument of Inference.initialize,
 * sgld.py:106-119; prior_factor NULL = 1). batch_rows > 0: iteration t uses the rows of mini-batch (t mod
 * floor(n_rows/batch_rows)) of the bound data (multiple of 4). velocity [P] device (SGHMC, in/out). noise
 * [n_iter, P] or NULL → device Philox. n_accept += 1 per iteration (sgld.py:86). */
int edhmc_sgmcmc_run(edhmc_t* h, int32_t kind, float* params, int64_t ldp, int64_t T, int64_t t0, int64_t n_iter,
                     float step_size, float friction, float lik_factor, const float* prior_factor, float* velocity,
                     const float* noise, int64_t batch_rows, void* stream);

/* ---- C vectorised chains (extension; the reference runs one chain per ed.HMC object) -------------------------
 * Every chain is an independent HMC chain on the same data: per leapfrog step the C gradients are one dense
 * contraction S = X·W, G = Xᵀ·(y − σ(S)) executed on the tcgen05 tensor cores in 3xTF32 (TMEM accumulators).
 *   params [T, C, P] float32 in place; r0 [n_iter, C, P] / u [n_iter, C] optional injected draws;
 *   theta [C, P]; logp [C] float64; grad [C, P] float32; trace [n_iter, C, 8] float64 (same columns as above).
 * Each chain follows exactly the single-chain semantics of edhmc_run (checked against independent runs). */
int edhmc_run_chains(edhmc_t* h, float* params, int64_t T, int64_t t0, int64_t n_iter, float step_size,
                     int32_t n_steps, const float* r0, const float* u, void* stream);
int edhmc_logp_grad_chains(edhmc_t* h, const float* theta, double* logp, float* grad, void* stream);
int edhmc_read_chain_state(edhmc_t* h, int64_t* n_accept_host /*[C]*/, double* logp_host /*[C]*/, void* stream);
int edhmc_set_chain_trace(edhmc_t* h, double* trace);
/* Development aid: [64][16] int64 per-role clock64 timeline of one CTA of the pipelined tensor-core pass. */
int edhmc_set_chain_debug(edhmc_t* h, long long* buf);

/* Development aid for the persistent plan of edhmc_run: thread 0 of every CTA stamps its first `n_passes` data passes
 * into buf [n_passes][grid_ctas][32] int64 (device): clock64 at {pass start, CTA sums ready, partials published, grid
 * barrier passed, totals ready (after the peer exchange when sharded), integrator done}, then %globaltimer (ns) at
 * {pass start, grid barrier passed}, then the cycles warps 0..7 spent waiting for their tiles to land (one-ring-per-CTA
 * plans), then (slots 16..) stamps of the leader CTA under the leader protocol (slots 3..5 are then stamped by CTA 0
 * only). NULL switches it off (default). bench.py reports the medians as `timeline`.
 * Replaces: nothing in the reference (introspection). */
int edhmc_set_timeline(edhmc_t* h, long long* buf, int32_t n_passes);

/* Posterior-predictive evaluation over the device-resident sample store: what ed.evaluate
 * (criticisms/evaluate.py:20-235) and ed.ppc compute for the GLMs of this path. For every row n of X and S draws
 * s = 0..S-1 of the latents — w_s = params[idx_w[s], 0..D), b_s = params[idx_b[s], bias_col] (independent index lists:
 * the reference draws every latent independently, empirical.py:98-110; idx_b NULL = no bias) — eta = x_n . w_s + b_s and
 *   mean_out[n]   = mean_s E[y | eta]: sigmoid(eta) (Bernoulli, evaluate.py:132-143), eta (Normal), exp(eta) (Poisson);
 *   loglik_out[n] = sum_s log p(y_n | eta) (float64; evaluate.py:222-227 averages it), or NULL.
 * One fused pass: X is read once, eta [N, S] is never materialised; sums are accumulated in a fixed order.
 * All pointers are device pointers. Stateless (no handle). */
int edhmc_predictive(const float* X, int64_t n_rows, int64_t ldx, int32_t n_features, const void* y, int32_t y_dtype,
                     int32_t family, float lik_scale, const float* params, int64_t ldp, const int32_t* idx_w,
                     const int32_t* idx_b, int32_t bias_col, int32_t n_draws, float* mean_out, double* loglik_out,
                     int32_t device, void* stream);

/* Read-bandwidth probe for the roofline denominators (bench.py): queues `iters` read sweeps over buf[0..bytes) on
 * `stream`; the caller times them with CUDA events. mode 0: LDG.128 grid-stride loads; mode 1: 1-D TMA bulk copies into
 * a shared-memory ring (the access path of the sampler's data pass). A buffer that fits the L2 gives the L2 read
 * throughput, one much larger than L2 the read-only HBM throughput. buf 16-byte aligned, bytes >= 1 MiB (mode 1 reads
 * whole 32 KiB chunks), sink: 4 writable device bytes. Replaces: nothing in the reference (measurement aid). */
int edhmc_probe_read(const void* buf, int64_t bytes, int32_t iters, int32_t mode, void* sink, void* stream);

/* Introspection for benches/tests: fills up to `cap` int64 values:
 * {grid_ctas, warps_per_cta, ring_stages, tile_rows, lanes_per_row, vec_width, smem_bytes,
 *  plan_in_use, passes_last_run, launches_last_run, ring_mode (0: one TMA ring per warp, 1: one ring per CTA, 2: re-laid
 *  32-row tiles read with LDG), and for ring mode 2 the 32-row tiles per CTA that a persistent launch keeps resident in
 *  shared memory and in tensor memory}.
 * Returns the number written. */
int edhmc_plan_info(edhmc_t* h, int64_t* out_host, int32_t cap);

/* Host-only: the tiling edhmc_create would choose for the wide / row-sharded vectorised-chain path (two-GEMM
 * plan, csrc/chains_wide.cu) on a device with num_sms multiprocessors. out[0..7] = {chain tiles, K of GEMM 1
 * (features rounded to 16), row tiles of 128, feature tiles of GEMM 2, feature-tile width, CTAs per chain tile in
 * GEMM 1, K splits of GEMM 2, padded feature count of GEMM 2}. No CUDA call is made (usable without a GPU).
 * Replaces: nothing in the reference (introspection for tests and capacity planning). */
int edhmc_chains_plan_probe(int64_t n_rows, int32_t n_features, int32_t n_chains, int32_t num_sms, int64_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* EDHMC_H_ */
