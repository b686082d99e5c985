/*
 * hmc_ref.c — C restatement of the reference's CPU execution of the HMC hot path (Bernoulli-logit /
 * Normal-identity likelihood, Normal priors), used as the CPU baseline and as a second oracle.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing under edward_b200/ links or loads this file.
 * PARITY UNPINNED: TensorFlow (the reference's arithmetic backend) is not available here, see
 * oracle/hmc_oracle.py for what pins this restatement instead.
 *
 * It executes the reference's SCHEDULE, not ours: per transition (edward/inferences/hmc.py:81-130)
 *   L+1 gradient evaluations   (hmc.py:199,206 — tf.gradients of a fresh log-joint copy each time)
 *   2 forward evaluations      (hmc.py:104-105 — log joint at the new and at the old state)
 * and every evaluation runs the un-fused TF op sequence over full [N] temporaries:
 *   CheckNumerics(X), CheckNumerics(w)          util/tensorflow.py:33-36
 *   MatMul(X, w)                                util/tensorflow.py:45
 *   GreaterEqual/Select/Neg/Exp/Log1p/Mul/Sub/Add, Sum      Bernoulli._log_prob [TF 1.5]
 *   (backward) elementwise gradients, MatMul(X^T, r)        tf.gradients
 * float32 throughout, reductions in float32 (per-thread partial sums combined in thread order; Eigen's
 * exact reduction tree is not reproducible outside TF). OpenMP threads stand in for TF's intra-op pool.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  const float* X;
  const int32_t* y_i32; /* Bernoulli data (int32, cast to float inside log_prob) */
  const float* y_f32;   /* Normal data */
  int64_t N;
  int D;
  int has_bias;
  int family; /* 0 Bernoulli-logit, 1 Normal-identity */
  float lik_scale;
  const float* prior_loc;   /* [P] */
  const float* prior_scale; /* [P] */
  int check_numerics;
  /* temporaries [N] */
  float *logits, *t1, *t2, *resid;
} ref_model;

int hmc_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int all_finite(const float* a, int64_t n) {
  int ok = 1;
#pragma omp parallel for reduction(&& : ok) schedule(static)
  for (int64_t i = 0; i < n; ++i) ok = ok && isfinite(a[i]);
  return ok;
}

/* ed.dot: logits[n] = sum_d X[n,d]*w[d]  (+ b) */
static int op_dot(ref_model* m, const float* theta) {
  if (m->check_numerics) {
    if (!all_finite(m->X, m->N * (int64_t)m->D)) return -3;
    if (!all_finite(theta, m->D)) return -3;
  }
  const int D = m->D;
  const float b = m->has_bias ? theta[D] : 0.0f;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < m->N; ++n) {
    const float* x = m->X + n * D;
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc += x[d] * theta[d];
    m->logits[n] = m->has_bias ? acc + b : acc;
  }
  return 0;
}

static float sum_f32(const float* a, int64_t n) {
  float total = 0.0f;
#pragma omp parallel
  {
    float local = 0.0f;
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; ++i) local += a[i];
#pragma omp critical
    total += local;
  }
  return total;
}

/* log-likelihood terms into t1[N] (separate passes like separate TF ops), returns their float32 sum */
static float op_loglik(ref_model* m) {
  const int64_t N = m->N;
  if (m->family == 0) {
    /* relu = where(l>=0,l,0); neg_abs = where(l>=0,-l,l)  (two Select ops) */
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const float l = m->logits[n];
      m->t1[n] = l >= 0.0f ? l : 0.0f;
      m->t2[n] = l >= 0.0f ? -l : l;
    }
    /* log1p(exp(neg_abs)) (Exp, Log1p) */
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) m->t2[n] = log1pf(expf(m->t2[n]));
    /* relu - l*y (Cast, Mul, Sub) ; add ; neg */
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const float yv = (float)m->y_i32[n];
      m->t1[n] = -((m->t1[n] - m->logits[n] * yv) + m->t2[n]);
    }
  } else {
    const float s = m->lik_scale;
    const float lognorm = 0.5f * logf(2.0f * (float)M_PI) + logf(s);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const float z = (m->y_f32[n] - m->logits[n]) / s;
      m->t1[n] = -0.5f * (z * z) - lognorm;
    }
  }
  return sum_f32(m->t1, N);
}

static float prior_logp(const ref_model* m, const float* theta, int lo, int hi) {
  float acc = 0.0f;
  for (int c = lo; c < hi; ++c) {
    const float z = (theta[c] - m->prior_loc[c]) / m->prior_scale[c];
    acc += -0.5f * (z * z) - (0.5f * logf(2.0f * (float)M_PI) + logf(m->prior_scale[c]));
  }
  return acc;
}

/* HMC._log_joint, hmc.py:161-192: 0.0 + prior(w) [+ prior(b)] + likelihood, float32 */
static int log_joint(ref_model* m, const float* theta, float* out) {
  int rc = op_dot(m, theta);
  if (rc) return rc;
  float lj = 0.0f;
  lj += prior_logp(m, theta, 0, m->D);
  if (m->has_bias) lj += prior_logp(m, theta, m->D, m->D + 1);
  lj += op_loglik(m);
  *out = lj;
  return 0;
}

/* tf.gradients(log_joint(theta), theta): forward pass + backward pass */
static int grad_log_joint(ref_model* m, const float* theta, float* grad) {
  float lj;
  int rc = log_joint(m, theta, &lj); /* the forward ops are part of the gradient sub-graph */
  if (rc) return rc;
  const int64_t N = m->N;
  const int D = m->D;
  if (m->family == 0) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
      const float l = m->logits[n];
      const float yv = (float)m->y_i32[n];
      const float e = expf(l >= 0.0f ? -l : l);
      const float q = (1.0f / (1.0f + e)) * e;
      m->resid[n] = l >= 0.0f ? (yv - 1.0f) + q : yv - q;
    }
  } else {
    const float s = m->lik_scale;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) m->resid[n] = ((m->y_f32[n] - m->logits[n]) / s) / s;
  }
  /* MatMul(X^T, resid) */
  for (int d = 0; d < D; ++d) grad[d] = 0.0f;
#pragma omp parallel
  {
    float* local = (float*)calloc((size_t)D, sizeof(float));
#pragma omp for schedule(static) nowait
    for (int64_t n = 0; n < N; ++n) {
      const float* x = m->X + n * D;
      const float rv = m->resid[n];
      for (int d = 0; d < D; ++d) local[d] += rv * x[d];
    }
#pragma omp critical
    for (int d = 0; d < D; ++d) grad[d] += local[d];
    free(local);
  }
  if (m->has_bias) grad[D] = sum_f32(m->resid, N);
  const int P = D + (m->has_bias ? 1 : 0);
  for (int c = 0; c < P; ++c) {
    const float z = (theta[c] - m->prior_loc[c]) / m->prior_scale[c];
    grad[c] += (-0.5f * (2.0f * z)) / m->prior_scale[c];
  }
  return 0;
}

/*
 * Runs n_iter transitions t0..t0+n_iter-1 in place on params[T,P] with injected momentum r0[n_iter,P]
 * and uniforms u[n_iter]. trace (nullable) [n_iter,8] = {logp_old, logp_new, k_old, k_new, ratio, log_u,
 * accept, 0}. Returns 0, -3 on NaN/Inf operands (CheckNumerics), -4 on out-of-range t, -7 on OOM.
 */
int hmc_ref_run(const float* X, const void* y, int64_t N, int D, int has_bias, int family, float lik_scale,
                const float* prior_loc, const float* prior_scale, float* params, int64_t T, int64_t t0,
                int64_t n_iter, float step_size, int L, const float* r0, const float* u, int check_numerics,
                int64_t* n_accept, double* trace) {
  if (t0 < 0 || t0 + n_iter > T) return -4;
  const int P = D + (has_bias ? 1 : 0);
  ref_model m;
  memset(&m, 0, sizeof(m));
  m.X = X;
  m.y_i32 = (const int32_t*)y;
  m.y_f32 = (const float*)y;
  m.N = N;
  m.D = D;
  m.has_bias = has_bias;
  m.family = family;
  m.lik_scale = lik_scale;
  m.prior_loc = prior_loc;
  m.prior_scale = prior_scale;
  m.check_numerics = check_numerics;
  m.logits = (float*)malloc((size_t)N * sizeof(float));
  m.t1 = (float*)malloc((size_t)N * sizeof(float));
  m.t2 = (float*)malloc((size_t)N * sizeof(float));
  m.resid = (float*)malloc((size_t)N * sizeof(float));
  float* z = (float*)malloc((size_t)P * 4 * sizeof(float));
  if (!m.logits || !m.t1 || !m.t2 || !m.resid || !z) return -7;
  float *r = z + P, *g = r + P, *old = g + P;
  const float eps = step_size, half_eps = 0.5f * step_size;
  int rc = 0;
  int64_t acc = 0;
  for (int64_t it = 0; it < n_iter && !rc; ++it) {
    const int64_t t = t0 + it;
    const int64_t tp = t > 0 ? t - 1 : 0;
    memcpy(old, params + tp * P, (size_t)P * sizeof(float));
    memcpy(z, old, (size_t)P * sizeof(float));
    memcpy(r, r0 + it * P, (size_t)P * sizeof(float));
    float k_old = 0.0f;
    for (int c = 0; c < P; ++c) k_old += r[c] * r[c];
    k_old *= 0.5f;
    /* leapfrog, hmc.py:195-210 */
    if ((rc = grad_log_joint(&m, z, g))) break;
    for (int s = 0; s < L; ++s) {
      for (int c = 0; c < P; ++c) {
        r[c] = r[c] + half_eps * g[c];
        z[c] = z[c] + eps * r[c];
      }
      if ((rc = grad_log_joint(&m, z, g))) break;
      for (int c = 0; c < P; ++c) r[c] = r[c] + half_eps * g[c];
    }
    if (rc) break;
    float k_new = 0.0f;
    for (int c = 0; c < P; ++c) k_new += r[c] * r[c];
    k_new *= 0.5f;
    float lp_new, lp_old;
    if ((rc = log_joint(&m, z, &lp_new))) break;
    if ((rc = log_joint(&m, old, &lp_old))) break;
    float ratio = k_old;
    ratio -= k_new;
    ratio += lp_new;
    ratio -= lp_old;
    const float log_u = logf(u[it]);
    const int accept = log_u < ratio;
    memcpy(params + t * P, accept ? z : old, (size_t)P * sizeof(float));
    acc += accept;
    if (trace) {
      double* tr = trace + it * 8;
      tr[0] = lp_old;
      tr[1] = lp_new;
      tr[2] = k_old;
      tr[3] = k_new;
      tr[4] = ratio;
      tr[5] = log_u;
      tr[6] = accept;
      tr[7] = 0.0;
    }
  }
  if (n_accept) *n_accept = acc;
  free(m.logits);
  free(m.t1);
  free(m.t2);
  free(m.resid);
  free(z);
  return rc;
}

/* One fused-schedule-free evaluation for unit tests: log joint (float32) and gradient at theta. */
int hmc_ref_logp_grad(const float* X, const void* y, int64_t N, int D, int has_bias, int family, float lik_scale,
                      const float* prior_loc, const float* prior_scale, const float* theta, float* logp,
                      float* grad) {
  ref_model m;
  memset(&m, 0, sizeof(m));
  m.X = X;
  m.y_i32 = (const int32_t*)y;
  m.y_f32 = (const float*)y;
  m.N = N;
  m.D = D;
  m.has_bias = has_bias;
  m.family = family;
  m.lik_scale = lik_scale;
  m.prior_loc = prior_loc;
  m.prior_scale = prior_scale;
  m.check_numerics = 1;
  m.logits = (float*)malloc((size_t)N * sizeof(float));
  m.t1 = (float*)malloc((size_t)N * sizeof(float));
  m.t2 = (float*)malloc((size_t)N * sizeof(float));
  m.resid = (float*)malloc((size_t)N * sizeof(float));
  if (!m.logits || !m.t1 || !m.t2 || !m.resid) return -7;
  int rc = log_joint(&m, theta, logp);
  if (!rc) rc = grad_log_joint(&m, theta, grad);
  free(m.logits);
  free(m.t1);
  free(m.t2);
  free(m.resid);
  return rc;
}
