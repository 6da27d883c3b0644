/*
 * ba_oracle.c — CPU ORACLE for HP2 (global bundle adjustment).  TEST INFRASTRUCTURE.
 *
 * Restates, on the flattened problem of include/psfm_b200.h:
 *   - problem assembly of BundleAdjuster::SetUp / AddImageToProblem / ParameterizeCameras
 *     (reference sfm/gmapper/src/optim/bundle_adjustment.cc:326-447, 500-544);
 *   - linear-solver selection rule (bundle_adjustment.cc:276-286);
 *   - COLMAP bd84ad6 BundleAdjustmentCostFunction<SimplePinholeCameraModel> and the
 *     constant-pose variant (colmap/base/cost_functions.h — external, call sites
 *     bundle_adjustment.cc:380-411): ceres::UnitQuaternionRotatePoint, perspective
 *     division, x = f u + cx;
 *   - Ceres 2.0.0: loss functions + Corrector, QuaternionParameterization,
 *     SubsetParameterization, TrustRegionMinimizer, LevenbergMarquardtStrategy,
 *     SchurEliminator + dense Cholesky (DENSE_/SPARSE_SCHUR give the same exact step),
 *     ImplicitSchurComplement + ConjugateGradientsSolver + SchurJacobiPreconditioner
 *     (ITERATIVE_SCHUR) — see SURVEY.md Appendix A and oracle/ceres_semantics.h.
 *
 * PARITY UNPINNED (no reference golden vectors exist; see psfm_oracle.h).
 * Derivatives are analytic; the quaternion block follows Ceres' two-stage form
 * (ambient 2x4 Jacobian of the UnitQuaternionRotatePoint polynomial times the 4x3
 * plus-Jacobian) on purpose: the CUDA path uses the closed form -2[RX]x, so the two
 * derivations check each other.
 *
 * Also serves as bench.py's timed CPU baseline ("port"), hence OpenMP.
 */
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "psfm_oracle.h"
#include "ceres_semantics.h"

#define MAXT 256

static double wall_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* PSFM_ORACLE_TIMING=1: wall time per phase of a solve, printed to stderr (where does the CPU
   baseline spend its time) */
static double g_tm[8];
static int g_tm_on = -1;
#define TM_BEGIN() const double tm__0 = (g_tm_on > 0) ? wall_s() : 0.0
#define TM_END(slot) do { if (g_tm_on > 0) g_tm[slot] += wall_s() - tm__0; } while (0)

#include "ba_schur_blocks.h"

int psfm_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

typedef struct {
  int F, P, M, C, NS;
  const psfm_ba_problem* pb;
  psfm_ba_options o;
  int nthreads;
  /* structure */
  int* pt_ptr;   /* P+1 */
  int* pt_obs;   /* M: observation ids grouped by point (identity: see spb) */
  /* Internal observation order = grouped by point, like Ceres after
     ReorderProgramForSchurTypeLinearSolver (rows of an e-block are contiguous): `pb` points at
     `spb`, a copy of the caller's problem whose observation arrays are the sorted ones, so every
     pass over the stored Jacobian streams through memory. */
  psfm_ba_problem spb;
  int32_t* s_img; int32_t* s_pt; double* s_xy;
  int* orig;     /* M: sorted position -> caller's observation index */
  unsigned char* slot_active; /* NS */
  unsigned char* img_has_obs; /* F */
  unsigned char* cam_has_obs; /* C */
  unsigned char* pose_var;    /* F */
  /* state */
  double *q, *t, *X, *K;
  double *qc, *tc, *Xc, *Kc;
  /* linearisation (loss-corrected; column-scaled after scale_columns()) */
  double* r;  /* 2M */
  double* Jc; /* 12M */
  double* Jp; /* 6M */
  double* Jk; /* 6M */
  double* scale_c; /* NS */
  double* scale_p; /* 3P */
  double* g_c;     /* NS  unscaled tangent gradient */
  double* g_p;     /* 3P */
  /* linear solve workspace */
  double* diag_c; double* diag_p; /* clamped squared column norms of scaled J */
  double* D_c; double* D_p;       /* lm diagonal */
  double* Hinv;   /* 9P  (E'E + D^2)^-1 */
  double* gs_c;   /* NS  scaled gradient J_s' r */
  double* gs_p;   /* 3P */
  double* step_c; double* step_p;
  double* S;      /* NS*NS (exact) */
  double* Spriv;  /* private copies of S for the assembly threads */
  int nspriv;
  double* tbuf;   /* per-thread scratch */
  size_t tbuf_stride;
  int num_linear_iterations;
  sblocks_t sb;   /* block-sparse reduced system (single shared camera): ba_schur_blocks.h */
  int sb_unusable;
} ctx_t;

/* ---------------------------------------------------------------- small math */

static void quat_mul(const double* a, const double* b, double* o) {
  /* Hamilton product, wxyz (ceres::QuaternionProduct) */
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

/* ceres::QuaternionParameterization::Plus */
static void quat_plus(const double* x, const double* d, double* o) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd;
    double qd[4] = {cos(nd), s * d[0], s * d[1], s * d[2]};
    quat_mul(qd, x, o);
  } else {
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
  }
}

/* ceres LossFunction::Evaluate: rho[0..2] for squared norm s */
static void loss_eval(int type, double a, double s, double* rho) {
  if (type == PSFM_LOSS_SOFT_L1) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double tmp = sqrt(sum);
    rho[0] = 2.0 * b * (tmp - 1.0);
    rho[1] = fmax(DBL_MIN, 1.0 / tmp);
    rho[2] = -(c * rho[1]) / (2.0 * sum);
  } else if (type == PSFM_LOSS_CAUCHY) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * log(sum);
    rho[1] = fmax(DBL_MIN, inv);
    rho[2] = -c * (inv * inv);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

/* 3x3 SPD inverse through Cholesky (Ceres InvertPSDMatrix -> LLT solve of identity) */
static int inv3_spd(const double* A, double* Ai) {
  double l00, l10, l11, l20, l21, l22;
  if (!(A[0] > 0.0)) return 1;
  l00 = sqrt(A[0]);
  l10 = A[3] / l00;
  l20 = A[6] / l00;
  double d = A[4] - l10 * l10;
  if (!(d > 0.0)) return 1;
  l11 = sqrt(d);
  l21 = (A[7] - l20 * l10) / l11;
  d = A[8] - l20 * l20 - l21 * l21;
  if (!(d > 0.0)) return 1;
  l22 = sqrt(d);
  /* inverse of L */
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  /* A^-1 = Li' Li */
  Ai[0] = i00 * i00 + i10 * i10 + i20 * i20;
  Ai[1] = i10 * i11 + i20 * i21;
  Ai[2] = i20 * i22;
  Ai[3] = Ai[1];
  Ai[4] = i11 * i11 + i21 * i21;
  Ai[5] = i21 * i22;
  Ai[6] = Ai[2];
  Ai[7] = Ai[5];
  Ai[8] = i22 * i22;
  return 0;
}

/* general small SPD inverse restricted to active dims (n<=3), identity elsewhere */
static void inv_small_masked(const double* A, const unsigned char* act, int n, double* Ai) {
  int idx[3], m = 0;
  for (int i = 0; i < n; ++i) if (act[i]) idx[m++] = i;
  for (int i = 0; i < n * n; ++i) Ai[i] = 0.0;
  for (int i = 0; i < n; ++i) if (!act[i]) Ai[i * n + i] = 1.0;
  if (m == 0) return;
  double B[9], Bi[9];
  if (m == 3) {
    memcpy(B, A, sizeof(B));
    if (inv3_spd(B, Bi)) { for (int i = 0; i < 9; ++i) Bi[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    memcpy(Ai, Bi, sizeof(Bi));
    return;
  }
  if (m == 1) { Ai[idx[0] * n + idx[0]] = 1.0 / A[idx[0] * n + idx[0]]; return; }
  /* m == 2 */
  {
    const double a = A[idx[0] * n + idx[0]], b = A[idx[0] * n + idx[1]], d = A[idx[1] * n + idx[1]];
    const double l00 = sqrt(a), l10 = b / l00, l11 = sqrt(d - l10 * l10);
    const double i00 = 1.0 / l00, i11 = 1.0 / l11, i10 = -l10 * i00 * i11;
    Ai[idx[0] * n + idx[0]] = i00 * i00 + i10 * i10;
    Ai[idx[0] * n + idx[1]] = i10 * i11;
    Ai[idx[1] * n + idx[0]] = i10 * i11;
    Ai[idx[1] * n + idx[1]] = i11 * i11;
  }
}

/* ---------------------------------------------------------------- residual block */

/* One observation: residual (2), and if J != NULL the ambient->tangent Jacobians
   jc[2x6] (rot3 | t3), jp[2x3], jk[2x3]; all loss-corrected; returns 1/2 rho(s). */
static double eval_obs(const ctx_t* c, int i, const double* q, const double* t, const double* X,
                       const double* K, double* r_out, double* jc, double* jp, double* jk) {
  const psfm_ba_problem* pb = c->pb;
  const int img = pb->obs_image[i], pt = pb->obs_point[i], cam = pb->image_camera[img];
  const double* qv = q + 4 * img;
  const double* tv = t + 3 * img;
  const double* x = X + 3 * pt;
  const double f = K[3 * cam], cx = K[3 * cam + 1], cy = K[3 * cam + 2];
  const double q0 = qv[0], q1 = qv[1], q2 = qv[2], q3 = qv[3];
  /* ceres::UnitQuaternionRotatePoint */
  const double t2 = q0 * q1, t3 = q0 * q2, t4 = q0 * q3, t5 = -q1 * q1, t6 = q1 * q2,
               t7 = q1 * q3, t8 = -q2 * q2, t9 = q2 * q3, t1 = -q3 * q3;
  double p[3];
  p[0] = 2.0 * ((t8 + t1) * x[0] + (t6 - t4) * x[1] + (t3 + t7) * x[2]) + x[0];
  p[1] = 2.0 * ((t4 + t6) * x[0] + (t5 + t1) * x[1] + (t9 - t2) * x[2]) + x[1];
  p[2] = 2.0 * ((t7 - t3) * x[0] + (t2 + t9) * x[1] + (t5 + t8) * x[2]) + x[2];
  p[0] += tv[0]; p[1] += tv[1]; p[2] += tv[2];
  const double u = p[0] / p[2], v = p[1] / p[2];
  double r[2];
  r[0] = f * u + cx - pb->obs_xy[2 * i];
  r[1] = f * v + cy - pb->obs_xy[2 * i + 1];
  const double s = r[0] * r[0] + r[1] * r[1];
  double rho[3];
  loss_eval(c->o.loss_function_type, c->o.loss_function_scale, s, rho);
  /* Corrector: rho[2] <= 0 for all three losses -> first-order branch */
  const double sq = sqrt(rho[1]);
  r_out[0] = sq * r[0];
  r_out[1] = sq * r[1];
  if (jc) {
    const double iz = 1.0 / p[2];
    /* d r / d p */
    const double a00 = f * iz, a02 = -f * u * iz, a11 = f * iz, a12 = -f * v * iz;
    /* d p / d X  (the polynomial's matrix; equals R(q) for unit q) */
    double R[9];
    R[0] = 1.0 + 2.0 * (t8 + t1); R[1] = 2.0 * (t6 - t4);       R[2] = 2.0 * (t3 + t7);
    R[3] = 2.0 * (t4 + t6);       R[4] = 1.0 + 2.0 * (t5 + t1); R[5] = 2.0 * (t9 - t2);
    R[6] = 2.0 * (t7 - t3);       R[7] = 2.0 * (t2 + t9);       R[8] = 1.0 + 2.0 * (t5 + t8);
    for (int k = 0; k < 3; ++k) {
      jp[k] = sq * (a00 * R[k] + a02 * R[6 + k]);
      jp[3 + k] = sq * (a11 * R[3 + k] + a12 * R[6 + k]);
    }
    /* d p / d q (ambient, 3x4) of the polynomial */
    double dq[12];
    dq[0] = 2.0 * (-q3 * x[1] + q2 * x[2]);
    dq[1] = 2.0 * (q2 * x[1] + q3 * x[2]);
    dq[2] = 2.0 * (-2.0 * q2 * x[0] + q1 * x[1] + q0 * x[2]);
    dq[3] = 2.0 * (-2.0 * q3 * x[0] - q0 * x[1] + q1 * x[2]);
    dq[4] = 2.0 * (q3 * x[0] - q1 * x[2]);
    dq[5] = 2.0 * (q2 * x[0] - 2.0 * q1 * x[1] - q0 * x[2]);
    dq[6] = 2.0 * (q1 * x[0] + q3 * x[2]);
    dq[7] = 2.0 * (q0 * x[0] - 2.0 * q3 * x[1] + q2 * x[2]);
    dq[8] = 2.0 * (-q2 * x[0] + q1 * x[1]);
    dq[9] = 2.0 * (q3 * x[0] + q0 * x[1] - 2.0 * q1 * x[2]);
    dq[10] = 2.0 * (-q0 * x[0] + q3 * x[1] - 2.0 * q2 * x[2]);
    dq[11] = 2.0 * (q1 * x[0] + q2 * x[1]);
    /* d r / d q (2x4) */
    double rq[8];
    for (int k = 0; k < 4; ++k) {
      rq[k] = a00 * dq[k] + a02 * dq[8 + k];
      rq[4 + k] = a11 * dq[4 + k] + a12 * dq[8 + k];
    }
    /* QuaternionParameterization::ComputeJacobian (4x3, row-major) */
    const double PJ[12] = {-q1, -q2, -q3, q0, q3, -q2, -q3, q0, q1, q2, -q1, q0};
    const unsigned char* act = c->slot_active + 6 * img;
    for (int row = 0; row < 2; ++row) {
      for (int k = 0; k < 3; ++k) {
        double acc = 0.0;
        for (int a = 0; a < 4; ++a) acc += rq[4 * row + a] * PJ[3 * a + k];
        jc[6 * row + k] = act[k] ? sq * acc : 0.0;
      }
    }
    jc[3] = act[3] ? sq * a00 : 0.0; jc[4] = 0.0;                    jc[5] = act[5] ? sq * a02 : 0.0;
    jc[9] = 0.0;                    jc[10] = act[4] ? sq * a11 : 0.0; jc[11] = act[5] ? sq * a12 : 0.0;
    const unsigned char* kact = c->slot_active + 6 * c->F + 3 * cam;
    jk[0] = kact[0] ? sq * u : 0.0; jk[1] = kact[1] ? sq : 0.0; jk[2] = 0.0;
    jk[3] = kact[0] ? sq * v : 0.0; jk[4] = 0.0;                jk[5] = kact[2] ? sq : 0.0;
  }
  return 0.5 * rho[0];
}

/* cost (and optionally r, J) at (q,t,X,K) */
static double evaluate(ctx_t* c, const double* q, const double* t, const double* X,
                       const double* K, int with_jac) {
  double cost = 0.0;
  const int M = c->M;
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(c->nthreads)
  for (int i = 0; i < M; ++i) {
    double rr[2];
    if (with_jac) {
      cost += eval_obs(c, i, q, t, X, K, c->r + 2 * i, c->Jc + 12 * (size_t)i,
                       c->Jp + 6 * (size_t)i, c->Jk + 6 * (size_t)i);
    } else {
      cost += eval_obs(c, i, q, t, X, K, rr, NULL, NULL, NULL);
    }
  }
  return cost;
}

/* g = J' r over the stored Jacobian; out_c [NS], out_p [3P] */
static void jt_r(ctx_t* c, double* out_c, double* out_p) {
  const int NS = c->NS, P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  memset(c->tbuf, 0, sizeof(double) * c->tbuf_stride * c->nthreads);
#pragma omp parallel num_threads(c->nthreads)
  {
#ifdef _OPENMP
    double* buf = c->tbuf + c->tbuf_stride * omp_get_thread_num();
#else
    double* buf = c->tbuf;
#endif
#pragma omp for schedule(static)
    for (int p = 0; p < P; ++p) {
      double gp[3] = {0, 0, 0};
      for (int e = c->pt_ptr[p]; e < c->pt_ptr[p + 1]; ++e) {
        const int i = c->pt_obs[e];
        const double* r = c->r + 2 * i;
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        for (int k = 0; k < 6; ++k) buf[6 * img + k] += jc[k] * r[0] + jc[6 + k] * r[1];
        for (int k = 0; k < 3; ++k) buf[6 * F + 3 * cam + k] += jk[k] * r[0] + jk[3 + k] * r[1];
        for (int k = 0; k < 3; ++k) gp[k] += jp[k] * r[0] + jp[3 + k] * r[1];
      }
      out_p[3 * p] = gp[0]; out_p[3 * p + 1] = gp[1]; out_p[3 * p + 2] = gp[2];
    }
  }
  for (int k = 0; k < NS; ++k) {
    double s = 0.0;
    for (int th = 0; th < c->nthreads; ++th) s += c->tbuf[c->tbuf_stride * th + k];
    out_c[k] = s;
  }
}

/* squared column norms of the stored Jacobian */
static void col_sqnorm(ctx_t* c, double* out_c, double* out_p) {
  const int NS = c->NS, P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  memset(c->tbuf, 0, sizeof(double) * c->tbuf_stride * c->nthreads);
#pragma omp parallel num_threads(c->nthreads)
  {
#ifdef _OPENMP
    double* buf = c->tbuf + c->tbuf_stride * omp_get_thread_num();
#else
    double* buf = c->tbuf;
#endif
#pragma omp for schedule(static)
    for (int p = 0; p < P; ++p) {
      double dp[3] = {0, 0, 0};
      for (int e = c->pt_ptr[p]; e < c->pt_ptr[p + 1]; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        for (int k = 0; k < 6; ++k) buf[6 * img + k] += jc[k] * jc[k] + jc[6 + k] * jc[6 + k];
        for (int k = 0; k < 3; ++k) buf[6 * F + 3 * cam + k] += jk[k] * jk[k] + jk[3 + k] * jk[3 + k];
        for (int k = 0; k < 3; ++k) dp[k] += jp[k] * jp[k] + jp[3 + k] * jp[3 + k];
      }
      out_p[3 * p] = dp[0]; out_p[3 * p + 1] = dp[1]; out_p[3 * p + 2] = dp[2];
    }
  }
  for (int k = 0; k < NS; ++k) {
    double s = 0.0;
    for (int th = 0; th < c->nthreads; ++th) s += c->tbuf[c->tbuf_stride * th + k];
    out_c[k] = s;
  }
}

/* jacobian_->ScaleColumns(jacobian_scaling_) */
static void scale_columns(ctx_t* c) {
  const int M = c->M, F = c->F;
  const psfm_ba_problem* pb = c->pb;
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
  for (int i = 0; i < M; ++i) {
    const int img = pb->obs_image[i], pt = pb->obs_point[i], cam = pb->image_camera[img];
    double* jc = c->Jc + 12 * (size_t)i;
    double* jp = c->Jp + 6 * (size_t)i;
    double* jk = c->Jk + 6 * (size_t)i;
    for (int k = 0; k < 6; ++k) { jc[k] *= c->scale_c[6 * img + k]; jc[6 + k] *= c->scale_c[6 * img + k]; }
    for (int k = 0; k < 3; ++k) { jp[k] *= c->scale_p[3 * pt + k]; jp[3 + k] *= c->scale_p[3 * pt + k]; }
    for (int k = 0; k < 3; ++k) {
      jk[k] *= c->scale_c[6 * F + 3 * cam + k]; jk[3 + k] *= c->scale_c[6 * F + 3 * cam + k];
    }
  }
}

/* ---------------------------------------------------------------- dense Cholesky */

/* in-place lower Cholesky of row-major n x n (upper part ignored); 0 on success */
static int chol_lower(double* A, int n, int nthreads) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = (k0 + NB < n) ? NB : n - k0;
    /* factor diagonal block */
    for (int j = k0; j < k0 + kb; ++j) {
      double d = A[(size_t)j * n + j];
      for (int k = k0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
      if (!(d > 0.0) || !isfinite(d)) return 1;
      d = sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k0 + kb; ++i) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / d;
      }
    }
    const int r0 = k0 + kb;
    if (r0 >= n) break;
    /* panel: rows below, solve X * L11' = A21 */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int i = r0; i < n; ++i) {
      for (int j = k0; j < k0 + kb; ++j) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
      }
    }
    /* trailing update A22 -= L21 L21' (lower part) */
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
    for (int i = r0; i < n; ++i) {
      const double* li = A + (size_t)i * n + k0;
      for (int j = r0; j <= i; ++j) {
        const double* lj = A + (size_t)j * n + k0;
        double s = 0.0;
        for (int k = 0; k < kb; ++k) s += li[k] * lj[k];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  return 0;
}

static void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

/* ---------------------------------------------------------------- Schur pieces */

/* Hinv_p = (E'E + D_p^2)^-1 ; gs_p = E' r */
static int build_point_blocks(ctx_t* c) {
  int fail = 0;
  const int P = c->P;
#pragma omp parallel for schedule(static) reduction(| : fail) num_threads(c->nthreads)
  for (int p = 0; p < P; ++p) {
    if (c->pt_ptr[p] == c->pt_ptr[p + 1]) continue;
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = c->pt_ptr[p]; e < c->pt_ptr[p + 1]; ++e) {
      const double* jp = c->Jp + 6 * (size_t)c->pt_obs[e];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[3 * a + b] += jp[a] * jp[b] + jp[3 + a] * jp[3 + b];
    }
    for (int a = 0; a < 3; ++a) H[4 * a] += c->D_p[3 * p + a] * c->D_p[3 * p + a];
    fail |= inv3_spd(H, c->Hinv + 9 * (size_t)p);
  }
  return fail;
}

/* y = S x (implicit), slot space */
static void schur_multiply(ctx_t* c, const double* x, double* y) {
  const int NS = c->NS, P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  memset(c->tbuf, 0, sizeof(double) * c->tbuf_stride * c->nthreads);
#pragma omp parallel num_threads(c->nthreads)
  {
#ifdef _OPENMP
    double* buf = c->tbuf + c->tbuf_stride * omp_get_thread_num();
#else
    double* buf = c->tbuf;
#endif
#pragma omp for schedule(static)
    for (int p = 0; p < P; ++p) {
      const int b = c->pt_ptr[p], e1 = c->pt_ptr[p + 1];
      if (b == e1) continue;
      double tp[3] = {0, 0, 0};
      for (int e = b; e < e1; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        double u0 = 0, u1 = 0;
        for (int k = 0; k < 6; ++k) { u0 += jc[k] * x[6 * img + k]; u1 += jc[6 + k] * x[6 * img + k]; }
        for (int k = 0; k < 3; ++k) { u0 += jk[k] * x[6 * F + 3 * cam + k]; u1 += jk[3 + k] * x[6 * F + 3 * cam + k]; }
        for (int k = 0; k < 3; ++k) tp[k] += jp[k] * u0 + jp[3 + k] * u1;
      }
      const double* Hi = c->Hinv + 9 * (size_t)p;
      double wp[3];
      for (int a = 0; a < 3; ++a) wp[a] = Hi[3 * a] * tp[0] + Hi[3 * a + 1] * tp[1] + Hi[3 * a + 2] * tp[2];
      for (int e = b; e < e1; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        double u0 = 0, u1 = 0;
        for (int k = 0; k < 6; ++k) { u0 += jc[k] * x[6 * img + k]; u1 += jc[6 + k] * x[6 * img + k]; }
        for (int k = 0; k < 3; ++k) { u0 += jk[k] * x[6 * F + 3 * cam + k]; u1 += jk[3 + k] * x[6 * F + 3 * cam + k]; }
        const double v0 = u0 - (jp[0] * wp[0] + jp[1] * wp[1] + jp[2] * wp[2]);
        const double v1 = u1 - (jp[3] * wp[0] + jp[4] * wp[1] + jp[5] * wp[2]);
        for (int k = 0; k < 6; ++k) buf[6 * img + k] += jc[k] * v0 + jc[6 + k] * v1;
        for (int k = 0; k < 3; ++k) buf[6 * F + 3 * cam + k] += jk[k] * v0 + jk[3 + k] * v1;
      }
    }
  }
  for (int k = 0; k < NS; ++k) {
    double s = 0.0;
    for (int th = 0; th < c->nthreads; ++th) s += c->tbuf[c->tbuf_stride * th + k];
    y[k] = c->slot_active[k] ? s + c->D_c[k] * c->D_c[k] * x[k] : 0.0;
  }
}

/* rhs = F'b - F'E (E'E)^-1 E'b, slot space */
static void schur_rhs(ctx_t* c, double* rhs) {
  const int NS = c->NS, P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  memset(c->tbuf, 0, sizeof(double) * c->tbuf_stride * c->nthreads);
#pragma omp parallel num_threads(c->nthreads)
  {
#ifdef _OPENMP
    double* buf = c->tbuf + c->tbuf_stride * omp_get_thread_num();
#else
    double* buf = c->tbuf;
#endif
#pragma omp for schedule(static)
    for (int p = 0; p < P; ++p) {
      const int b = c->pt_ptr[p], e1 = c->pt_ptr[p + 1];
      if (b == e1) continue;
      const double* Hi = c->Hinv + 9 * (size_t)p;
      const double* gp = c->gs_p + 3 * p;
      double wp[3];
      for (int a = 0; a < 3; ++a) wp[a] = Hi[3 * a] * gp[0] + Hi[3 * a + 1] * gp[1] + Hi[3 * a + 2] * gp[2];
      for (int e = b; e < e1; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        const double v0 = jp[0] * wp[0] + jp[1] * wp[1] + jp[2] * wp[2];
        const double v1 = jp[3] * wp[0] + jp[4] * wp[1] + jp[5] * wp[2];
        for (int k = 0; k < 6; ++k) buf[6 * img + k] += jc[k] * v0 + jc[6 + k] * v1;
        for (int k = 0; k < 3; ++k) buf[6 * F + 3 * cam + k] += jk[k] * v0 + jk[3 + k] * v1;
      }
    }
  }
  for (int k = 0; k < NS; ++k) {
    double s = 0.0;
    for (int th = 0; th < c->nthreads; ++th) s += c->tbuf[c->tbuf_stride * th + k];
    rhs[k] = c->slot_active[k] ? c->gs_c[k] - s : 0.0;
  }
}

/* y_p = Hinv (E'b - E'F y_c) */
static void back_substitute(ctx_t* c, const double* yc, double* yp) {
  const int P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
  for (int p = 0; p < P; ++p) {
    const int b = c->pt_ptr[p], e1 = c->pt_ptr[p + 1];
    if (b == e1) { yp[3 * p] = yp[3 * p + 1] = yp[3 * p + 2] = 0.0; continue; }
    double tp[3] = {c->gs_p[3 * p], c->gs_p[3 * p + 1], c->gs_p[3 * p + 2]};
    for (int e = b; e < e1; ++e) {
      const int i = c->pt_obs[e];
      const double* jc = c->Jc + 12 * (size_t)i;
      const double* jp = c->Jp + 6 * (size_t)i;
      const double* jk = c->Jk + 6 * (size_t)i;
      const int img = pb->obs_image[i], cam = pb->image_camera[img];
      double u0 = 0, u1 = 0;
      for (int k = 0; k < 6; ++k) { u0 += jc[k] * yc[6 * img + k]; u1 += jc[6 + k] * yc[6 * img + k]; }
      for (int k = 0; k < 3; ++k) { u0 += jk[k] * yc[6 * F + 3 * cam + k]; u1 += jk[3 + k] * yc[6 * F + 3 * cam + k]; }
      for (int k = 0; k < 3; ++k) tp[k] -= jp[k] * u0 + jp[3 + k] * u1;
    }
    const double* Hi = c->Hinv + 9 * (size_t)p;
    for (int a = 0; a < 3; ++a) yp[3 * p + a] = Hi[3 * a] * tp[0] + Hi[3 * a + 1] * tp[1] + Hi[3 * a + 2] * tp[2];
  }
}

/* explicit S (slot space, dense, symmetric).  Only block pairs with row slot <= col slot
   are accumulated; the lower triangle is mirrored at the end. */
#define S_ADD(idx, val)                                  \
  do {                                                   \
    if (atomic) {                                        \
      _Pragma("omp atomic") Sx[idx] += (val);            \
    } else {                                             \
      Sx[idx] += (val);                                  \
    }                                                    \
  } while (0)

static void build_schur_dense(ctx_t* c, double* S) {
  const int NS = c->NS, P = c->P, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  const size_t nn = (size_t)NS * NS;
  /* the assembly runs on at most 16 threads, each with a private copy of S (kept for the
     whole solve); more copies cost more in zeroing/merging than they gain */
  const int nbuf = c->nthreads < 16 ? c->nthreads : 16;
  int use_private = (nn * sizeof(double) * (size_t)nbuf) < ((size_t)6 << 30);
  double* priv = NULL;
  if (use_private && nbuf > 1) {
    if (!c->Spriv || c->nspriv != nbuf - 1) {
      free(c->Spriv);
      c->Spriv = (double*)malloc(nn * (size_t)(nbuf - 1) * sizeof(double));
      c->nspriv = nbuf - 1;
    }
    priv = c->Spriv;
    if (!priv) use_private = 0;
  }
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
  for (long long k = 0; k < (long long)nn; ++k) S[k] = 0.0;
  if (priv) {
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
    for (long long k = 0; k < (long long)(nn * (size_t)(nbuf - 1)); ++k) priv[k] = 0.0;
  }
#pragma omp parallel num_threads(nbuf)
  {
#ifdef _OPENMP
    const int th = omp_get_thread_num();
#else
    const int th = 0;
#endif
    double* Sx = (use_private && th > 0) ? priv + nn * (size_t)(th - 1) : S;
    const int atomic = !use_private && nbuf > 1;
    int wcap = 256;
    double* W = (double*)malloc(sizeof(double) * 27 * (size_t)wcap);  /* per obs: 6x3 pose | 3x3 intr */
    double* WH = (double*)malloc(sizeof(double) * 27 * (size_t)wcap); /* W * Hinv */
#pragma omp for schedule(dynamic, 64)
    for (int p = 0; p < P; ++p) {
      const int b = c->pt_ptr[p], e1 = c->pt_ptr[p + 1], L = e1 - b;
      if (L == 0) continue;
      if (L > wcap) {
        wcap = L;
        W = (double*)realloc(W, sizeof(double) * 27 * (size_t)wcap);
        WH = (double*)realloc(WH, sizeof(double) * 27 * (size_t)wcap);
      }
      const double* Hi = c->Hinv + 9 * (size_t)p;
      for (int e = b; e < e1; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jp = c->Jp + 6 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        double* w = W + 27 * (size_t)(e - b);
        for (int a = 0; a < 6; ++a)
          for (int k = 0; k < 3; ++k) w[3 * a + k] = jc[a] * jp[k] + jc[6 + a] * jp[3 + k];
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 3; ++k) w[18 + 3 * a + k] = jk[a] * jp[k] + jk[3 + a] * jp[3 + k];
        double* wh = WH + 27 * (size_t)(e - b);
        for (int a = 0; a < 9; ++a)
          for (int k = 0; k < 3; ++k)
            wh[3 * a + k] = w[3 * a] * Hi[k] + w[3 * a + 1] * Hi[3 + k] + w[3 * a + 2] * Hi[6 + k];
        /* F'F contributions */
        const int sc = 6 * img, sk = 6 * F + 3 * cam;
        for (int a = 0; a < 6; ++a) {
          for (int k = 0; k < 6; ++k) S_ADD((size_t)(sc + a) * NS + sc + k, jc[a] * jc[k] + jc[6 + a] * jc[6 + k]);
          for (int k = 0; k < 3; ++k) S_ADD((size_t)(sc + a) * NS + sk + k, jc[a] * jk[k] + jc[6 + a] * jk[3 + k]);
        }
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 3; ++k) S_ADD((size_t)(sk + a) * NS + sk + k, jk[a] * jk[k] + jk[3 + a] * jk[3 + k]);
      }
      /* - W Hinv W' over ordered pairs of f-blocks with row slot <= col slot */
      for (int ei = 0; ei < L; ++ei) {
        const int ii = c->pt_obs[b + ei];
        const int imgi = pb->obs_image[ii], cami = pb->image_camera[imgi];
        const double* whi = WH + 27 * (size_t)ei;
        const int ri[2] = {6 * imgi, 6 * F + 3 * cami};
        const int dd[2] = {6, 3};
        for (int ej = 0; ej < L; ++ej) {
          const int jj = c->pt_obs[b + ej];
          const int imgj = pb->obs_image[jj], camj = pb->image_camera[imgj];
          const double* wj = W + 27 * (size_t)ej;
          const int rj[2] = {6 * imgj, 6 * F + 3 * camj};
          for (int bi = 0; bi < 2; ++bi)
            for (int bj = 0; bj < 2; ++bj) {
              if (ri[bi] > rj[bj]) continue;
              const double* A = whi + (bi ? 18 : 0);
              const double* B = wj + (bj ? 18 : 0);
              for (int a = 0; a < dd[bi]; ++a)
                for (int k = 0; k < dd[bj]; ++k)
                  S_ADD((size_t)(ri[bi] + a) * NS + rj[bj] + k,
                        -(A[3 * a] * B[3 * k] + A[3 * a + 1] * B[3 * k + 1] + A[3 * a + 2] * B[3 * k + 2]));
            }
        }
      }
    }
    free(W); free(WH);
  }
  if (use_private && nbuf > 1) {
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
    for (long long k = 0; k < (long long)nn; ++k) {
      double s = S[k];
      for (int th = 0; th < nbuf - 1; ++th) s += priv[nn * (size_t)th + k];
      S[k] = s;
    }
  }
#pragma omp parallel for schedule(static) num_threads(c->nthreads)
  for (int i = 0; i < NS; ++i)
    for (int j = i + 1; j < NS; ++j) S[(size_t)j * NS + i] = S[(size_t)i * NS + j];
  for (int k = 0; k < NS; ++k) {
    if (c->slot_active[k]) S[(size_t)k * NS + k] += c->D_c[k] * c->D_c[k];
    else {
      for (int j = 0; j < NS; ++j) { S[(size_t)k * NS + j] = 0.0; S[(size_t)j * NS + k] = 0.0; }
      S[(size_t)k * NS + k] = 1.0;
    }
  }
}

/* block-Jacobi preconditioner of S: blocks = parameter blocks (rot3 | t3 | intr3);
   writes the INVERSE blocks to Minv [(2F + C) * 9] */
static void build_schur_jacobi(ctx_t* c, double* Minv) {
  const int F = c->F, C = c->C, P = c->P;
  const psfm_ba_problem* pb = c->pb;
  const int NB = 2 * F + C;
  double* B = (double*)calloc((size_t)NB * 9 * c->nthreads, sizeof(double));
#pragma omp parallel num_threads(c->nthreads)
  {
#ifdef _OPENMP
    double* Bx = B + (size_t)NB * 9 * omp_get_thread_num();
#else
    double* Bx = B;
#endif
#pragma omp for schedule(static)
    for (int p = 0; p < P; ++p) {
      const int b = c->pt_ptr[p], e1 = c->pt_ptr[p + 1];
      if (b == e1) continue;
      const double* Hi = c->Hinv + 9 * (size_t)p;
      for (int e = b; e < e1; ++e) {
        const int i = c->pt_obs[e];
        const double* jc = c->Jc + 12 * (size_t)i;
        const double* jk = c->Jk + 6 * (size_t)i;
        const int img = pb->obs_image[i], cam = pb->image_camera[img];
        /* F'F diagonal blocks */
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 3; ++k) {
            Bx[(size_t)(2 * img) * 9 + 3 * a + k] += jc[a] * jc[k] + jc[6 + a] * jc[6 + k];
            Bx[(size_t)(2 * img + 1) * 9 + 3 * a + k] += jc[3 + a] * jc[3 + k] + jc[9 + a] * jc[9 + k];
            Bx[(size_t)(2 * F + cam) * 9 + 3 * a + k] += jk[a] * jk[k] + jk[3 + a] * jk[3 + k];
          }
      }
      /* - sum over obs pairs sharing the block: W_i Hinv W_j' */
      for (int ei = b; ei < e1; ++ei) {
        const int i = c->pt_obs[ei];
        const int imgi = pb->obs_image[i], cami = pb->image_camera[imgi];
        const double* jci = c->Jc + 12 * (size_t)i;
        const double* jpi = c->Jp + 6 * (size_t)i;
        const double* jki = c->Jk + 6 * (size_t)i;
        double Wi[27], WHi[27]; /* rot 3x3, t 3x3, intr 3x3 (each rows of block, cols point) */
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 3; ++k) {
            Wi[3 * a + k] = jci[a] * jpi[k] + jci[6 + a] * jpi[3 + k];
            Wi[9 + 3 * a + k] = jci[3 + a] * jpi[k] + jci[9 + a] * jpi[3 + k];
            Wi[18 + 3 * a + k] = jki[a] * jpi[k] + jki[3 + a] * jpi[3 + k];
          }
        for (int a = 0; a < 9; ++a)
          for (int k = 0; k < 3; ++k)
            WHi[3 * a + k] = Wi[3 * a] * Hi[k] + Wi[3 * a + 1] * Hi[3 + k] + Wi[3 * a + 2] * Hi[6 + k];
        for (int ej = b; ej < e1; ++ej) {
          const int j = c->pt_obs[ej];
          const int imgj = pb->obs_image[j], camj = pb->image_camera[imgj];
          const double* jcj = c->Jc + 12 * (size_t)j;
          const double* jpj = c->Jp + 6 * (size_t)j;
          const double* jkj = c->Jk + 6 * (size_t)j;
          if (imgj == imgi) {
            for (int a = 0; a < 3; ++a)
              for (int k = 0; k < 3; ++k) {
                double wr[3], wt[3];
                for (int m = 0; m < 3; ++m) {
                  wr[m] = jcj[k] * jpj[m] + jcj[6 + k] * jpj[3 + m];
                  wt[m] = jcj[3 + k] * jpj[m] + jcj[9 + k] * jpj[3 + m];
                }
                Bx[(size_t)(2 * imgi) * 9 + 3 * a + k] -= WHi[3 * a] * wr[0] + WHi[3 * a + 1] * wr[1] + WHi[3 * a + 2] * wr[2];
                Bx[(size_t)(2 * imgi + 1) * 9 + 3 * a + k] -= WHi[9 + 3 * a] * wt[0] + WHi[9 + 3 * a + 1] * wt[1] + WHi[9 + 3 * a + 2] * wt[2];
              }
          }
          if (camj == cami) {
            for (int a = 0; a < 3; ++a)
              for (int k = 0; k < 3; ++k) {
                double wk[3];
                for (int m = 0; m < 3; ++m) wk[m] = jkj[k] * jpj[m] + jkj[3 + k] * jpj[3 + m];
                Bx[(size_t)(2 * F + cami) * 9 + 3 * a + k] -= WHi[18 + 3 * a] * wk[0] + WHi[18 + 3 * a + 1] * wk[1] + WHi[18 + 3 * a + 2] * wk[2];
              }
          }
        }
      }
    }
  }
  for (int nb = 0; nb < NB; ++nb) {
    double A[9];
    for (int k = 0; k < 9; ++k) {
      double s = 0.0;
      for (int th = 0; th < c->nthreads; ++th) s += B[(size_t)NB * 9 * th + (size_t)nb * 9 + k];
      A[k] = s;
    }
    const int slot0 = (nb < 2 * F) ? 3 * nb : 6 * F + 3 * (nb - 2 * F);
    for (int a = 0; a < 3; ++a) A[4 * a] += c->D_c[slot0 + a] * c->D_c[slot0 + a];
    inv_small_masked(A, c->slot_active + slot0, 3, Minv + (size_t)nb * 9);
    /* inactive dims: identity in Minv but vectors are zero there anyway */
  }
  free(B);
}

static double dotn(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ceres ConjugateGradientsSolver::Solve on the implicit Schur complement.
   returns termination: 0 success, 1 no-convergence, 2 failure */
static int schur_pcg(ctx_t* c, const double* b, const double* Minv, double q_tol, double r_tol,
                     int max_it, int min_it, double* x, int* iters) {
  const int n = c->NS;
  const int NB = 2 * c->F + c->C;
  double* r = (double*)malloc(sizeof(double) * n * 4);
  double *pv = r + n, *z = r + 2 * n, *tmp = r + 3 * n;
  *iters = 0;
  memset(x, 0, sizeof(double) * n);
  const double norm_b = sqrt(dotn(b, b, n));
  if (norm_b == 0.0) { free(r); return 0; }
  const double tol_r = r_tol * norm_b;
  schur_multiply(c, x, tmp);
  for (int i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
  double norm_r = sqrt(dotn(r, r, n));
  if (min_it == 0 && norm_r <= tol_r) { free(r); return 0; }
  double rho = 1.0;
  double Q0 = 0.0;
  for (int i = 0; i < n; ++i) Q0 += x[i] * (b[i] + r[i]);
  Q0 = -1.0 * Q0;
  int term = 1;
  for (int it = 1;; ++it) {
    *iters = it;
    /* z = M^-1 r */
    for (int nb = 0; nb < NB; ++nb) {
      const int s0 = (nb < 2 * c->F) ? 3 * nb : 6 * c->F + 3 * (nb - 2 * c->F);
      const double* Mi = Minv + (size_t)nb * 9;
      for (int a = 0; a < 3; ++a)
        z[s0 + a] = Mi[3 * a] * r[s0] + Mi[3 * a + 1] * r[s0 + 1] + Mi[3 * a + 2] * r[s0 + 2];
    }
    const double last_rho = rho;
    rho = dotn(r, z, n);
    if (rho == 0.0 || !isfinite(rho)) { term = 2; break; }
    if (it == 1) memcpy(pv, z, sizeof(double) * n);
    else {
      const double beta = rho / last_rho;
      if (beta == 0.0 || !isfinite(beta)) { term = 2; break; }
      for (int i = 0; i < n; ++i) pv[i] = z[i] + beta * pv[i];
    }
    double* qv = z;
    schur_multiply(c, pv, qv);
    const double pq = dotn(pv, qv, n);
    if (pq <= 0.0 || isinf(pq)) { term = 1; break; }
    const double alpha = rho / pq;
    if (isinf(alpha)) { term = 2; break; }
    for (int i = 0; i < n; ++i) x[i] = x[i] + alpha * pv[i];
    if (it % CERES_CG_RESIDUAL_RESET_PERIOD == 0) {
      schur_multiply(c, x, tmp);
      for (int i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
    } else {
      for (int i = 0; i < n; ++i) r[i] = r[i] - alpha * qv[i];
    }
    double Q1 = 0.0;
    for (int i = 0; i < n; ++i) Q1 += x[i] * (b[i] + r[i]);
    Q1 = -1.0 * Q1;
    const double zeta = it * (Q1 - Q0) / Q1;
    if (zeta < q_tol && it >= min_it) { term = 0; break; }
    Q0 = Q1;
    norm_r = sqrt(dotn(r, r, n));
    if (norm_r <= tol_r && it >= min_it) { term = 0; break; }
    if (it >= max_it) break;
  }
  free(r);
  return term;
}

/* ---------------------------------------------------------------- LM strategy */

static int resolve_solver(const ctx_t* c) {
  int s = c->o.linear_solver;
  if (s == PSFM_BA_SOLVER_AUTO) {
    /* bundle_adjustment.cc:276-286 (num_images = images in the config) */
    s = (c->F <= 1000) ? PSFM_BA_SOLVER_EXACT_SCHUR : PSFM_BA_SOLVER_ITERATIVE_SCHUR;
  }
  return s;
}

/* solves (J'J + D^2) y = J' r through the Schur complement; step = -y.
   returns 0 ok, 2 failure */
static int linear_solve(ctx_t* c, int solver) {
  const int NS = c->NS, P = c->P;
  { TM_BEGIN(); jt_r(c, c->gs_c, c->gs_p); /* scaled-J gradient */
    if (build_point_blocks(c)) return 2; TM_END(1); }
  double* yc = c->step_c;
  double* yp = c->step_p;
  c->num_linear_iterations = 0;
  if (solver == PSFM_BA_SOLVER_EXACT_SCHUR) {
    /* SPARSE_SCHUR restatement (block-sparse S, band Cholesky) for the pipeline's single shared
       camera; PSFM_ORACLE_DENSE_SCHUR=1 forces the general dense assembly (the two are compared
       in tests/test_oracle_ba.py) */
    if (c->C == 1 && !c->sb.ready && !c->sb_unusable) {
      if (getenv("PSFM_ORACLE_DENSE_SCHUR") ||
          sblocks_pattern(&c->sb, c->F, P, c->pt_ptr, c->pt_obs, c->pb->obs_image, c->nthreads)) {
        sblocks_free(&c->sb);
        c->sb_unusable = 1;
      }
    }
    if (c->sb.ready) {
      { TM_BEGIN(); sblocks_assemble(&c->sb, c->F, P, c->pt_ptr, c->pt_obs, c->pb->obs_image, c->Jc, c->Jp, c->Jk, c->Hinv, c->nthreads); TM_END(2); }
      { TM_BEGIN(); schur_rhs(c, yc); TM_END(3); }
      if (2 * (6 * c->sb.span + 5) < 6 * c->F) {
        TM_BEGIN();
        const int bad = sblocks_band_solve(&c->sb, c->F, c->slot_active, c->D_c, yc);
        TM_END(4);
        if (bad) return 2;
      } else {
        if (!c->S) c->S = (double*)malloc(sizeof(double) * (size_t)NS * NS);
        sblocks_to_dense(&c->sb, c->F, c->slot_active, c->D_c, c->S);
        if (chol_lower(c->S, NS, c->nthreads)) return 2;
        chol_solve(c->S, NS, yc);
      }
    } else {
      if (!c->S) c->S = (double*)malloc(sizeof(double) * (size_t)NS * NS);
      build_schur_dense(c, c->S);
      schur_rhs(c, yc);
      if (chol_lower(c->S, NS, c->nthreads)) return 2;
      chol_solve(c->S, NS, yc);
    }
    c->num_linear_iterations = 1;
  } else {
    double* rhs = (double*)malloc(sizeof(double) * NS);
    double* Minv = (double*)malloc(sizeof(double) * 9 * (size_t)(2 * c->F + c->C));
    schur_rhs(c, rhs);
    build_schur_jacobi(c, Minv);
    int iters = 0;
    const int term = schur_pcg(c, rhs, Minv, c->o.eta, -1.0, c->o.max_linear_solver_iterations,
                               CERES_MIN_LINEAR_SOLVER_ITERATIONS, yc, &iters);
    c->num_linear_iterations = iters;
    free(rhs); free(Minv);
    if (term == 2) return 2;
  }
  { TM_BEGIN(); back_substitute(c, yc, yp); TM_END(5); }
  for (int k = 0; k < NS; ++k) { if (!isfinite(yc[k])) return 2; yc[k] = -yc[k]; }
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) num_threads(c->nthreads)
  for (int k = 0; k < 3 * P; ++k) { if (!isfinite(yp[k])) bad = 1; yp[k] = -yp[k]; }
  return bad ? 2 : 0;
}

/* model_cost_change = -(J s)'(r + J s / 2) */
static double model_cost_change(ctx_t* c) {
  const int M = c->M, F = c->F;
  const psfm_ba_problem* pb = c->pb;
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(c->nthreads)
  for (int i = 0; i < M; ++i) {
    const double* jc = c->Jc + 12 * (size_t)i;
    const double* jp = c->Jp + 6 * (size_t)i;
    const double* jk = c->Jk + 6 * (size_t)i;
    const int img = pb->obs_image[i], pt = pb->obs_point[i], cam = pb->image_camera[img];
    double m0 = 0, m1 = 0;
    for (int k = 0; k < 6; ++k) { m0 += jc[k] * c->step_c[6 * img + k]; m1 += jc[6 + k] * c->step_c[6 * img + k]; }
    for (int k = 0; k < 3; ++k) { m0 += jp[k] * c->step_p[3 * pt + k]; m1 += jp[3 + k] * c->step_p[3 * pt + k]; }
    for (int k = 0; k < 3; ++k) { m0 += jk[k] * c->step_c[6 * F + 3 * cam + k]; m1 += jk[3 + k] * c->step_c[6 * F + 3 * cam + k]; }
    acc += m0 * (c->r[2 * i] + m0 / 2.0) + m1 * (c->r[2 * i + 1] + m1 / 2.0);
  }
  return -acc;
}

/* candidate = Plus(x, step o scale); returns |x - x_c|^2 */
static double make_candidate(ctx_t* c) {
  const int F = c->F, P = c->P, C = c->C;
  double sn = 0.0;
  for (int i = 0; i < F; ++i) {
    memcpy(c->qc + 4 * i, c->q + 4 * i, 4 * sizeof(double));
    memcpy(c->tc + 3 * i, c->t + 3 * i, 3 * sizeof(double));
    if (!c->pose_var[i]) continue;
    const unsigned char* act = c->slot_active + 6 * i;
    if (act[0]) {
      double d[3];
      for (int k = 0; k < 3; ++k) d[k] = c->step_c[6 * i + k] * c->scale_c[6 * i + k];
      quat_plus(c->q + 4 * i, d, c->qc + 4 * i);
      for (int k = 0; k < 4; ++k) { const double e = c->q[4 * i + k] - c->qc[4 * i + k]; sn += e * e; }
    }
    for (int k = 0; k < 3; ++k)
      if (act[3 + k]) {
        const double d = c->step_c[6 * i + 3 + k] * c->scale_c[6 * i + 3 + k];
        c->tc[3 * i + k] = c->t[3 * i + k] + d;
        const double e = c->t[3 * i + k] - c->tc[3 * i + k];
        sn += e * e;
      }
  }
  for (int k = 0; k < 3 * C; ++k) {
    c->Kc[k] = c->K[k];
    if (c->slot_active[6 * F + k]) {
      c->Kc[k] = c->K[k] + c->step_c[6 * F + k] * c->scale_c[6 * F + k];
      const double e = c->K[k] - c->Kc[k];
      sn += e * e;
    }
  }
  double snp = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : snp) num_threads(c->nthreads)
  for (int p = 0; p < P; ++p) {
    for (int k = 0; k < 3; ++k) {
      if (c->pt_ptr[p] == c->pt_ptr[p + 1]) { c->Xc[3 * p + k] = c->X[3 * p + k]; continue; }
      c->Xc[3 * p + k] = c->X[3 * p + k] + c->step_p[3 * p + k] * c->scale_p[3 * p + k];
      const double e = c->X[3 * p + k] - c->Xc[3 * p + k];
      snp += e * e;
    }
  }
  return sn + snp;
}

/* |x|^2 over the parameter blocks of the reduced program (ambient sizes) */
static double x_sqnorm(const ctx_t* c) {
  const int F = c->F, P = c->P, C = c->C;
  double s = 0.0;
  for (int i = 0; i < F; ++i) {
    if (!c->pose_var[i]) continue;
    const unsigned char* act = c->slot_active + 6 * i;
    if (act[0]) for (int k = 0; k < 4; ++k) s += c->q[4 * i + k] * c->q[4 * i + k];
    if (act[3] || act[4] || act[5]) for (int k = 0; k < 3; ++k) s += c->t[3 * i + k] * c->t[3 * i + k];
  }
  for (int cc = 0; cc < C; ++cc) {
    const unsigned char* act = c->slot_active + 6 * F + 3 * cc;
    if (act[0] || act[1] || act[2]) for (int k = 0; k < 3; ++k) s += c->K[3 * cc + k] * c->K[3 * cc + k];
  }
  double sp = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sp) num_threads(c->nthreads)
  for (int p = 0; p < P; ++p) {
    if (c->pt_ptr[p] == c->pt_ptr[p + 1]) continue;
    for (int k = 0; k < 3; ++k) sp += c->X[3 * p + k] * c->X[3 * p + k];
  }
  return s + sp;
}

/* gradient_max_norm = |Plus(x, -g) - x|_inf */
static double gradient_max_norm(const ctx_t* c) {
  const int F = c->F, P = c->P, C = c->C;
  double m = 0.0;
  for (int i = 0; i < F; ++i) {
    if (!c->pose_var[i]) continue;
    const unsigned char* act = c->slot_active + 6 * i;
    if (act[0]) {
      double d[3] = {-c->g_c[6 * i], -c->g_c[6 * i + 1], -c->g_c[6 * i + 2]};
      double qn[4];
      quat_plus(c->q + 4 * i, d, qn);
      for (int k = 0; k < 4; ++k) m = fmax(m, fabs(c->q[4 * i + k] - qn[k]));
    }
    for (int k = 0; k < 3; ++k) if (act[3 + k]) m = fmax(m, fabs(c->g_c[6 * i + 3 + k]));
  }
  for (int k = 0; k < 3 * C; ++k) if (c->slot_active[6 * F + k]) m = fmax(m, fabs(c->g_c[6 * F + k]));
  for (int k = 0; k < 3 * P; ++k) m = fmax(m, fabs(c->g_p[k]));
  return m;
}

/* ---------------------------------------------------------------- context */

static int ctx_init(ctx_t* c, const psfm_ba_problem* pb, const psfm_ba_options* o, int nthreads) {
  memset(c, 0, sizeof(*c));
  c->pb = pb; c->o = *o;
  c->F = pb->num_images; c->P = pb->num_points; c->M = pb->num_observations; c->C = pb->num_cameras;
  c->NS = 6 * c->F + 3 * c->C;
  /* num_threads = min(cpu_count, 64): sfm/main_sfm.py:144 (--GlobalMapper.num_threads) */
  if (nthreads <= 0) nthreads = psfm_oracle_num_threads() < 64 ? psfm_oracle_num_threads() : 64;
  if (nthreads > MAXT) nthreads = MAXT;
  c->nthreads = nthreads;
  const int F = c->F, P = c->P, M = c->M, C = c->C, NS = c->NS;
  for (int i = 0; i < M; ++i) {
    if (pb->obs_image[i] < 0 || pb->obs_image[i] >= F || pb->obs_point[i] < 0 || pb->obs_point[i] >= P) return PSFM_ERR_INVALID;
  }
  for (int i = 0; i < F; ++i) if (pb->image_camera[i] < 0 || pb->image_camera[i] >= C) return PSFM_ERR_INVALID;
  c->pt_ptr = (int*)calloc((size_t)P + 2, sizeof(int));
  c->pt_obs = (int*)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
  c->img_has_obs = (unsigned char*)calloc((size_t)F + 1, 1);
  c->cam_has_obs = (unsigned char*)calloc((size_t)C + 1, 1);
  c->pose_var = (unsigned char*)calloc((size_t)F + 1, 1);
  c->slot_active = (unsigned char*)calloc((size_t)NS + 1, 1);
  for (int i = 0; i < M; ++i) {
    c->pt_ptr[pb->obs_point[i] + 1]++;
    c->img_has_obs[pb->obs_image[i]] = 1;
    c->cam_has_obs[pb->image_camera[pb->obs_image[i]]] = 1;
  }
  for (int p = 0; p < P; ++p) c->pt_ptr[p + 1] += c->pt_ptr[p];
  {
    int* fill = (int*)malloc(sizeof(int) * (size_t)(P + 1));
    memcpy(fill, c->pt_ptr, sizeof(int) * (size_t)(P + 1));
    c->orig = (int*)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
    c->s_img = (int32_t*)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    c->s_pt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    c->s_xy = (double*)malloc(sizeof(double) * 2 * (size_t)(M > 0 ? M : 1));
    for (int i = 0; i < M; ++i) c->orig[fill[pb->obs_point[i]]++] = i;     /* stable counting sort by point */
    free(fill);
    for (int e = 0; e < M; ++e) {
      const int i = c->orig[e];
      c->pt_obs[e] = e;
      c->s_img[e] = pb->obs_image[i]; c->s_pt[e] = pb->obs_point[i];
      c->s_xy[2 * e] = pb->obs_xy[2 * i]; c->s_xy[2 * e + 1] = pb->obs_xy[2 * i + 1];
    }
    c->spb = *pb;
    c->spb.obs_image = c->s_img; c->spb.obs_point = c->s_pt; c->spb.obs_xy = c->s_xy;
    c->pb = &c->spb;
  }
  for (int i = 0; i < F; ++i) {
    const int constant_pose = !o->refine_extrinsics || (pb->pose_constant && pb->pose_constant[i]);
    c->pose_var[i] = c->img_has_obs[i] && !constant_pose;
    if (!c->pose_var[i]) continue;
    const unsigned tm = pb->tvec_constant_mask ? pb->tvec_constant_mask[i] : 0;
    for (int k = 0; k < 3; ++k) {
      c->slot_active[6 * i + k] = o->refine_rotation ? 1 : 0;
      c->slot_active[6 * i + 3 + k] = ((tm >> k) & 1) ? 0 : 1;
    }
  }
  {
    /* ParameterizeCameras (bundle_adjustment.cc:500-544), SIMPLE_PINHOLE:
       FocalLengthIdxs={0}, PrincipalPointIdxs={1,2}, ExtraParamsIdxs={} */
    const int constant_camera = !o->refine_focal_length && !o->refine_principal_point && !o->refine_extra_params;
    for (int cc = 0; cc < C; ++cc) {
      if (!c->cam_has_obs[cc]) continue;
      if (constant_camera || (pb->camera_constant && pb->camera_constant[cc])) continue;
      c->slot_active[6 * F + 3 * cc] = o->refine_focal_length ? 1 : 0;
      c->slot_active[6 * F + 3 * cc + 1] = o->refine_principal_point ? 1 : 0;
      c->slot_active[6 * F + 3 * cc + 2] = o->refine_principal_point ? 1 : 0;
    }
  }
#define ALLOC(ptr, n) do { ptr = (double*)calloc((size_t)(n) + 1, sizeof(double)); if (!ptr) return PSFM_ERR_INVALID; } while (0)
  ALLOC(c->q, 4 * F); ALLOC(c->t, 3 * F); ALLOC(c->X, 3 * P); ALLOC(c->K, 3 * C);
  ALLOC(c->qc, 4 * F); ALLOC(c->tc, 3 * F); ALLOC(c->Xc, 3 * P); ALLOC(c->Kc, 3 * C);
  ALLOC(c->r, 2 * (size_t)M); ALLOC(c->Jc, 12 * (size_t)M); ALLOC(c->Jp, 6 * (size_t)M); ALLOC(c->Jk, 6 * (size_t)M);
  ALLOC(c->scale_c, NS); ALLOC(c->scale_p, 3 * P); ALLOC(c->g_c, NS); ALLOC(c->g_p, 3 * P);
  ALLOC(c->diag_c, NS); ALLOC(c->diag_p, 3 * P); ALLOC(c->D_c, NS); ALLOC(c->D_p, 3 * P);
  ALLOC(c->Hinv, 9 * (size_t)P); ALLOC(c->gs_c, NS); ALLOC(c->gs_p, 3 * P);
  ALLOC(c->step_c, NS); ALLOC(c->step_p, 3 * P);
  c->tbuf_stride = (size_t)NS + 8;
  ALLOC(c->tbuf, c->tbuf_stride * (size_t)c->nthreads);
#undef ALLOC
  /* image.NormalizeQvec() for every image of the config (bundle_adjustment.cc:355) */
  for (int i = 0; i < F; ++i) {
    const double* qs = pb->qvec + 4 * i;
    const double n = sqrt(qs[0] * qs[0] + qs[1] * qs[1] + qs[2] * qs[2] + qs[3] * qs[3]);
    if (n == 0.0) { c->q[4 * i] = 1.0; c->q[4 * i + 1] = c->q[4 * i + 2] = c->q[4 * i + 3] = 0.0; }
    else for (int k = 0; k < 4; ++k) c->q[4 * i + k] = qs[k] / n;
  }
  memcpy(c->t, pb->tvec, sizeof(double) * 3 * (size_t)F);
  memcpy(c->X, pb->xyz, sizeof(double) * 3 * (size_t)P);
  memcpy(c->K, pb->cam_params, sizeof(double) * 3 * (size_t)C);
  for (int k = 0; k < NS; ++k) c->scale_c[k] = 1.0;
  for (int k = 0; k < 3 * P; ++k) c->scale_p[k] = 1.0;
  return PSFM_OK;
}

static void ctx_free(ctx_t* c) {
  free(c->pt_ptr); free(c->pt_obs); free(c->img_has_obs); free(c->cam_has_obs); free(c->pose_var);
  free(c->slot_active); free(c->q); free(c->t); free(c->X); free(c->K); free(c->qc); free(c->tc);
  free(c->Xc); free(c->Kc); free(c->r); free(c->Jc); free(c->Jp); free(c->Jk); free(c->scale_c);
  free(c->scale_p); free(c->g_c); free(c->g_p); free(c->diag_c); free(c->diag_p); free(c->D_c);
  free(c->D_p); free(c->Hinv); free(c->gs_c); free(c->gs_p); free(c->step_c); free(c->step_p);
  free(c->S); free(c->Spriv); free(c->tbuf);
  free(c->orig); free(c->s_img); free(c->s_pt); free(c->s_xy);
  sblocks_free(&c->sb);
}

/* EvaluateGradientAndJacobian: cost, r, J, g = J'r (unscaled), optional scaling */
static double evaluate_gradient_and_jacobian(ctx_t* c, int iteration) {
  const int NS = c->NS, P = c->P;
  const double cost = evaluate(c, c->q, c->t, c->X, c->K, 1);
  jt_r(c, c->g_c, c->g_p);
  if (c->o.jacobi_scaling) {
    if (iteration == 0) {
      col_sqnorm(c, c->scale_c, c->scale_p);
      for (int k = 0; k < NS; ++k) c->scale_c[k] = 1.0 / (1.0 + sqrt(c->scale_c[k]));
      for (int k = 0; k < 3 * P; ++k) c->scale_p[k] = 1.0 / (1.0 + sqrt(c->scale_p[k]));
    }
    scale_columns(c);
  }
  return cost;
}

static void lm_diagonal(ctx_t* c, double radius, int reuse) {
  const int NS = c->NS, P = c->P;
  if (!reuse) {
    col_sqnorm(c, c->diag_c, c->diag_p);
    for (int k = 0; k < NS; ++k) c->diag_c[k] = fmin(fmax(c->diag_c[k], c->o.min_lm_diagonal), c->o.max_lm_diagonal);
    for (int k = 0; k < 3 * P; ++k) c->diag_p[k] = fmin(fmax(c->diag_p[k], c->o.min_lm_diagonal), c->o.max_lm_diagonal);
  }
  for (int k = 0; k < NS; ++k) c->D_c[k] = c->slot_active[k] ? sqrt(c->diag_c[k] / radius) : 0.0;
  for (int k = 0; k < 3 * P; ++k) c->D_p[k] = sqrt(c->diag_p[k] / radius);
}

int psfm_oracle_ba_solve(psfm_ba_problem* pb, const psfm_ba_options* opts, psfm_ba_summary* sum,
                         int32_t num_threads) {
  psfm_ba_options o;
  if (opts) o = *opts; else psfm_oracle_ba_default_options(&o);
  psfm_ba_summary s;
  memset(&s, 0, sizeof(s));
  if (pb->num_observations == 0) {
    if (o.print_summary) printf("Zero residual for BA\n");
    if (sum) *sum = s;
    return PSFM_ZERO_RESIDUALS;
  }
  ctx_t c;
  int rc = ctx_init(&c, pb, &o, num_threads);
  if (rc != PSFM_OK) { ctx_free(&c); return rc; }
  const double t0 = wall_s();
  const int solver = resolve_solver(&c);
  s.linear_solver_used = solver;
  s.world_size = 1;
  s.num_residuals_reduced = 2 * c.M;
  {
    int np = 0;
    for (int k = 0; k < c.NS; ++k) np += c.slot_active[k];
    for (int p = 0; p < c.P; ++p) if (c.pt_ptr[p] != c.pt_ptr[p + 1]) np += 3;
    s.num_effective_parameters_reduced = np;
  }
  /* ---- TrustRegionMinimizer::Minimize ---- */
  double radius = o.initial_trust_region_radius;
  double decrease_factor = CERES_LM_DECREASE_FACTOR0;
  int reuse_diagonal = 0;
  int num_consecutive_invalid = 0;
  int iteration = 0;
  if (g_tm_on < 0) g_tm_on = getenv("PSFM_ORACLE_TIMING") ? 1 : 0;
  memset(g_tm, 0, sizeof(g_tm));
  double x_cost;
  { TM_BEGIN(); x_cost = evaluate_gradient_and_jacobian(&c, 0); TM_END(0); }
  double x_norm = sqrt(x_sqnorm(&c));
  double gmax = gradient_max_norm(&c);
  s.initial_cost = x_cost;
  s.num_linearize = 1;
  int term = PSFM_TERM_NO_CONVERGENCE;
  int step_ok_prev = 1; /* iteration 0 counts as successful */
  if (o.minimizer_progress_to_stdout)
    printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter\n"
           "%4d % .6e  % .3e  % .3e\n", 0, x_cost, 0.0, gmax);
  for (;;) {
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (step_ok_prev && iteration > 0) s.num_successful_steps++;
    if (iteration >= o.max_num_iterations) { term = PSFM_TERM_NO_CONVERGENCE; break; }
    if (gmax <= o.gradient_tolerance) { term = PSFM_TERM_CONVERGENCE_GRADIENT; break; }
    if (radius <= o.min_trust_region_radius) { term = PSFM_TERM_MIN_RADIUS; break; }
    ++iteration;
    step_ok_prev = 0;
    /* LevenbergMarquardtStrategy::ComputeStep */
    lm_diagonal(&c, radius, reuse_diagonal);
    const int ls = linear_solve(&c, solver);
    reuse_diagonal = 1;
    s.num_linear_iterations += c.num_linear_iterations;
    double mcc = 0.0;
    int valid = 0;
    if (ls == 0) {
      mcc = model_cost_change(&c);
      valid = mcc > 0.0;
    }
    if (!valid) {
      /* HandleInvalidStep */
      if (++num_consecutive_invalid >= o.max_num_consecutive_invalid_steps) { term = PSFM_TERM_FAILURE; s.num_unsuccessful_steps++; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
      s.num_unsuccessful_steps++;
      continue;
    }
    num_consecutive_invalid = 0;
    double step_sq, cand_cost;
    { TM_BEGIN(); step_sq = make_candidate(&c); cand_cost = evaluate(&c, c.qc, c.tc, c.Xc, c.Kc, 0); TM_END(6); }
    /* ParameterToleranceReached */
    const double step_norm = sqrt(step_sq);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { term = PSFM_TERM_CONVERGENCE_PARAMETER; break; }
    /* FunctionToleranceReached */
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * x_cost) { term = PSFM_TERM_CONVERGENCE_FUNCTION; break; }
    const double rho = cost_change / mcc;
    if (rho > o.min_relative_decrease) {
      double* tmp;
      tmp = c.q; c.q = c.qc; c.qc = tmp;
      tmp = c.t; c.t = c.tc; c.tc = tmp;
      tmp = c.X; c.X = c.Xc; c.Xc = tmp;
      tmp = c.K; c.K = c.Kc; c.Kc = tmp;
      x_norm = sqrt(x_sqnorm(&c));
      { TM_BEGIN(); x_cost = evaluate_gradient_and_jacobian(&c, iteration); TM_END(0); }
      s.num_linearize++;
      gmax = gradient_max_norm(&c);
      /* StepAccepted */
      const double t = 2.0 * rho - 1.0;
      radius = radius / fmax(CERES_LM_MIN_SHRINK, 1.0 - t * t * t);
      radius = fmin(o.max_trust_region_radius, radius);
      decrease_factor = CERES_LM_DECREASE_FACTOR0;
      reuse_diagonal = 0;
      step_ok_prev = 1;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
      s.num_unsuccessful_steps++;
    }
    if (o.minimizer_progress_to_stdout)
      printf("%4d % .6e  % .3e  % .3e  % .3e  % .3e  % .3e  %d\n", iteration, step_ok_prev ? x_cost : cand_cost,
             cost_change, gmax, step_norm, rho, radius, c.num_linear_iterations);
  }
  s.num_iterations = iteration;
  s.termination = term;
  s.final_cost = x_cost;
  s.total_time_in_seconds = wall_s() - t0;
  if (g_tm_on > 0)
    fprintf(stderr, "[oracle timing] %d threads, total %.3f s: linearise %.3f | J'r + point blocks %.3f | Schur assembly %.3f | rhs %.3f | "
            "band Cholesky %.3f | back substitution %.3f | candidate + cost %.3f\n", c.nthreads, s.total_time_in_seconds,
            g_tm[0], g_tm[1], g_tm[2], g_tm[3], g_tm[4], g_tm[5], g_tm[6]);
  /* write back (in place, like the reference) */
  memcpy(pb->qvec, c.q, sizeof(double) * 4 * (size_t)c.F);
  memcpy(pb->tvec, c.t, sizeof(double) * 3 * (size_t)c.F);
  memcpy(pb->xyz, c.X, sizeof(double) * 3 * (size_t)c.P);
  memcpy(pb->cam_params, c.K, sizeof(double) * 3 * (size_t)c.C);
  if (o.print_summary) {
    static const char* tn[] = {"Convergence", "Convergence", "Convergence", "No convergence", "Failure", "Convergence"};
    printf("    Residuals : %d\n   Parameters : %d\n   Iterations : %d\n         Time : %g [s]\n"
           " Initial cost : %g [px]\n   Final cost : %g [px]\n  Termination : %s\n\n",
           s.num_residuals_reduced, s.num_effective_parameters_reduced,
           s.num_successful_steps + s.num_unsuccessful_steps, s.total_time_in_seconds,
           sqrt(s.initial_cost / s.num_residuals_reduced), sqrt(s.final_cost / s.num_residuals_reduced),
           tn[term]);
  }
  if (sum) *sum = s;
  ctx_free(&c);
  return PSFM_OK;
}

int psfm_oracle_ba_evaluate(const psfm_ba_problem* pb, const psfm_ba_options* opts, double* cost,
                            double* residuals, double* gradient_cam, double* gradient_pts,
                            int32_t num_threads) {
  psfm_ba_options o;
  if (opts) o = *opts; else psfm_oracle_ba_default_options(&o);
  ctx_t c;
  int rc = ctx_init(&c, pb, &o, num_threads);
  if (rc != PSFM_OK) { ctx_free(&c); return rc; }
  const double cst = evaluate(&c, c.q, c.t, c.X, c.K, 1);
  jt_r(&c, c.g_c, c.g_p);
  if (cost) *cost = cst;
  if (residuals)
    for (int e = 0; e < c.M; ++e) {   /* back to the caller's observation order */
      residuals[2 * (size_t)c.orig[e]] = c.r[2 * (size_t)e];
      residuals[2 * (size_t)c.orig[e] + 1] = c.r[2 * (size_t)e + 1];
    }
  if (gradient_cam) memcpy(gradient_cam, c.g_c, sizeof(double) * (size_t)c.NS);
  if (gradient_pts) memcpy(gradient_pts, c.g_p, sizeof(double) * 3 * (size_t)c.P);
  ctx_free(&c);
  return PSFM_OK;
}

int psfm_oracle_ba_jacobians(const psfm_ba_problem* pb, const psfm_ba_options* opts, double* jc,
                             double* jp, double* jk) {
  psfm_ba_options o;
  if (opts) o = *opts; else psfm_oracle_ba_default_options(&o);
  ctx_t c;
  int rc = ctx_init(&c, pb, &o, 1);
  if (rc != PSFM_OK) { ctx_free(&c); return rc; }
  evaluate(&c, c.q, c.t, c.X, c.K, 1);
  for (int e = 0; e < c.M; ++e) {     /* back to the caller's observation order */
    const size_t i = (size_t)c.orig[e];
    if (jc) memcpy(jc + 12 * i, c.Jc + 12 * (size_t)e, sizeof(double) * 12);
    if (jp) memcpy(jp + 6 * i, c.Jp + 6 * (size_t)e, sizeof(double) * 6);
    if (jk) memcpy(jk + 6 * i, c.Jk + 6 * (size_t)e, sizeof(double) * 6);
  }
  ctx_free(&c);
  return PSFM_OK;
}

int psfm_oracle_ba_linear_step(const psfm_ba_problem* pb, const psfm_ba_options* opts, double radius,
                               int32_t solver, double* step_cam, double* step_pts,
                               int32_t* num_linear_iterations, int32_t num_threads) {
  psfm_ba_options o;
  if (opts) o = *opts; else psfm_oracle_ba_default_options(&o);
  ctx_t c;
  int rc = ctx_init(&c, pb, &o, num_threads);
  if (rc != PSFM_OK) { ctx_free(&c); return rc; }
  evaluate_gradient_and_jacobian(&c, 0);
  lm_diagonal(&c, radius, 0);
  if (solver == PSFM_BA_SOLVER_AUTO) solver = resolve_solver(&c);
  const int ls = linear_solve(&c, solver);
  if (step_cam) memcpy(step_cam, c.step_c, sizeof(double) * (size_t)c.NS);
  if (step_pts) memcpy(step_pts, c.step_p, sizeof(double) * 3 * (size_t)c.P);
  if (num_linear_iterations) *num_linear_iterations = c.num_linear_iterations;
  ctx_free(&c);
  return ls == 0 ? PSFM_OK : PSFM_ERR_INVALID;
}

/* ---------------------------------------------------------------- option defaults
   (duplicated from the product library on purpose: the oracle must not link it) */
void psfm_oracle_ba_default_options(psfm_ba_options* o) {
  memset(o, 0, sizeof(*o));
  o->loss_function_type = PSFM_LOSS_TRIVIAL;      /* bundle_adjustment.h:51 */
  o->loss_function_scale = 1.0;                   /* :54 */
  o->refine_focal_length = 1;                     /* :57 */
  o->refine_principal_point = 0;                  /* :60 */
  o->refine_extra_params = 1;                     /* :63 */
  o->refine_extrinsics = 1;                       /* :66 */
  o->refine_rotation = 1;                         /* :69 */
  o->print_summary = 1;                           /* :72 */
  o->minimizer_progress_to_stdout = 0;            /* :86 */
  o->function_tolerance = 0.0;                    /* :83 */
  o->gradient_tolerance = 0.0;                    /* :84 */
  o->parameter_tolerance = 0.0;                   /* :85 */
  o->max_num_iterations = 100;                    /* :87 */
  o->max_linear_solver_iterations = 200;          /* :88 */
  o->max_num_consecutive_invalid_steps = 10;      /* :89 */
  o->linear_solver = PSFM_BA_SOLVER_AUTO;
  o->eta = CERES_ETA;
  o->exact_r_tolerance = 1e-13;
  o->exact_max_iterations = 0;
  o->initial_trust_region_radius = CERES_INITIAL_TRUST_REGION_RADIUS;
  o->max_trust_region_radius = CERES_MAX_TRUST_REGION_RADIUS;
  o->min_trust_region_radius = CERES_MIN_TRUST_REGION_RADIUS;
  o->min_relative_decrease = CERES_MIN_RELATIVE_DECREASE;
  o->min_lm_diagonal = CERES_MIN_LM_DIAGONAL;
  o->max_lm_diagonal = CERES_MAX_LM_DIAGONAL;
  o->jacobi_scaling = 1;
  o->pcg_check_period = 0;
}

void psfm_oracle_ba_global_options(psfm_ba_options* o) {
  /* GlobalMapperOptions::GlobalBundleAdjustment, controllers/global_mapper.cc:41-71 */
  psfm_oracle_ba_default_options(o);
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1.0;
  o->parameter_tolerance = 1e-8;
  o->max_num_iterations = 50;
  o->max_linear_solver_iterations = 100;
  o->minimizer_progress_to_stdout = 1;
  o->print_summary = 1;
  o->refine_rotation = 0;
  o->refine_focal_length = 0;
  o->refine_principal_point = 0;
  o->refine_extra_params = 0;
  o->loss_function_type = PSFM_LOSS_SOFT_L1;
}
