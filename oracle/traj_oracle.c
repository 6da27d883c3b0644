/*
 * traj_oracle.c — CPU ORACLE for HP1 (path-consistency trajectory optimiser).
 * TEST INFRASTRUCTURE.
 *
 * Restates particlesfm::optimize_location
 *   (reference point_trajectory/optimize/src/trajectory_optimize.cpp:30-96):
 *   one Ceres problem with N residual blocks of 6 residuals / 4 parameters
 *   (PathConsistencyError, path_consistency_cost.h:42-59), TrivialLoss,
 *   SPARSE_NORMAL_CHOLESKY, DOGLEG, max 200 iterations, default tolerances;
 *   bilinear flow lookup = ceres::BiLinearInterpolator of linear_interpolation.h:97-123
 *   over ceres::Grid2D<double,2> (row-major, interleaved, index-clamped).
 * and Ceres 2.0.0's TrustRegionMinimizer + DoglegStrategy(TRADITIONAL) — see
 * SURVEY.md Appendix A.1/A.3 and ceres_semantics.h.  Because the normal matrix is
 * block diagonal (N independent 4x4 SPD blocks) the sparse Cholesky is N dense 4x4
 * Cholesky factorisations; everything else (radius, mu, rho, tolerances, the dogleg
 * interpolation) is GLOBAL over the 4N-vector, exactly as in the reference.
 *
 * PARITY UNPINNED (see psfm_oracle.h).
 *
 * Global sums use the "canonical sum" (DESIGN.md): 256-wide chunks reduced by a
 * butterfly tree per 32 + a left-to-right sum of the 8 group sums, then the chunk
 * partials are combined as tree256(q), q[t] = p[t] + p[t+256] + ... .  The CUDA kernel
 * implements the same definition, and this file is compiled with -ffp-contract=off, so
 * the two produce bit-identical iterates (the reference's integer track connectivity
 * depends on thresholds applied to these doubles — SURVEY.md §7 "hard parts").
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "psfm_oracle.h"
#include "ceres_semantics.h"

static double wall_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------- canonical sum */

static double tree256(const double* w) {
  double s[8];
  for (int g = 0; g < 8; ++g) {
    double a[32];
    memcpy(a, w + 32 * g, sizeof(a));
    for (int off = 16; off >= 1; off >>= 1)
      for (int i = 0; i < off; ++i) a[i] = a[i] + a[i + off];
    s[g] = a[0];
  }
  double t = s[0];
  for (int g = 1; g < 8; ++g) t = t + s[g];
  return t;
}

static double canon_sum(const double* v, int n, int mode, int nthreads) {
  if (mode == 1) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    return s;
  }
  const int nchunks = (n + 255) / 256;
  double* p = (double*)malloc(sizeof(double) * (size_t)(nchunks > 0 ? nchunks : 1));
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int c = 0; c < nchunks; ++c) {
    double w[256];
    const int base = 256 * c;
    for (int t = 0; t < 256; ++t) w[t] = (base + t < n) ? v[base + t] : 0.0;
    p[c] = tree256(w);
  }
  double q[256];
  for (int t = 0; t < 256; ++t) {
    double a = 0.0;
    for (int k = t; k < nchunks; k += 256) a = a + p[k];
    q[t] = a;
  }
  free(p);
  return tree256(q);
}

/* ------------------------------------------------------------- interpolation */

/* ceres::Grid2D<double,2>::GetValue (row-major, interleaved, clamped);
   the f32 map widened to f64 is exact. */
static inline void grid_get(const float* flow, int w, int h, int r, int c, double* f) {
  const int ri = r < 0 ? 0 : (r > h - 1 ? h - 1 : r);
  const int ci = c < 0 ? 0 : (c > w - 1 ? w - 1 : c);
  const float* px = flow + ((size_t)ri * w + ci) * 2;
  f[0] = (double)px[0];
  f[1] = (double)px[1];
}

/* BiLinearInterpolator::Evaluate (linear_interpolation.h:97-123) */
void psfm_oracle_bilinear(const float* flow, int32_t w, int32_t h, double r, double c, double* f,
                          double* dfdr, double* dfdc) {
  const int row = (int)floor(r);
  const int col = (int)floor(c);
  double p0[2], p1[2], f0[2], f1[2], d0[2], d1[2];
  const double xc = c - col, xr = r - row;
  grid_get(flow, w, h, row, col, p0);
  grid_get(flow, w, h, row, col + 1, p1);
  for (int k = 0; k < 2; ++k) { f0[k] = (1 - xc) * p0[k] + xc * p1[k]; d0[k] = p1[k] - p0[k]; }
  grid_get(flow, w, h, row + 1, col, p0);
  grid_get(flow, w, h, row + 1, col + 1, p1);
  for (int k = 0; k < 2; ++k) { f1[k] = (1 - xc) * p0[k] + xc * p1[k]; d1[k] = p1[k] - p0[k]; }
  for (int k = 0; k < 2; ++k) {
    f[k] = (1 - xr) * f0[k] + xr * f1[k];
    if (dfdr) dfdr[k] = f1[k] - f0[k];
    if (dfdc) dfdc[k] = (1 - xr) * d0[k] + xr * d1[k];
  }
}

/* ------------------------------------------------------------- the problem */

typedef struct {
  int n, w, h, mode, nthreads;
  const double *ref1, *ref2, *scale;
  const float* flow;
  double* x;    /* 4n current */
  double* xc;   /* 4n candidate */
  double* r;    /* 6n residuals at x */
  double* jac;  /* 4n: a,b,c,d of rows 4,5 (unscaled) */
  double* sc;   /* 4n jacobi scaling */
  double* diag; /* 4n dogleg diagonal */
  double* gt;   /* 4n scaled gradient g~ = (Js' r)/diag */
  double* gn;   /* 4n gauss-newton step (D-scaled space) */
  double* tmp;  /* 8n scratch for sums */
} tctx;

/* residuals + the four non-trivial Jacobian entries at xx (path_consistency_cost.h:42-59) */
static inline void eval_block(const tctx* c, int i, const double* xx, double* r, double* jac) {
  double ft[2], dr[2], dc[2];
  psfm_oracle_bilinear(c->flow, c->w, c->h, xx[1], xx[0], ft, dr, dc);
  const double s = c->scale[i];
  r[0] = xx[0] - c->ref1[2 * i];
  r[1] = xx[1] - c->ref1[2 * i + 1];
  r[2] = (xx[2] - c->ref2[2 * i]) * s;
  r[3] = (xx[3] - c->ref2[2 * i + 1]) * s;
  r[4] = (xx[2] - xx[0]) - ft[0];
  r[5] = (xx[3] - xx[1]) - ft[1];
  if (jac) {
    jac[0] = -1.0 - dc[0]; /* d r4 / d x1 */
    jac[1] = 0.0 - dr[0];  /* d r4 / d y1 */
    jac[2] = 0.0 - dc[1];  /* d r5 / d x1 */
    jac[3] = -1.0 - dr[1]; /* d r5 / d y1 */
  }
}

/* scaled Jacobian entries of block i */
typedef struct { double e0, e1, e2, e3, A, B, C, D, f2, f3; } sjac;
static inline void scaled_jac(const tctx* c, int i, sjac* J) {
  const double* sc = c->sc + 4 * i;
  const double* j = c->jac + 4 * i;
  const double s = c->scale[i];
  J->e0 = sc[0]; J->e1 = sc[1]; J->e2 = s * sc[2]; J->e3 = s * sc[3];
  J->A = j[0] * sc[0]; J->B = j[1] * sc[1]; J->C = j[2] * sc[0]; J->D = j[3] * sc[1];
  J->f2 = sc[2]; J->f3 = sc[3];
}

int psfm_oracle_traj_evaluate(const double* uv12, const double* ref1, const double* ref2,
                              const double* scale, const float* flow12, int32_t n, int32_t w,
                              int32_t h, double* residuals, double* jacobians) {
  tctx c;
  memset(&c, 0, sizeof(c));
  c.n = n; c.w = w; c.h = h; c.ref1 = ref1; c.ref2 = ref2; c.scale = scale; c.flow = flow12;
  for (int i = 0; i < n; ++i) {
    double r[6], j[4];
    eval_block(&c, i, uv12 + 4 * i, r, j);
    if (residuals) memcpy(residuals + 6 * i, r, sizeof(r));
    if (jacobians) {
      double* J = jacobians + 24 * (size_t)i;
      memset(J, 0, sizeof(double) * 24);
      J[0] = 1.0; J[5] = 1.0; J[10] = scale[i]; J[15] = scale[i];
      J[16] = j[0]; J[17] = j[1]; J[18] = 1.0;
      J[20] = j[2]; J[21] = j[3]; J[23] = 1.0;
    }
  }
  return PSFM_OK;
}

void psfm_oracle_traj_default_options(psfm_traj_options* o);

int psfm_oracle_traj_optimize(const double* uv12, const double* ref1, const double* ref2,
                              const double* scale, const float* flow12, int32_t n, int32_t w,
                              int32_t h, const psfm_traj_options* opts, double* out_uv12,
                              psfm_traj_summary* summary, int32_t num_threads,
                              int32_t reduction_mode) {
  psfm_traj_options o;
  if (opts) o = *opts; else psfm_oracle_traj_default_options(&o);
  psfm_traj_summary S;
  memset(&S, 0, sizeof(S));
  if (n <= 0) { if (summary) *summary = S; return PSFM_OK; }
  const double t0 = wall_s();
  tctx c;
  memset(&c, 0, sizeof(c));
  c.n = n; c.w = w; c.h = h; c.mode = reduction_mode;
#ifdef _OPENMP
  c.nthreads = num_threads > 0 ? num_threads : omp_get_max_threads();
#else
  c.nthreads = 1;
#endif
  c.ref1 = ref1; c.ref2 = ref2; c.scale = scale; c.flow = flow12;
  const size_t n4 = 4 * (size_t)n;
  c.x = (double*)malloc(sizeof(double) * n4);
  c.xc = (double*)malloc(sizeof(double) * n4);
  c.r = (double*)malloc(sizeof(double) * 6 * (size_t)n);
  c.jac = (double*)malloc(sizeof(double) * n4);
  c.sc = (double*)malloc(sizeof(double) * n4);
  c.diag = (double*)malloc(sizeof(double) * n4);
  c.gt = (double*)malloc(sizeof(double) * n4);
  c.gn = (double*)malloc(sizeof(double) * n4);
  c.tmp = (double*)malloc(sizeof(double) * 8 * (size_t)n);
  memcpy(c.x, uv12, sizeof(double) * n4);
  for (size_t k = 0; k < n4; ++k) c.sc[k] = 1.0;
  double* T0 = c.tmp; double* T1 = c.tmp + n; double* T2 = c.tmp + 2 * (size_t)n;
  double* T3 = c.tmp + 3 * (size_t)n; double* T4 = c.tmp + 4 * (size_t)n;
  const int nt = c.nthreads, md = c.mode;

  double radius = o.initial_trust_region_radius;
  double mu = CERES_DOGLEG_MIN_MU;
  int reuse = 0;
  double dogleg_step_norm = 0.0;
  double alpha = 0.0, gt_norm = 0.0, gn_norm = 0.0, gt_dot_gn = 0.0;
  int num_invalid = 0;
  int iteration = 0;
  int term = PSFM_TERM_NO_CONVERGENCE;
  double x_cost = 0.0, x_norm = 0.0, gmax = 0.0;
  int need_eval = 1;      /* evaluate r, J, gradient at x (iteration 0 / after acceptance) */
  int step_ok_prev = 1;
  int gn_valid = 0;

  for (;;) {
    if (need_eval) {
      /* EvaluateGradientAndJacobian at x */
      double gm = 0.0;
#pragma omp parallel for schedule(static) reduction(max : gm) num_threads(nt)
      for (int i = 0; i < n; ++i) {
        double* r = c.r + 6 * (size_t)i;
        double* j = c.jac + 4 * (size_t)i;
        const double* x = c.x + 4 * (size_t)i;
        eval_block(&c, i, x, r, j);
        const double s = c.scale[i];
        T0[i] = 0.5 * (((((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]) + r[3] * r[3]) + r[4] * r[4]) + r[5] * r[5]);
        T1[i] = ((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]) + x[3] * x[3];
        /* unscaled gradient J' r */
        const double g0 = (r[0] + j[0] * r[4]) + j[2] * r[5];
        const double g1 = (r[1] + j[1] * r[4]) + j[3] * r[5];
        const double g2 = s * r[2] + r[4];
        const double g3 = s * r[3] + r[5];
        gm = fmax(gm, fmax(fmax(fabs(g0), fabs(g1)), fmax(fabs(g2), fabs(g3))));
        if (iteration == 0 && o.jacobi_scaling) {
          /* jacobian_scaling = 1 / (1 + sqrt(squared column norm)) */
          double* sc = c.sc + 4 * (size_t)i;
          sc[0] = 1.0 / (1.0 + sqrt((1.0 + j[0] * j[0]) + j[2] * j[2]));
          sc[1] = 1.0 / (1.0 + sqrt((1.0 + j[1] * j[1]) + j[3] * j[3]));
          sc[2] = 1.0 / (1.0 + sqrt(s * s + 1.0));
          sc[3] = 1.0 / (1.0 + sqrt(s * s + 1.0));
        }
      }
      x_cost = canon_sum(T0, n, md, nt);
      x_norm = sqrt(canon_sum(T1, n, md, nt));
      gmax = gm;
      if (iteration == 0) S.initial_cost = x_cost;
      need_eval = 0;
      reuse = 0;
    }
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (step_ok_prev && iteration > 0) S.num_successful_steps++;
    if (iteration >= o.max_num_iterations) { term = PSFM_TERM_NO_CONVERGENCE; break; }
    if (gmax <= o.gradient_tolerance) { term = PSFM_TERM_CONVERGENCE_GRADIENT; break; }
    if (radius <= o.min_trust_region_radius) { term = PSFM_TERM_MIN_RADIUS; break; }
    ++iteration;
    step_ok_prev = 0;

    /* ---- DoglegStrategy::ComputeStep ---- */
    int ls_failed = 0;
    if (!reuse) {
      reuse = 1;
      /* diagonal_, gradient_, Cauchy point */
#pragma omp parallel for schedule(static) num_threads(nt)
      for (int i = 0; i < n; ++i) {
        sjac J;
        scaled_jac(&c, i, &J);
        const double* r = c.r + 6 * (size_t)i;
        double* dg = c.diag + 4 * (size_t)i;
        double* gt = c.gt + 4 * (size_t)i;
        const double h00 = (J.e0 * J.e0 + J.A * J.A) + J.C * J.C;
        const double h11 = (J.e1 * J.e1 + J.B * J.B) + J.D * J.D;
        const double h22 = J.e2 * J.e2 + J.f2 * J.f2;
        const double h33 = J.e3 * J.e3 + J.f3 * J.f3;
        dg[0] = sqrt(fmin(fmax(h00, CERES_MIN_LM_DIAGONAL), CERES_MAX_LM_DIAGONAL));
        dg[1] = sqrt(fmin(fmax(h11, CERES_MIN_LM_DIAGONAL), CERES_MAX_LM_DIAGONAL));
        dg[2] = sqrt(fmin(fmax(h22, CERES_MIN_LM_DIAGONAL), CERES_MAX_LM_DIAGONAL));
        dg[3] = sqrt(fmin(fmax(h33, CERES_MIN_LM_DIAGONAL), CERES_MAX_LM_DIAGONAL));
        const double gs0 = (J.e0 * r[0] + J.A * r[4]) + J.C * r[5];
        const double gs1 = (J.e1 * r[1] + J.B * r[4]) + J.D * r[5];
        const double gs2 = J.e2 * r[2] + J.f2 * r[4];
        const double gs3 = J.e3 * r[3] + J.f3 * r[5];
        gt[0] = gs0 / dg[0]; gt[1] = gs1 / dg[1]; gt[2] = gs2 / dg[2]; gt[3] = gs3 / dg[3];
        T0[i] = ((gt[0] * gt[0] + gt[1] * gt[1]) + gt[2] * gt[2]) + gt[3] * gt[3];
        /* Jg = Js * (g~ / diag) */
        const double v0 = gt[0] / dg[0], v1 = gt[1] / dg[1], v2 = gt[2] / dg[2], v3 = gt[3] / dg[3];
        const double m0 = J.e0 * v0, m1 = J.e1 * v1, m2 = J.e2 * v2, m3 = J.e3 * v3;
        const double m4 = (J.A * v0 + J.B * v1) + J.f2 * v2;
        const double m5 = (J.C * v0 + J.D * v1) + J.f3 * v3;
        T1[i] = ((((m0 * m0 + m1 * m1) + m2 * m2) + m3 * m3) + m4 * m4) + m5 * m5;
      }
      const double gt2 = canon_sum(T0, n, md, nt);
      const double jg2 = canon_sum(T1, n, md, nt);
      gt_norm = sqrt(gt2);
      alpha = gt2 / jg2;
      /* ComputeGaussNewtonStep */
      ls_failed = 1;
      gn_valid = 0;
      while (mu < CERES_DOGLEG_MAX_MU) {
        int fail = 0;
        const double sqmu = sqrt(mu);
#pragma omp parallel for schedule(static) reduction(| : fail) num_threads(nt)
        for (int i = 0; i < n; ++i) {
          sjac J;
          scaled_jac(&c, i, &J);
          const double* r = c.r + 6 * (size_t)i;
          const double* dg = c.diag + 4 * (size_t)i;
          double* gn = c.gn + 4 * (size_t)i;
          const double* gt = c.gt + 4 * (size_t)i;
          const double l0 = dg[0] * sqmu, l1 = dg[1] * sqmu, l2 = dg[2] * sqmu, l3 = dg[3] * sqmu;
          /* H = Js'Js + diag(lm^2), upper part */
          const double h00 = ((J.e0 * J.e0 + J.A * J.A) + J.C * J.C) + l0 * l0;
          const double h01 = J.A * J.B + J.C * J.D;
          const double h02 = J.A * J.f2;
          const double h03 = J.C * J.f3;
          const double h11 = ((J.e1 * J.e1 + J.B * J.B) + J.D * J.D) + l1 * l1;
          const double h12 = J.B * J.f2;
          const double h13 = J.D * J.f3;
          const double h22 = (J.e2 * J.e2 + J.f2 * J.f2) + l2 * l2;
          const double h33 = (J.e3 * J.e3 + J.f3 * J.f3) + l3 * l3;
          const double b0 = (J.e0 * r[0] + J.A * r[4]) + J.C * r[5];
          const double b1 = (J.e1 * r[1] + J.B * r[4]) + J.D * r[5];
          const double b2 = J.e2 * r[2] + J.f2 * r[4];
          const double b3 = J.e3 * r[3] + J.f3 * r[5];
          /* 4x4 Cholesky H = L L' (h23 = 0) */
          int bad = 0;
          double L00, L10, L20, L30, L11, L21, L31, L22, L32, L33, d;
          if (!(h00 > 0.0)) bad = 1;
          L00 = sqrt(h00);
          L10 = h01 / L00; L20 = h02 / L00; L30 = h03 / L00;
          d = h11 - L10 * L10; if (!(d > 0.0)) bad = 1;
          L11 = sqrt(d);
          L21 = (h12 - L20 * L10) / L11;
          L31 = (h13 - L30 * L10) / L11;
          d = (h22 - L20 * L20) - L21 * L21; if (!(d > 0.0)) bad = 1;
          L22 = sqrt(d);
          L32 = ((0.0 - L30 * L20) - L31 * L21) / L22;
          d = ((h33 - L30 * L30) - L31 * L31) - L32 * L32; if (!(d > 0.0)) bad = 1;
          L33 = sqrt(d);
          /* forward / backward substitution */
          const double y0 = b0 / L00;
          const double y1 = (b1 - L10 * y0) / L11;
          const double y2 = ((b2 - L20 * y0) - L21 * y1) / L22;
          const double y3 = (((b3 - L30 * y0) - L31 * y1) - L32 * y2) / L33;
          const double z3 = y3 / L33;
          const double z2 = (y2 - L32 * z3) / L22;
          const double z1 = ((y1 - L21 * z2) - L31 * z3) / L11;
          const double z0 = (((y0 - L10 * z1) - L20 * z2) - L30 * z3) / L00;
          if (!isfinite(z0) || !isfinite(z1) || !isfinite(z2) || !isfinite(z3)) bad = 1;
          fail |= bad;
          /* gauss_newton_step_ *= -diagonal_ */
          gn[0] = z0 * (-dg[0]); gn[1] = z1 * (-dg[1]); gn[2] = z2 * (-dg[2]); gn[3] = z3 * (-dg[3]);
          T0[i] = ((gn[0] * gn[0] + gn[1] * gn[1]) + gn[2] * gn[2]) + gn[3] * gn[3];
          T1[i] = ((gt[0] * gn[0] + gt[1] * gn[1]) + gt[2] * gn[2]) + gt[3] * gn[3];
        }
        if (fail) { mu *= CERES_DOGLEG_MU_INCREASE; continue; }
        gn_norm = sqrt(canon_sum(T0, n, md, nt));
        gt_dot_gn = canon_sum(T1, n, md, nt);
        ls_failed = 0;
        gn_valid = 1;
        break;
      }
    }
    double mcc = 0.0, cand_cost = 0.0, step_sq = 0.0;
    int valid = 0;
    if (!ls_failed && gn_valid) {
      /* ComputeTraditionalDoglegStep, candidate, model cost change */
      int kase;
      double beta = 0.0, cauchy_scale = 0.0;
      if (gn_norm <= radius) kase = 1;
      else if (gt_norm * alpha >= radius) { kase = 2; cauchy_scale = -(radius / gt_norm); }
      else {
        kase = 3;
        const double b_dot_a = -alpha * gt_dot_gn;
        const double an = alpha * gt_norm;
        const double a_sq = an * an;
        const double bma_sq = (a_sq - 2 * b_dot_a) + gn_norm * gn_norm;
        const double cc = b_dot_a - a_sq;
        const double dd = sqrt(cc * cc + bma_sq * (radius * radius - a_sq));
        beta = (cc <= 0) ? (dd - cc) / bma_sq : (radius * radius - a_sq) / (dd + cc);
        cauchy_scale = -alpha * (1.0 - beta);
      }
#pragma omp parallel for schedule(static) num_threads(nt)
      for (int i = 0; i < n; ++i) {
        sjac J;
        scaled_jac(&c, i, &J);
        const double* r = c.r + 6 * (size_t)i;
        const double* dg = c.diag + 4 * (size_t)i;
        const double* gt = c.gt + 4 * (size_t)i;
        const double* gn = c.gn + 4 * (size_t)i;
        const double* sc = c.sc + 4 * (size_t)i;
        const double* x = c.x + 4 * (size_t)i;
        double* xc = c.xc + 4 * (size_t)i;
        double s0, s1, s2, s3;
        if (kase == 1) { s0 = gn[0]; s1 = gn[1]; s2 = gn[2]; s3 = gn[3]; }
        else if (kase == 2) { s0 = cauchy_scale * gt[0]; s1 = cauchy_scale * gt[1]; s2 = cauchy_scale * gt[2]; s3 = cauchy_scale * gt[3]; }
        else {
          s0 = cauchy_scale * gt[0] + beta * gn[0]; s1 = cauchy_scale * gt[1] + beta * gn[1];
          s2 = cauchy_scale * gt[2] + beta * gn[2]; s3 = cauchy_scale * gt[3] + beta * gn[3];
        }
        T3[i] = ((s0 * s0 + s1 * s1) + s2 * s2) + s3 * s3;
        const double p0 = s0 / dg[0], p1 = s1 / dg[1], p2 = s2 / dg[2], p3 = s3 / dg[3];
        /* model residuals m = Js * step */
        const double m0 = J.e0 * p0, m1 = J.e1 * p1, m2 = J.e2 * p2, m3 = J.e3 * p3;
        const double m4 = (J.A * p0 + J.B * p1) + J.f2 * p2;
        const double m5 = (J.C * p0 + J.D * p1) + J.f3 * p3;
        T0[i] = ((((m0 * (r[0] + m0 / 2.0) + m1 * (r[1] + m1 / 2.0)) + m2 * (r[2] + m2 / 2.0)) +
                  m3 * (r[3] + m3 / 2.0)) + m4 * (r[4] + m4 / 2.0)) + m5 * (r[5] + m5 / 2.0);
        /* delta = step o scale ; candidate = x + delta */
        xc[0] = x[0] + p0 * sc[0]; xc[1] = x[1] + p1 * sc[1];
        xc[2] = x[2] + p2 * sc[2]; xc[3] = x[3] + p3 * sc[3];
        double rc[6];
        eval_block(&c, i, xc, rc, NULL);
        T1[i] = 0.5 * (((((rc[0] * rc[0] + rc[1] * rc[1]) + rc[2] * rc[2]) + rc[3] * rc[3]) + rc[4] * rc[4]) + rc[5] * rc[5]);
        const double e0 = x[0] - xc[0], e1 = x[1] - xc[1], e2 = x[2] - xc[2], e3 = x[3] - xc[3];
        T2[i] = ((e0 * e0 + e1 * e1) + e2 * e2) + e3 * e3;
      }
      mcc = -canon_sum(T0, n, md, nt);
      cand_cost = canon_sum(T1, n, md, nt);
      step_sq = canon_sum(T2, n, md, nt);
      if (kase == 1) dogleg_step_norm = gn_norm;
      else if (kase == 2) dogleg_step_norm = radius;
      else dogleg_step_norm = sqrt(canon_sum(T3, n, md, nt));
      valid = mcc > 0.0;
    }
    (void)T4;
    if (!valid) {
      /* HandleInvalidStep -> DoglegStrategy::StepIsInvalid */
      S.num_unsuccessful_steps++;
      if (++num_invalid >= o.max_num_consecutive_invalid_steps) { term = PSFM_TERM_FAILURE; break; }
      mu *= CERES_DOGLEG_MU_INCREASE;
      reuse = 0;
      continue;
    }
    num_invalid = 0;
    const double step_norm = sqrt(step_sq);
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { term = PSFM_TERM_CONVERGENCE_PARAMETER; break; }
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * x_cost) { term = PSFM_TERM_CONVERGENCE_FUNCTION; break; }
    const double rho = cost_change / mcc;
    if (rho > o.min_relative_decrease) {
      double* t = c.x; c.x = c.xc; c.xc = t;
      need_eval = 1;
      step_ok_prev = 1;
      /* DoglegStrategy::StepAccepted */
      if (rho < CERES_DOGLEG_DECREASE_THRESHOLD) radius *= 0.5;
      if (rho > CERES_DOGLEG_INCREASE_THRESHOLD) radius = fmax(radius, 3.0 * dogleg_step_norm);
      mu = fmax(CERES_DOGLEG_MIN_MU, 2.0 * mu / CERES_DOGLEG_MU_INCREASE);
      reuse = 0;
    } else {
      S.num_unsuccessful_steps++;
      radius *= 0.5;
      reuse = 1;
    }
  }
  S.num_iterations = iteration;
  S.termination = term;
  S.final_cost = x_cost;
  memcpy(out_uv12, c.x, sizeof(double) * n4);
  S.total_ms = 1e3 * (wall_s() - t0);
  S.solve_ms = S.total_ms;
  if (summary) *summary = S;
  free(c.x); free(c.xc); free(c.r); free(c.jac); free(c.sc); free(c.diag); free(c.gt); free(c.gn); free(c.tmp);
  return PSFM_OK;
}

void psfm_oracle_traj_default_options(psfm_traj_options* o) {
  o->max_num_iterations = 200;                 /* trajectory_optimize.cpp:76 */
  o->function_tolerance = CERES_FUNCTION_TOLERANCE;
  o->gradient_tolerance = CERES_GRADIENT_TOLERANCE;
  o->parameter_tolerance = CERES_PARAMETER_TOLERANCE;
  o->initial_trust_region_radius = CERES_INITIAL_TRUST_REGION_RADIUS;
  o->max_trust_region_radius = CERES_MAX_TRUST_REGION_RADIUS;
  o->min_trust_region_radius = CERES_MIN_TRUST_REGION_RADIUS;
  o->min_relative_decrease = CERES_MIN_RELATIVE_DECREASE;
  o->max_num_consecutive_invalid_steps = CERES_MAX_NUM_CONSECUTIVE_INVALID_STEPS;
  o->jacobi_scaling = 1;
}
