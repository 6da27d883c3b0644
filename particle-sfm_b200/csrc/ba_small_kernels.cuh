// ba_small_kernels.cuh — O(#images) kernels of HP2: Jacobi scaling, LM diagonal and
// Schur-Jacobi blocks of the reduced camera system, the PCG vector updates (device
// resident scalars, no host round trip per iteration), manifold Plus() for the poses.
// Semantics follow Ceres 2.0.0.
#pragma once
#include "ba_kernels.cuh"

namespace psfm {
namespace ba {

// ------------------------------------------------------------------ Jacobi scaling

// jacobian_scaling = 1 / (1 + sqrt(squared column norm)), computed once at iteration 0
// (TrustRegionMinimizer::EvaluateGradientAndJacobian); inactive slots keep scale 0.
__global__ void k_scale_cams(const double* acc_cam, const double* acc_intr, const unsigned char* active,
                             int F, int C, double* scale_c) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = 6 * F + 3 * C;
  if (s >= NS) return;
  double n2;
  if (s < 6 * F) {
    const int img = s / 6, k = s % 6;
    const int d[3] = {0, 3, 5};
    n2 = acc_cam[(size_t)img * NVL + (k < 3 ? d[k] : 6 + d[k - 3])];
  } else {
    const int c = (s - 6 * F) / 3, k = (s - 6 * F) % 3;
    const int d[3] = {0, 3, 5};
    n2 = acc_intr[(size_t)c * NVI + d[k]];
  }
  scale_c[s] = active[s] ? 1.0 / (1.0 + sqrt(n2)) : 0.0;
}

__global__ void k_scale_points(const double* hpp, int P, double* scale_p) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  scale_p[3 * (size_t)p] = 1.0 / (1.0 + sqrt(hpp[p]));
  scale_p[3 * (size_t)p + 1] = 1.0 / (1.0 + sqrt(hpp[3 * (size_t)P + p]));
  scale_p[3 * (size_t)p + 2] = 1.0 / (1.0 + sqrt(hpp[5 * (size_t)P + p]));
}

__global__ void k_fill(double* p, double v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// dst[i] (+)= sum over replicas of src[r][i]; the replicas are zeroed for the next use
__global__ void k_fold_replicas(double* __restrict__ dst, double* __restrict__ rep, size_t n, size_t stride, int nrep,
                                const double* __restrict__ scale, const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int r0 = 0; r0 < nrep; r0 += 8) {        // eight independent loads in flight, fixed summation order
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (r0 + u < nrep) ? rep[(size_t)(r0 + u) * stride + i] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r0 + u < nrep) rep[(size_t)(r0 + u) * stride + i] = 0.0;
  }
  dst[i] = scale ? scale[i] * s : s;
}

// two folds in one launch (the camera-side sums and the band blocks of the fused Schur kernel)
__global__ void k_fold_replicas2(double* __restrict__ dst, double* __restrict__ rep_a, size_t na, int nrep_a, double* __restrict__ rep_b,
                                 size_t nb, int nrep_b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na + nb) return;
  double* rep = rep_a;
  size_t stride = na;
  int nrep = nrep_a;
  double* out = dst + i;
  if (i >= na) { i -= na; rep = rep_b; stride = nb; nrep = nrep_b; }
  double s = 0.0;
  for (int r0 = 0; r0 < nrep; r0 += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (r0 + u < nrep) ? rep[(size_t)(r0 + u) * stride + i] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r0 + u < nrep) rep[(size_t)(r0 + u) * stride + i] = 0.0;
  }
  *out = s;
}

// xs = s o x  (the tile kernels work with the unscaled Jacobian: J_scaled x = J (s o x))
__global__ void k_scale_vec(const double* x, const double* scale, size_t n, double* xs, const int* skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xs[i] = scale[i] * x[i];
}

// pose (q, t) -> pose16 rows: R (row-major 9), t (3), pad
__global__ void k_pose_table(const double* pose, int F, double* pose16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  const double* p = pose + 8 * (size_t)i;
  const double q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
  double* o = pose16 + 16 * (size_t)i;
  o[0] = 1.0 - 2.0 * (q2 * q2 + q3 * q3); o[1] = 2.0 * (q1 * q2 - q0 * q3); o[2] = 2.0 * (q1 * q3 + q0 * q2);
  o[3] = 2.0 * (q1 * q2 + q0 * q3); o[4] = 1.0 - 2.0 * (q1 * q1 + q3 * q3); o[5] = 2.0 * (q2 * q3 - q0 * q1);
  o[6] = 2.0 * (q1 * q3 - q0 * q2); o[7] = 2.0 * (q2 * q3 + q0 * q1); o[8] = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
  o[9] = p[4]; o[10] = p[5]; o[11] = p[6];
  o[12] = o[13] = o[14] = o[15] = 0.0;
}

// sum of squares of a vector (grid-stride), atomically added to *out
__global__ void __launch_bounds__(256) k_sqnorm(const double* v, size_t n, double* out) {
  __shared__ double sred[32];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += v[i] * v[i];
  s = block_sum(s, sred);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// ------------------------------------------------------------------ small dense helpers

// inverse of the SPD 3x3 block restricted to the active dims; identity elsewhere
__device__ inline void inv3_masked(const double* A /*upper 6*/, const unsigned char* act, double* Ai /*9*/) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[1], A[3], A[4]}, {A[2], A[4], A[5]}};
  for (int j = 0; j < 3; ++j)
    if (!act[j]) {
      for (int k = 0; k < 3; ++k) { a[j][k] = 0.0; a[k][j] = 0.0; }
      a[j][j] = 1.0;
    }
  // Cholesky of the (now block-diagonal w.r.t. inactive dims) SPD matrix
  const double l00 = sqrt(a[0][0]);
  const double l10 = a[1][0] / l00, l20 = a[2][0] / l00;
  const double l11 = sqrt(a[1][1] - l10 * l10);
  const double l21 = (a[2][1] - l20 * l10) / l11;
  const double l22 = sqrt(a[2][2] - l20 * l20 - l21 * l21);
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  Ai[0] = i00 * i00 + i10 * i10 + i20 * i20;
  Ai[1] = Ai[3] = i10 * i11 + i20 * i21;
  Ai[2] = Ai[6] = i20 * i22;
  Ai[4] = i11 * i11 + i21 * i21;
  Ai[5] = Ai[7] = i21 * i22;
  Ai[8] = i22 * i22;
}

__device__ inline void quat_mul(const double* a, const double* b, double* o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

// ceres::QuaternionParameterization::Plus
__device__ inline void quat_plus(const double* x, const double* d, double* o) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd;
    const double qd[4] = {cos(nd), s * d[0], s * d[1], s * d[2]};
    quat_mul(qd, x, o);
  } else {
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
  }
}

// ------------------------------------------------------------------ reduced-system set-up

struct CamFinArgs {
  const double* lin_cam;    // [F][NVL]  F'F blocks, F'r
  const double* lin_intr;   // [C][NVI]
  const double* prep_cam;   // [F][prep_stride]  -(W hinv W') blocks at 0 (when prep_blocks), -(W w) at prep_goff
  int prep_stride, prep_goff, prep_blocks;   // k_schur_prep: NVL, 12, 1;  k_schur_tile sums: NVX2, 21, 0
  const double* prep_intr;  // [C][NVI]
  const unsigned char* active;  // [NS]
  const double* scale_c;        // [NS] jacobi scaling (0 on inactive slots)
  double radius, min_diag, max_diag;
  int F, C;
  double* Dc2;    // [NS] LM diagonal squared (0 on inactive slots)
  double* Minv;   // [(2F + C)][9] inverse Schur-Jacobi blocks
  double* rhs;    // [NS] reduced right-hand side
};

// one thread per parameter block: rot(img), t(img), intr(cam)
__global__ void k_cam_finalize(const CamFinArgs a) {
  const int nb = blockIdx.x * blockDim.x + threadIdx.x;
  const int NB = 2 * a.F + a.C;
  if (nb >= NB) return;
  const double *A, *Cc, *g, *gc;
  int slot0;
  if (nb < 2 * a.F) {
    const int img = nb >> 1, b = nb & 1;
    A = a.lin_cam + (size_t)img * NVL + 6 * b;
    g = a.lin_cam + (size_t)img * NVL + 12 + 3 * b;
    Cc = a.prep_blocks ? a.prep_cam + (size_t)img * a.prep_stride + 6 * b : nullptr;
    gc = a.prep_cam + (size_t)img * a.prep_stride + a.prep_goff + 3 * b;
    slot0 = 6 * img + 3 * b;
  } else {
    const int c = nb - 2 * a.F;
    A = a.lin_intr + (size_t)c * NVI;
    g = A + 6;
    Cc = a.prep_intr + (size_t)c * NVI;
    gc = Cc + 6;
    slot0 = 6 * a.F + 3 * c;
  }
  const unsigned char* act = a.active + slot0;
  const double* sc = a.scale_c + slot0;
  const int dg[3] = {0, 3, 5};
  // the tile kernels accumulate with the UNSCALED Jacobian: apply diag(s) on both sides here
  const double ss[6] = {sc[0] * sc[0], sc[0] * sc[1], sc[0] * sc[2], sc[1] * sc[1], sc[1] * sc[2], sc[2] * sc[2]};
  double Mb[6];
  for (int k = 0; k < 6; ++k) Mb[k] = ss[k] * (A[k] + (Cc ? Cc[k] : 0.0));
  for (int j = 0; j < 3; ++j) {
    double d2 = 0.0;
    if (act[j]) d2 = fmin(fmax(ss[dg[j]] * A[dg[j]], a.min_diag), a.max_diag) / a.radius;
    a.Dc2[slot0 + j] = d2;
    Mb[dg[j]] += d2;
    a.rhs[slot0 + j] = act[j] ? sc[j] * (g[j] + gc[j]) : 0.0;
  }
  inv3_masked(Mb, act, a.Minv + (size_t)nb * 9);
}

// gradient_max_norm contribution of the camera-side blocks: |Plus(x, -g) - x|_inf with
// g = J' r the (unscaled) tangent gradient
__global__ void k_cam_gmax(const double* lin_cam, const double* lin_intr,
                           const unsigned char* active, const double* pose, int F, int C, double* gmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (j < F) {
    const double* g = lin_cam + (size_t)j * NVL + 12;
    const unsigned char* act = active + 6 * (size_t)j;
    if (act[0]) {
      const double d[3] = {-g[0], -g[1], -g[2]};
      const double* q = pose + 8 * (size_t)j;
      double qn[4];
      quat_plus(q, d, qn);
      for (int k = 0; k < 4; ++k) m = fmax(m, fabs(q[k] - qn[k]));
    }
    for (int k = 0; k < 3; ++k)
      if (act[3 + k]) m = fmax(m, fabs(g[3 + k]));
  } else if (j < F + C) {
    const int c = j - F;
    const double* g = lin_intr + (size_t)c * NVI + 6;
    for (int k = 0; k < 3; ++k)
      if (active[6 * (size_t)F + 3 * c + k]) m = fmax(m, fabs(g[k]));
  }
  if (m > 0.0) atomic_max_nonneg(gmax, m);
}

// ------------------------------------------------------------------ PCG (ConjugateGradientsSolver::Solve)

enum PcgFlag { PCG_RUNNING = 0, PCG_SUCCESS = 1, PCG_INDEFINITE = 2, PCG_FAILURE = 3, PCG_MAXITER = 4 };

struct PcgState {
  double rho, Q0, norm_b, norm_r;
  int it;          // iteration about to run / last run
  int flag;
  int pad0, pad1;
};

struct PcgArgs {
  PcgState* st;
  const double* b;      // rhs
  const double* Minv;   // [(2F+C)][9]
  const double* Dc2;    // [NS]
  double* x;
  double* r;
  double* p;
  double* z;
  double* y;            // S*p accumulator (without the D^2 term)
  int NS, NB, F;
  double q_tol, r_tol;  // r_tol < 0 disables the residual test (LM uses -1)
  int max_it, min_it;
};

__device__ inline void pcg_precondition(const PcgArgs& a) {
  // z = M^-1 r, block by block
  for (int nb = threadIdx.x; nb < a.NB; nb += blockDim.x) {
    const int s0 = (nb < 2 * a.F) ? 3 * nb : 6 * a.F + 3 * (nb - 2 * a.F);
    const double* Mi = a.Minv + (size_t)nb * 9;
    const double r0 = a.r[s0], r1 = a.r[s0 + 1], r2 = a.r[s0 + 2];
    a.z[s0] = Mi[0] * r0 + Mi[1] * r1 + Mi[2] * r2;
    a.z[s0 + 1] = Mi[3] * r0 + Mi[4] * r1 + Mi[5] * r2;
    a.z[s0 + 2] = Mi[6] * r0 + Mi[7] * r1 + Mi[8] * r2;
  }
  __syncthreads();
}

__device__ inline double pcg_dot(const double* u, const double* v, int n, double* sred) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += u[i] * v[i];
  __shared__ double bc;
  s = block_sum(s, sred);
  if (threadIdx.x == 0) bc = s;
  __syncthreads();
  s = bc;
  __syncthreads();
  return s;
}

__device__ inline bool zero_or_inf(double x) { return x == 0.0 || isinf(x) || isnan(x); }

// start of iteration `it`: z = M^-1 r ; rho = r.z ; p = z (+ beta p) ; y = 0
__device__ inline void pcg_next_direction(const PcgArgs& a, double* sred, bool first) {
  pcg_precondition(a);
  const double last_rho = a.st->rho;
  const double rho = pcg_dot(a.r, a.z, a.NS, sred);
  if (zero_or_inf(rho)) {
    if (threadIdx.x == 0) a.st->flag = PCG_FAILURE;
    return;
  }
  double beta = 0.0;
  if (!first) {
    beta = rho / last_rho;
    if (zero_or_inf(beta)) {
      if (threadIdx.x == 0) a.st->flag = PCG_FAILURE;
      return;
    }
  }
  for (int i = threadIdx.x; i < a.NS; i += blockDim.x) {
    a.p[i] = first ? a.z[i] : a.z[i] + beta * a.p[i];
    a.y[i] = 0.0;
  }
  if (threadIdx.x == 0) a.st->rho = rho;
}

__global__ void __launch_bounds__(1024) k_pcg_init(const PcgArgs a) {
  __shared__ double sred[32];
  for (int i = threadIdx.x; i < a.NS; i += blockDim.x) { a.x[i] = 0.0; a.r[i] = a.b[i]; a.y[i] = 0.0; }
  __syncthreads();
  const double nb2 = pcg_dot(a.b, a.b, a.NS, sred);
  if (threadIdx.x == 0) {
    a.st->norm_b = sqrt(nb2);
    a.st->norm_r = sqrt(nb2);
    a.st->rho = 1.0;
    a.st->Q0 = 0.0;   // -x.(b + r) with x = 0
    a.st->it = 1;
    a.st->flag = (nb2 == 0.0) ? PCG_SUCCESS : PCG_RUNNING;
  }
  __syncthreads();
  if (nb2 == 0.0) return;
  pcg_next_direction(a, sred, true);
}

// mode 0: normal iteration (A + B); mode 1: refresh iteration part A (x updated, y zeroed
// for the S*x product); mode 2: refresh part A2 + B (r = b - S x).
__global__ void __launch_bounds__(1024) k_pcg_update(const PcgArgs a, const int mode) {
  __shared__ double sred[32];
  if (a.st->flag != PCG_RUNNING) return;
  const int it = a.st->it;
  if (mode != 2) {
    // q = S p = y + D^2 p
    for (int i = threadIdx.x; i < a.NS; i += blockDim.x) a.y[i] += a.Dc2[i] * a.p[i];
    __syncthreads();
    const double pq = pcg_dot(a.p, a.y, a.NS, sred);
    if (pq <= 0.0 || isinf(pq) || isnan(pq)) {
      if (threadIdx.x == 0) a.st->flag = PCG_INDEFINITE;
      return;
    }
    const double alpha = a.st->rho / pq;
    if (isinf(alpha)) {
      if (threadIdx.x == 0) a.st->flag = PCG_FAILURE;
      return;
    }
    for (int i = threadIdx.x; i < a.NS; i += blockDim.x) {
      a.x[i] = a.x[i] + alpha * a.p[i];
      if (mode == 0) a.r[i] = a.r[i] - alpha * a.y[i];
      else a.y[i] = 0.0;
    }
    __syncthreads();
    if (mode == 1) return;
  } else {
    // r = b - S x  (every residual_reset_period iterations)
    for (int i = threadIdx.x; i < a.NS; i += blockDim.x) a.r[i] = a.b[i] - (a.y[i] + a.Dc2[i] * a.x[i]);
    __syncthreads();
  }
  // quadratic-model termination: Q = -x.(b + r)
  double s = 0.0;
  for (int i = threadIdx.x; i < a.NS; i += blockDim.x) s += a.x[i] * (a.b[i] + a.r[i]);
  __shared__ double bc;
  s = block_sum(s, sred);
  if (threadIdx.x == 0) bc = s;
  __syncthreads();
  const double Q1 = -1.0 * bc;
  __syncthreads();
  const double zeta = it * (Q1 - a.st->Q0) / Q1;
  const double nr2 = pcg_dot(a.r, a.r, a.NS, sred);
  const double norm_r = sqrt(nr2);
  int flag = PCG_RUNNING;
  if (zeta < a.q_tol && it >= a.min_it) flag = PCG_SUCCESS;
  else if (a.r_tol >= 0.0 && norm_r <= a.r_tol * a.st->norm_b && it >= a.min_it) flag = PCG_SUCCESS;
  else if (it >= a.max_it) flag = PCG_MAXITER;
  __syncthreads();
  if (threadIdx.x == 0) {
    a.st->Q0 = Q1;
    a.st->norm_r = norm_r;
    a.st->flag = flag;
    if (flag == PCG_RUNNING) a.st->it = it + 1;
  }
  __syncthreads();
  if (flag != PCG_RUNNING) return;
  pcg_next_direction(a, sred, false);
}

// ------------------------------------------------------------------ candidate poses / intrinsics

struct ApplyArgs {
  const double* yc;        // reduced-system solution; step = -yc (scaled space)
  const double* scale_c;
  const unsigned char* active;
  const double* pose;      // [F*8]
  const double* K;         // [C*3]
  double* pose_c;
  double* K_c;
  int F, C;
  double* acc;             // [0] |x - x_c|^2 (ambient)   [1] |x_c|^2 over non-constant blocks
};

__global__ void __launch_bounds__(256) k_apply_cams(const ApplyArgs a) {
  __shared__ double sred[64];
  const int j = blockIdx.x * 256 + threadIdx.x;
  double d2 = 0.0, x2 = 0.0;
  if (j < a.F) {
    const double* ps = a.pose + 8 * (size_t)j;
    double* pc = a.pose_c + 8 * (size_t)j;
    const unsigned char* act = a.active + 6 * (size_t)j;
    const double* y = a.yc + 6 * (size_t)j;
    const double* sc = a.scale_c + 6 * (size_t)j;
    for (int k = 0; k < 8; ++k) pc[k] = ps[k];
    if (act[0]) {
      const double d[3] = {-y[0] * sc[0], -y[1] * sc[1], -y[2] * sc[2]};
      double qn[4];
      quat_plus(ps, d, qn);
      for (int k = 0; k < 4; ++k) {
        pc[k] = qn[k];
        const double e = ps[k] - qn[k];
        d2 += e * e;
        x2 += qn[k] * qn[k];
      }
    }
    bool anyt = false;
    for (int k = 0; k < 3; ++k)
      if (act[3 + k]) {
        anyt = true;
        pc[4 + k] = ps[4 + k] + (-y[3 + k] * sc[3 + k]);
        const double e = ps[4 + k] - pc[4 + k];
        d2 += e * e;
      }
    if (anyt) for (int k = 0; k < 3; ++k) x2 += pc[4 + k] * pc[4 + k];
  } else if (j < a.F + a.C) {
    const int c = j - a.F;
    const size_t s0 = 6 * (size_t)a.F + 3 * c;
    bool any = false;
    for (int k = 0; k < 3; ++k) {
      double v = a.K[3 * c + k];
      if (a.active[s0 + k]) {
        any = true;
        const double nv = v + (-a.yc[s0 + k] * a.scale_c[s0 + k]);
        const double e = v - nv;
        d2 += e * e;
        v = nv;
      }
      a.K_c[3 * c + k] = v;
    }
    if (any) for (int k = 0; k < 3; ++k) x2 += a.K_c[3 * c + k] * a.K_c[3 * c + k];
  }
  double v[2] = {d2, x2};
  const double s = block_sum_multi<2>(v, sred);
  if (threadIdx.x < 2) atomicAdd(a.acc + threadIdx.x, s);
}

}  // namespace ba
}  // namespace psfm
