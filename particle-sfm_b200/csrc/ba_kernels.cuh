// ba_kernels.cuh — sm_100a kernels of HP2 (global bundle adjustment).
//
// Replaces what ceres::Solve does inside BundleAdjuster::Solve
// (reference sfm/gmapper/src/optim/bundle_adjustment.cc:306): per-observation
// reprojection residual/Jacobian evaluation (COLMAP BundleAdjustmentCostFunction<
// SimplePinholeCameraModel>, call sites bundle_adjustment.cc:380-411), loss correction
// (CreateLossFunction :54-69), per-point 3x3 Schur elimination and the implicit
// reduced-camera-system product used by PCG.
//
// Data layout: observations sorted by point and packed into TILES of
// <= TILE observations that never split a point; one CTA per tile, one thread per
// observation.  Per-point sums are reduced inside the tile through shared memory (no
// atomics, fixed order); per-image sums are reduced per (image-segment, component)
// inside the tile with a host-precomputed in-tile image ordering, then one fp64 RED per
// segment goes to the global accumulator.  Jacobians are stored SoA by component
// ([k][M]) so that every global access of a warp is a contiguous 256-byte run.
#pragma once
#include "psfm_common.cuh"

namespace psfm {
namespace ba {

constexpr int NVL = 18;   // camera-side accumulator stride per image
constexpr int NVI = 9;    // intrinsics accumulator stride per camera
// Concurrently resident tiles are consecutive tiles and therefore touch the SAME few
// images: their fp64 REDs would serialise on a handful of L2 addresses.  Every per-image
// accumulator is therefore replicated NREP times (replica = tile index mod NREP) and the
// replicas are folded by k_fold_replicas before use.
constexpr int NREP = 32;
constexpr double kHuge = 1.7976931348623157e308;

struct TileCtx {
  const int* tile_start;   // [T+1] first observation of each tile
  const int* tile_pt;      // [T+1] first (internal) point of each tile
  const int* pt_ptr;       // [P+1] observation range of each point
  const int* obs_img;      // [M]
  const int* obs_pt;       // [M] internal point id
  const double2* obs_xy;   // [M]
  const unsigned short* tile_perm;  // [M] e-th observation of the tile in image order
  const int* cseg_ptr;     // [T+1] image segments of each tile
  const int* cseg_img;     // [nseg]
  const unsigned short* cseg_off;  // [nseg] start of the segment in the tile's image order
  const int* img_cam;      // [F]
  int F, P, M, C, T;
};

struct Jac {        // SoA by component, index k*M + i
  double* r;        // [2][M] loss-corrected residuals
  double* jc;       // [12][M] (ROT) or [6][M]: row0 cols.., row1 cols.. of the pose block
  double* jp;       // [6][M]
  double* jk;       // [4][M] focal column (2) then sq*s_cx, sq*s_cy (principal point)
};

struct LossP {
  int type;
  double a;
};

// ceres::{Trivial,SoftLOne,Cauchy}Loss::Evaluate -> rho0, rho1 (rho2 <= 0 always, so the
// Corrector takes the first-order branch: residual and Jacobian scaled by sqrt(rho1)).
__device__ __forceinline__ void loss_eval(const LossP l, const double s, double& rho0, double& rho1) {
  if (l.type == PSFM_LOSS_SOFT_L1) {
    const double b = l.a * l.a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double tmp = sqrt(sum);
    rho0 = 2.0 * b * (tmp - 1.0);
    rho1 = fmax(2.2250738585072014e-308, 1.0 / tmp);
  } else if (l.type == PSFM_LOSS_CAUCHY) {
    const double b = l.a * l.a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    rho0 = b * log(sum);
    rho1 = fmax(2.2250738585072014e-308, 1.0 / sum);
  } else {
    rho0 = s;
    rho1 = 1.0;
  }
}

// ------------------------------------------------------------------ tile plumbing

template <int TILE, int NV>
struct TileSmem {
  double* sv;            // [NV][TILE] per-observation values being reduced
  double* sw;            // [4][TILE] per-point scratch
  double* sred;          // [8*32] block-reduction scratch
  int* pstart;           // [TILE+1] local observation offset of each point of the tile
  int* coff;             // [TILE+1] image-segment offsets (image order)
  int* cimg;             // [TILE]
  unsigned short* perm;  // [TILE]
  static constexpr size_t bytes() {
    return sizeof(double) * (size_t)(NV * TILE + 4 * TILE + 256) + sizeof(int) * (size_t)(3 * TILE + 2) +
           sizeof(unsigned short) * (size_t)TILE + 16;
  }
  __device__ __forceinline__ void carve(unsigned char* base) {
    sv = reinterpret_cast<double*>(base);
    sw = sv + NV * TILE;
    sred = sw + 4 * TILE;
    pstart = reinterpret_cast<int*>(sred + 256);
    coff = pstart + TILE + 1;
    cimg = coff + TILE + 1;
    perm = reinterpret_cast<unsigned short*>(cimg + TILE);
  }
};

struct TileInfo {
  int base, n, pt0, np, ns;
};

// Tile header: five scalar loads.  The caller then issues ALL of its global loads
// (observation data, Jacobian rows, per-point blocks) before tile_fill_smem(), so that a
// CTA pays one global-memory latency for the whole batch instead of one per stage.
__device__ __forceinline__ TileInfo tile_header(const TileCtx& tc, int& cs0) {
  TileInfo ti;
  const int tile = blockIdx.x;
  ti.base = __ldg(tc.tile_start + tile);
  ti.n = __ldg(tc.tile_start + tile + 1) - ti.base;
  ti.pt0 = __ldg(tc.tile_pt + tile);
  ti.np = __ldg(tc.tile_pt + tile + 1) - ti.pt0;
  cs0 = __ldg(tc.cseg_ptr + tile);
  ti.ns = __ldg(tc.cseg_ptr + tile + 1) - cs0;
  return ti;
}

template <int TILE, int NV>
__device__ __forceinline__ void tile_fill_smem(const TileCtx& tc, TileSmem<TILE, NV>& sm, const TileInfo& ti,
                                               int cs0, bool need_cam) {
  const int tid = threadIdx.x;
  for (int j = tid; j <= ti.np; j += TILE) sm.pstart[j] = __ldg(tc.pt_ptr + ti.pt0 + j) - ti.base;
  if (need_cam) {
    for (int j = tid; j < ti.ns; j += TILE) {
      sm.coff[j] = __ldg(tc.cseg_off + cs0 + j);
      sm.cimg[j] = __ldg(tc.cseg_img + cs0 + j);
    }
    if (tid == 0) sm.coff[ti.ns] = ti.n;
    if (tid < ti.n) sm.perm[tid] = __ldg(tc.tile_perm + ti.base + tid);
  }
  __syncthreads();
}

template <int TILE, int NV>
__device__ __forceinline__ TileInfo tile_prologue(const TileCtx& tc, TileSmem<TILE, NV>& sm, bool need_cam) {
  int cs0;
  const TileInfo ti = tile_header(tc, cs0);
  tile_fill_smem<TILE, NV>(tc, sm, ti, cs0, need_cam);
  return ti;
}

// N block sums with one barrier pair; results valid in threads 0..N-1 (thread j holds
// value j) — all threads must call.
template <int N>
__device__ __forceinline__ double block_sum_multi(const double (&v)[N], double* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double s = warp_sum(v[j]);
    if (lane == 0) sbuf[j * 32 + wid] = s;
  }
  __syncthreads();
  double out = 0.0;
  if (threadIdx.x < N) {
    for (int w = 0; w < nw; ++w) out += sbuf[threadIdx.x * 32 + w];
  }
  return out;
}

// ------------------------------------------------------------------ projection

struct Proj {
  double R[9];
  double w[3];   // R X
  double p[3];   // R X + t
  double u, v, iz;
  double f;
};

__device__ __forceinline__ void load_pose(const double* __restrict__ pose, int img, double (&q)[4], double (&t)[3]) {
  const double2* pp = reinterpret_cast<const double2*>(pose + 8 * (size_t)img);
  const double2 a = __ldg(pp), b = __ldg(pp + 1), c = __ldg(pp + 2), d = __ldg(pp + 3);
  q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
  t[0] = c.x; t[1] = c.y; t[2] = d.x;
}

__device__ __forceinline__ void project(const double (&q)[4], const double (&t)[3], const double* X, Proj& pr) {
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  pr.R[0] = 1.0 - 2.0 * (q2 * q2 + q3 * q3);
  pr.R[1] = 2.0 * (q1 * q2 - q0 * q3);
  pr.R[2] = 2.0 * (q1 * q3 + q0 * q2);
  pr.R[3] = 2.0 * (q1 * q2 + q0 * q3);
  pr.R[4] = 1.0 - 2.0 * (q1 * q1 + q3 * q3);
  pr.R[5] = 2.0 * (q2 * q3 - q0 * q1);
  pr.R[6] = 2.0 * (q1 * q3 - q0 * q2);
  pr.R[7] = 2.0 * (q2 * q3 + q0 * q1);
  pr.R[8] = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    pr.w[a] = pr.R[3 * a] * X[0] + pr.R[3 * a + 1] * X[1] + pr.R[3 * a + 2] * X[2];
    pr.p[a] = pr.w[a] + t[a];
  }
  pr.iz = 1.0 / pr.p[2];
  pr.u = pr.p[0] * pr.iz;
  pr.v = pr.p[1] * pr.iz;
}

// ------------------------------------------------------------------ K1: Jacobian sweep

struct LinArgs {
  const double* pose;     // [F*8] q(4) t(3) pad
  const double* X;        // [P*3]
  const double* K;        // [C*3]
  const double* scale_c;  // [6F+3C] jacobi scaling, 0 on inactive slots
  const double* scale_p;  // [3P]
  LossP loss;
  int intr;               // 0: intrinsics constant, 1: focal only, 3: focal + principal point
  Jac J;
  double* hpp;            // [6][P]  E'E (upper: 00 01 02 11 12 22)
  double* gp;             // [3][P]  E'r
  double* wk;             // [9][P]  (G'E) rows: focal (3) | cx (3) | cy (3)
  double* acc_cam;        // [NREP][F][NVL] rot F'F (6) | t F'F (6) | g (6)
  size_t rep_stride;      // doubles between replicas
  double* acc_intr;       // [C][NVI] G'G (6: ff fx fy xx xy yy) | g_k (3)
  double* acc_cost;       // [1]
};

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_linearize(const TileCtx tc, const LinArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE, 18> sm;
  sm.carve(smem_raw);
  int cs0;
  const TileInfo ti = tile_header(tc, cs0);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const int M = tc.M;
  const size_t i = (size_t)ti.base + tid;
  int img = 0, pt = 0;
  double2 xy = make_double2(0.0, 0.0);
  if (act) {
    img = __ldg(tc.obs_img + i);
    pt = __ldg(tc.obs_pt + i);
    xy = tc.obs_xy[i];
  }
  tile_fill_smem<TILE, 18>(tc, sm, ti, cs0, true);

  double r0 = 0, r1 = 0, cost = 0;
  double jr[2][3], jt[2][3], jp[2][3], jk[4];
#pragma unroll
  for (int a_ = 0; a_ < 2; ++a_)
#pragma unroll
    for (int b_ = 0; b_ < 3; ++b_) { jr[a_][b_] = 0; jt[a_][b_] = 0; jp[a_][b_] = 0; }
  jk[0] = jk[1] = jk[2] = jk[3] = 0;
  int cam = 0;
  if (act) {
    cam = tc.img_cam[img];
    double q[4], t[3];
    load_pose(a.pose, img, q, t);
    const double X[3] = {a.X[3 * (size_t)pt], a.X[3 * (size_t)pt + 1], a.X[3 * (size_t)pt + 2]};
    Proj pr;
    project(q, t, X, pr);
    const double f = a.K[3 * cam], cx = a.K[3 * cam + 1], cy = a.K[3 * cam + 2];
    const double e0 = f * pr.u + cx - xy.x, e1 = f * pr.v + cy - xy.y;
    double rho0, rho1;
    loss_eval(a.loss, e0 * e0 + e1 * e1, rho0, rho1);
    const double sq = sqrt(rho1);
    cost = 0.5 * rho0;
    r0 = sq * e0;
    r1 = sq * e1;
    const double a00 = sq * f * pr.iz, a02 = -a00 * pr.u, a12 = -a00 * pr.v;
    const double* sp = a.scale_p + 3 * (size_t)pt;
    const double* sc = a.scale_c + 6 * (size_t)img;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double s = sp[k];
      jp[0][k] = (a00 * pr.R[k] + a02 * pr.R[6 + k]) * s;
      jp[1][k] = (a00 * pr.R[3 + k] + a12 * pr.R[6 + k]) * s;
    }
    jt[0][0] = a00 * sc[3]; jt[0][2] = a02 * sc[5];
    jt[1][1] = a00 * sc[4]; jt[1][2] = a12 * sc[5];
    if (ROT) {
      // d r / d delta = (d r / d p) * (-2 [R X]x)  (QuaternionParameterization, Plus = exp(d) * q)
      jr[0][0] = 2.0 * a02 * pr.w[1] * sc[0];
      jr[0][1] = 2.0 * (a00 * pr.w[2] - a02 * pr.w[0]) * sc[1];
      jr[0][2] = -2.0 * a00 * pr.w[1] * sc[2];
      jr[1][0] = 2.0 * (a12 * pr.w[1] - a00 * pr.w[2]) * sc[0];
      jr[1][1] = -2.0 * a12 * pr.w[0] * sc[1];
      jr[1][2] = 2.0 * a00 * pr.w[0] * sc[2];
    }
    if (a.intr >= 1) {
      const double* sk = a.scale_c + 6 * (size_t)tc.F + 3 * cam;
      jk[0] = sq * pr.u * sk[0];
      jk[1] = sq * pr.v * sk[0];
      if (a.intr == 3) { jk[2] = sq * sk[1]; jk[3] = sq * sk[2]; }
    }
    // ---- store the linearisation (SoA, coalesced) ----
    a.J.r[i] = r0;
    a.J.r[(size_t)M + i] = r1;
    if (ROT) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a.J.jc[(size_t)k * M + i] = jr[0][k];
        a.J.jc[(size_t)(3 + k) * M + i] = jt[0][k];
        a.J.jc[(size_t)(6 + k) * M + i] = jr[1][k];
        a.J.jc[(size_t)(9 + k) * M + i] = jt[1][k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a.J.jc[(size_t)k * M + i] = jt[0][k];
        a.J.jc[(size_t)(3 + k) * M + i] = jt[1][k];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a.J.jp[(size_t)k * M + i] = jp[0][k];
      a.J.jp[(size_t)(3 + k) * M + i] = jp[1][k];
    }
    if (a.intr >= 1) {
      a.J.jk[i] = jk[0];
      a.J.jk[(size_t)M + i] = jk[1];
      if (a.intr == 3) { a.J.jk[2 * (size_t)M + i] = jk[2]; a.J.jk[3 * (size_t)M + i] = jk[3]; }
    }
  }

  // ---- point side: E'E (6), E'r (3), G'E (3 per free intrinsic) ----
  const int nvp = 9 + 3 * a.intr;
  {
    double* sv = sm.sv + tid;
    sv[0 * TILE] = jp[0][0] * jp[0][0] + jp[1][0] * jp[1][0];
    sv[1 * TILE] = jp[0][0] * jp[0][1] + jp[1][0] * jp[1][1];
    sv[2 * TILE] = jp[0][0] * jp[0][2] + jp[1][0] * jp[1][2];
    sv[3 * TILE] = jp[0][1] * jp[0][1] + jp[1][1] * jp[1][1];
    sv[4 * TILE] = jp[0][1] * jp[0][2] + jp[1][1] * jp[1][2];
    sv[5 * TILE] = jp[0][2] * jp[0][2] + jp[1][2] * jp[1][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) sv[(6 + k) * TILE] = jp[0][k] * r0 + jp[1][k] * r1;
    if (a.intr >= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) sv[(9 + k) * TILE] = jk[0] * jp[0][k] + jk[1] * jp[1][k];
      if (a.intr == 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          sv[(12 + k) * TILE] = jk[2] * jp[0][k];
          sv[(15 + k) * TILE] = jk[3] * jp[1][k];
        }
      }
    }
  }
  __syncthreads();
  for (int pair = tid; pair < nvp * ti.np; pair += TILE) {
    const int k = pair / ti.np, lp = pair - k * ti.np;
    double acc = 0.0;
    const double* row = sm.sv + k * TILE;
    for (int e = sm.pstart[lp]; e < sm.pstart[lp + 1]; ++e) acc += row[e];
    const size_t gpt = (size_t)ti.pt0 + lp;
    if (k < 6) a.hpp[(size_t)k * tc.P + gpt] = acc;
    else if (k < 9) a.gp[(size_t)(k - 6) * tc.P + gpt] = acc;
    else a.wk[(size_t)(k - 9) * tc.P + gpt] = acc;
  }
  __syncthreads();

  // ---- image side: diagonal 3x3 blocks of F'F (rot | t) and F'r ----
  {
    double* sv = sm.sv + tid;
    if (ROT) {
      sv[0 * TILE] = jr[0][0] * jr[0][0] + jr[1][0] * jr[1][0];
      sv[1 * TILE] = jr[0][0] * jr[0][1] + jr[1][0] * jr[1][1];
      sv[2 * TILE] = jr[0][0] * jr[0][2] + jr[1][0] * jr[1][2];
      sv[3 * TILE] = jr[0][1] * jr[0][1] + jr[1][1] * jr[1][1];
      sv[4 * TILE] = jr[0][1] * jr[0][2] + jr[1][1] * jr[1][2];
      sv[5 * TILE] = jr[0][2] * jr[0][2] + jr[1][2] * jr[1][2];
#pragma unroll
      for (int k = 0; k < 3; ++k) sv[(12 + k) * TILE] = jr[0][k] * r0 + jr[1][k] * r1;
    }
    sv[6 * TILE] = jt[0][0] * jt[0][0];
    sv[7 * TILE] = 0.0;
    sv[8 * TILE] = jt[0][0] * jt[0][2];
    sv[9 * TILE] = jt[1][1] * jt[1][1];
    sv[10 * TILE] = jt[1][1] * jt[1][2];
    sv[11 * TILE] = jt[0][2] * jt[0][2] + jt[1][2] * jt[1][2];
    sv[15 * TILE] = jt[0][0] * r0;
    sv[16 * TILE] = jt[1][1] * r1;
    sv[17 * TILE] = jt[0][2] * r0 + jt[1][2] * r1;
  }
  __syncthreads();
  {
    // components handled: ROT -> 0..17, else 6..11 and 15..17 (9 values)
    const int nvc = ROT ? 18 : 9;
    for (int pair = tid; pair < nvc * ti.ns; pair += TILE) {
      int k = pair / ti.ns;
      const int s = pair - k * ti.ns;
      if (!ROT) k = (k < 6) ? k + 6 : k + 9;
      if (k == 7) continue;  // structurally zero
      const double* row = sm.sv + k * TILE;
      double acc = 0.0;
      for (int e = sm.coff[s]; e < sm.coff[s + 1]; ++e) acc += row[sm.perm[e]];
      atomicAdd(a.acc_cam + (size_t)(blockIdx.x & (NREP - 1)) * a.rep_stride + (size_t)sm.cimg[s] * NVL + k, acc);
    }
  }
  // ---- cost and intrinsics (block sums) ----
  {
    double v[3];
    v[0] = cost;
    v[1] = jk[0] * jk[0] + jk[1] * jk[1];
    v[2] = jk[0] * r0 + jk[1] * r1;
    const double s = block_sum_multi<3>(v, sm.sred);
    // single shared camera when intrinsics are free (checked on the host); cost always
    if (tid == 0) atomicAdd(a.acc_cost, s);
    if (a.intr >= 1) {
      if (tid == 1) atomicAdd(a.acc_intr + 0, s);
      if (tid == 2) atomicAdd(a.acc_intr + 6, s);
    }
    if (a.intr == 3) {
      double u[7];
      u[0] = jk[0] * jk[2];            // f-cx
      u[1] = jk[1] * jk[3];            // f-cy
      u[2] = jk[2] * jk[2];            // cx-cx
      u[3] = 0.0;                      // cx-cy
      u[4] = jk[3] * jk[3];            // cy-cy
      u[5] = jk[2] * r0;               // g cx
      u[6] = jk[3] * r1;               // g cy
      const double s2 = block_sum_multi<7>(u, sm.sred);
      if (tid < 5) atomicAdd(a.acc_intr + 1 + tid, s2);
      else if (tid < 7) atomicAdd(a.acc_intr + 7 + (tid - 5), s2);
    }
  }
}

// ------------------------------------------------------------------ K2: point blocks

struct PtsArgs {
  const double* hpp;   // [6][P]
  const double* gp;    // [3][P]
  const double* wk;    // [9][P]
  const double* scale_p;  // [3P]
  double radius, min_diag, max_diag;
  int intr;
  int P;
  double* hinv;        // [6][P] (E'E + D^2)^-1 upper
  double* w;           // [3][P] hinv * E'r
  double* acc_intr;    // [NVI] -(G'E) hinv (E'G) (6) | -(G'E) hinv E'r (3)
  double* acc_fail;    // [1] > 0 when a block is not positive definite
  double* gmax;        // [1] max |unscaled gradient| over point parameters (atomic max)
};

__global__ void __launch_bounds__(256) k_point_blocks(const PtsArgs a) {
  __shared__ double sred[9 * 32];
  const int p = blockIdx.x * 256 + threadIdx.x;
  double gm = 0.0, fail = 0.0;
  double ci[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) ci[k] = 0.0;
  if (p < a.P) {
    const size_t P = a.P;
    double h00 = a.hpp[p], h01 = a.hpp[P + p], h02 = a.hpp[2 * P + p], h11 = a.hpp[3 * P + p],
           h12 = a.hpp[4 * P + p], h22 = a.hpp[5 * P + p];
    const double g0 = a.gp[p], g1 = a.gp[P + p], g2 = a.gp[2 * P + p];
    // LevenbergMarquardtStrategy: D^2 = clamp(diag(J'J)) / radius
    h00 += fmin(fmax(h00, a.min_diag), a.max_diag) / a.radius;
    h11 += fmin(fmax(h11, a.min_diag), a.max_diag) / a.radius;
    h22 += fmin(fmax(h22, a.min_diag), a.max_diag) / a.radius;
    // Cholesky-based inverse
    bool bad = !(h00 > 0.0);
    const double l00 = sqrt(h00);
    const double l10 = h01 / l00, l20 = h02 / l00;
    double d = h11 - l10 * l10;
    bad |= !(d > 0.0);
    const double l11 = sqrt(d);
    const double l21 = (h12 - l20 * l10) / l11;
    d = h22 - l20 * l20 - l21 * l21;
    bad |= !(d > 0.0);
    const double l22 = sqrt(d);
    const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
    const double i10 = -l10 * i00 * i11;
    const double i21 = -l21 * i11 * i22;
    const double i20 = -(l20 * i00 + l21 * i10) * i22;
    double v00 = i00 * i00 + i10 * i10 + i20 * i20, v01 = i10 * i11 + i20 * i21, v02 = i20 * i22,
           v11 = i11 * i11 + i21 * i21, v12 = i21 * i22, v22 = i22 * i22;
    if (bad) { fail = 1.0; v00 = v11 = v22 = 1.0; v01 = v02 = v12 = 0.0; }
    a.hinv[p] = v00; a.hinv[P + p] = v01; a.hinv[2 * P + p] = v02;
    a.hinv[3 * P + p] = v11; a.hinv[4 * P + p] = v12; a.hinv[5 * P + p] = v22;
    const double w0 = v00 * g0 + v01 * g1 + v02 * g2;
    const double w1 = v01 * g0 + v11 * g1 + v12 * g2;
    const double w2 = v02 * g0 + v12 * g1 + v22 * g2;
    a.w[p] = w0; a.w[P + p] = w1; a.w[2 * P + p] = w2;
    const double* sp = a.scale_p + 3 * (size_t)p;
    gm = fmax(fmax(fabs(g0 / sp[0]), fabs(g1 / sp[1])), fabs(g2 / sp[2]));
    if (a.intr >= 1) {
      double W[3][3], WH[3][3];
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) W[j][k] = (j < a.intr) ? a.wk[(size_t)(3 * j + k) * P + p] : 0.0;
      for (int j = 0; j < 3; ++j) {
        WH[j][0] = W[j][0] * v00 + W[j][1] * v01 + W[j][2] * v02;
        WH[j][1] = W[j][0] * v01 + W[j][1] * v11 + W[j][2] * v12;
        WH[j][2] = W[j][0] * v02 + W[j][1] * v12 + W[j][2] * v22;
      }
      int c = 0;
      for (int j = 0; j < 3; ++j)
        for (int k = j; k < 3; ++k) ci[c++] = -(WH[j][0] * W[k][0] + WH[j][1] * W[k][1] + WH[j][2] * W[k][2]);
      for (int j = 0; j < 3; ++j) ci[6 + j] = -(W[j][0] * w0 + W[j][1] * w1 + W[j][2] * w2);
    }
  }
  const double m = block_max(gm, sred);
  if (threadIdx.x == 0 && m > 0.0) atomic_max_nonneg(a.gmax, m);
  const double f = block_sum(fail, sred);
  if (threadIdx.x == 0 && f > 0.0) atomicAdd(a.acc_fail, f);
  if (a.intr >= 1) {
    const double s = block_sum_multi<9>(ci, sred);
    if (threadIdx.x < 9) atomicAdd(a.acc_intr + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ K3: Schur-Jacobi / rhs corrections

struct PrepArgs {
  Jac J;
  const double* hinv;   // [6][P]
  const double* w;      // [3][P]
  double* acc_cam;      // [NREP][F][NVL]: -(W hinv W') rot (6) | t (6) | -(W w) (6)
  size_t rep_stride;
};

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_schur_prep(const TileCtx tc, const PrepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE, 18> sm;
  sm.carve(smem_raw);
  int cs0;
  const TileInfo ti = tile_header(tc, cs0);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const size_t M = tc.M, P = tc.P;
  const size_t i = (size_t)ti.base + tid;
  constexpr int NCP = ROT ? 6 : 3;
  double jp[2][3], hv[6], w[3], jcr0[NCP], jcr1[NCP];
  if (act) {
    const int pt = __ldg(tc.obs_pt + i);
#pragma unroll
    for (int k = 0; k < 3; ++k) { jp[0][k] = a.J.jp[k * M + i]; jp[1][k] = a.J.jp[(3 + k) * M + i]; }
#pragma unroll
    for (int k = 0; k < NCP; ++k) { jcr0[k] = a.J.jc[(size_t)k * M + i]; jcr1[k] = a.J.jc[(size_t)(NCP + k) * M + i]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) hv[k] = a.hinv[k * P + pt];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = a.w[k * P + pt];
  }
  tile_fill_smem<TILE, 18>(tc, sm, ti, cs0, true);
  double* sv = sm.sv + tid;
  if (act) {
    // block b: 0 = rot (ROT only), 1 = translation
#pragma unroll
    for (int b = (ROT ? 0 : 1); b < 2; ++b) {
      double jc0[3], jc1[3];
      const int off = ROT ? 3 * b : 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        jc0[k] = jcr0[off + k];
        jc1[k] = jcr1[off + k];
      }
      double W[3][3], WH[3][3];
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) W[j][k] = jc0[j] * jp[0][k] + jc1[j] * jp[1][k];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        WH[j][0] = W[j][0] * hv[0] + W[j][1] * hv[1] + W[j][2] * hv[2];
        WH[j][1] = W[j][0] * hv[1] + W[j][1] * hv[3] + W[j][2] * hv[4];
        WH[j][2] = W[j][0] * hv[2] + W[j][1] * hv[4] + W[j][2] * hv[5];
      }
      int c = 6 * b;
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = j; k < 3; ++k) { sv[c * TILE] = -(WH[j][0] * W[k][0] + WH[j][1] * W[k][1] + WH[j][2] * W[k][2]); ++c; }
#pragma unroll
      for (int j = 0; j < 3; ++j) sv[(12 + 3 * b + j) * TILE] = -(W[j][0] * w[0] + W[j][1] * w[1] + W[j][2] * w[2]);
    }
  }
  __syncthreads();
  const int nvc = ROT ? 18 : 9;
  for (int pair = tid; pair < nvc * ti.ns; pair += TILE) {
    int k = pair / ti.ns;
    const int s = pair - k * ti.ns;
    if (!ROT) k = (k < 6) ? k + 6 : k + 9;
    const double* row = sm.sv + k * TILE;
    double acc = 0.0;
    for (int e = sm.coff[s]; e < sm.coff[s + 1]; ++e) acc += row[sm.perm[e]];
    atomicAdd(a.acc_cam + (size_t)(blockIdx.x & (NREP - 1)) * a.rep_stride + (size_t)sm.cimg[s] * NVL + k, acc);
  }
}

// ------------------------------------------------------------------ K4: implicit S * p

struct SpArgs {
  Jac J;
  const double* hinv;   // [6][P]
  const double* x;      // [6F + 3C] input vector (slot layout)
  double* y;            // [NREP][6F + 3C] += F'(I - E hinv E') F x  (D^2 x is added by the PCG kernel)
  size_t rep_stride;
  const int* flag;      // PCG state: != 0 -> nothing to do
  int intr;
};

#ifndef PSFM_SP_MINB
#define PSFM_SP_MINB 3
#endif
template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? PSFM_SP_MINB : 1)) k_schur_product(const TileCtx tc, const SpArgs a) {
  if (*a.flag != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE, 6> sm;
  sm.carve(smem_raw);
  int cs0;
  const TileInfo ti = tile_header(tc, cs0);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const size_t M = tc.M, P = tc.P;
  const size_t i = (size_t)ti.base + tid;
  constexpr int NC = ROT ? 6 : 3;
  double jc0[NC], jc1[NC], jp[2][3], jk[4] = {0, 0, 0, 0};
  double hp[6] = {0, 0, 0, 0, 0, 0};
  double u0 = 0, u1 = 0;
  int img = 0, lp = 0;
  const double* xk = a.x + 6 * (size_t)tc.F;   // single shared camera when intr > 0
#pragma unroll
  for (int k = 0; k < NC; ++k) { jc0[k] = 0; jc1[k] = 0; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { jp[0][k] = 0; jp[1][k] = 0; }
  // ---- every global load of the tile is issued here, before the first barrier ----
  if (act) {
    img = __ldg(tc.obs_img + i);
    lp = __ldg(tc.obs_pt + i) - ti.pt0;
#pragma unroll
    for (int k = 0; k < NC; ++k) { jc0[k] = a.J.jc[(size_t)k * M + i]; jc1[k] = a.J.jc[(size_t)(NC + k) * M + i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { jp[0][k] = a.J.jp[k * M + i]; jp[1][k] = a.J.jp[(3 + k) * M + i]; }
    if (a.intr >= 1) {
      jk[0] = a.J.jk[i]; jk[1] = a.J.jk[M + i];
      if (a.intr == 3) { jk[2] = a.J.jk[2 * M + i]; jk[3] = a.J.jk[3 * M + i]; }
    }
  }
  if (tid < ti.np) {
#pragma unroll
    for (int k = 0; k < 6; ++k) hp[k] = a.hinv[k * P + (size_t)ti.pt0 + tid];
  }
  tile_fill_smem<TILE, 6>(tc, sm, ti, cs0, true);
  if (act) {
    const double* xi = a.x + 6 * (size_t)img + (ROT ? 0 : 3);
#pragma unroll
    for (int k = 0; k < NC; ++k) { const double xv = __ldg(xi + k); u0 += jc0[k] * xv; u1 += jc1[k] * xv; }
    if (a.intr >= 1) {
      const double xf = __ldg(xk);
      u0 += jk[0] * xf; u1 += jk[1] * xf;
      if (a.intr == 3) {
        u0 += jk[2] * __ldg(xk + 1);
        u1 += jk[3] * __ldg(xk + 2);
      }
    }
  }
  double* sv = sm.sv + tid;
#pragma unroll
  for (int k = 0; k < 3; ++k) sv[k * TILE] = jp[0][k] * u0 + jp[1][k] * u1;
  __syncthreads();
  // per point: t = E'u ; w = hinv t
  for (int pair = tid; pair < 3 * ti.np; pair += TILE) {
    const int k = pair / ti.np, l = pair - k * ti.np;
    const double* row = sm.sv + k * TILE;
    double acc = 0.0;
    for (int e = sm.pstart[l]; e < sm.pstart[l + 1]; ++e) acc += row[e];
    sm.sw[k * TILE + l] = acc;
  }
  __syncthreads();
  if (tid < ti.np) {
    const double t0 = sm.sw[tid], t1 = sm.sw[TILE + tid], t2 = sm.sw[2 * TILE + tid];
    sm.sw[tid] = hp[0] * t0 + hp[1] * t1 + hp[2] * t2;
    sm.sw[TILE + tid] = hp[1] * t0 + hp[3] * t1 + hp[4] * t2;
    sm.sw[2 * TILE + tid] = hp[2] * t0 + hp[4] * t1 + hp[5] * t2;
  }
  __syncthreads();
  const double w0 = sm.sw[lp], w1 = sm.sw[TILE + lp], w2 = sm.sw[2 * TILE + lp];
  const double v0 = u0 - (jp[0][0] * w0 + jp[0][1] * w1 + jp[0][2] * w2);
  const double v1 = u1 - (jp[1][0] * w0 + jp[1][1] * w1 + jp[1][2] * w2);
  __syncthreads();   // sv is rewritten below; sw reads are done
#pragma unroll
  for (int k = 0; k < NC; ++k) sv[k * TILE] = act ? (jc0[k] * v0 + jc1[k] * v1) : 0.0;
  __syncthreads();
  for (int pair = tid; pair < NC * ti.ns; pair += TILE) {
    const int k = pair / ti.ns, s = pair - k * ti.ns;
    const double* row = sm.sv + k * TILE;
    double acc = 0.0;
    for (int e = sm.coff[s]; e < sm.coff[s + 1]; ++e) acc += row[sm.perm[e]];
    atomicAdd(a.y + (size_t)(blockIdx.x & (NREP - 1)) * a.rep_stride + 6 * (size_t)sm.cimg[s] + (ROT ? 0 : 3) + k, acc);
  }
  if (a.intr >= 1) {
    double v[3];
    v[0] = act ? jk[0] * v0 + jk[1] * v1 : 0.0;
    v[1] = act ? jk[2] * v0 : 0.0;
    v[2] = act ? jk[3] * v1 : 0.0;
    const double s = block_sum_multi<3>(v, sm.sred);
    if (tid < a.intr) atomicAdd(a.y + (size_t)(blockIdx.x & (NREP - 1)) * a.rep_stride + 6 * (size_t)tc.F + tid, s);
  }
}

// ------------------------------------------------------------------ K5: back-substitution + model cost + candidate points

struct BackArgs {
  Jac J;
  const double* hinv;   // [6][P]
  const double* w;      // [3][P] hinv * E'r
  const double* yc;     // [6F + 3C] reduced-system solution (step_c = -yc)
  const double* scale_p;
  const double* X;      // [3P] current points
  double* Xc;           // [3P] candidate points
  double* acc;          // [0] sum m(r + m/2)   [1] |dX|^2   [2] |Xc|^2
  int intr;
};

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_back_substitute(const TileCtx tc, const BackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE, 6> sm;
  sm.carve(smem_raw);
  int cs0;
  const TileInfo ti = tile_header(tc, cs0);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const size_t M = tc.M, P = tc.P;
  const size_t i = (size_t)ti.base + tid;
  constexpr int NC = ROT ? 6 : 3;
  double jc0[NC], jc1[NC], jp[2][3], jk[4] = {0, 0, 0, 0};
  double u0 = 0, u1 = 0, r0 = 0, r1 = 0;
  int img = 0, lp = 0;
#pragma unroll
  for (int k = 0; k < NC; ++k) { jc0[k] = 0; jc1[k] = 0; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { jp[0][k] = 0; jp[1][k] = 0; }
  // ---- all global loads up front ----
  if (act) {
    img = __ldg(tc.obs_img + i);
    lp = __ldg(tc.obs_pt + i) - ti.pt0;
#pragma unroll
    for (int k = 0; k < NC; ++k) { jc0[k] = a.J.jc[(size_t)k * M + i]; jc1[k] = a.J.jc[(size_t)(NC + k) * M + i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { jp[0][k] = a.J.jp[k * M + i]; jp[1][k] = a.J.jp[(3 + k) * M + i]; }
    if (a.intr >= 1) {
      jk[0] = a.J.jk[i]; jk[1] = a.J.jk[M + i];
      if (a.intr == 3) { jk[2] = a.J.jk[2 * M + i]; jk[3] = a.J.jk[3 * M + i]; }
    }
    r0 = a.J.r[i]; r1 = a.J.r[M + i];
  }
  double hp[6] = {0, 0, 0, 0, 0, 0}, wp[3] = {0, 0, 0}, spp[3] = {0, 0, 0}, Xp[3] = {0, 0, 0};
  if (tid < ti.np) {
    const size_t g = (size_t)ti.pt0 + tid;
#pragma unroll
    for (int k = 0; k < 6; ++k) hp[k] = a.hinv[k * P + g];
#pragma unroll
    for (int k = 0; k < 3; ++k) { wp[k] = a.w[k * P + g]; spp[k] = a.scale_p[3 * g + k]; Xp[k] = a.X[3 * g + k]; }
  }
  tile_fill_smem<TILE, 6>(tc, sm, ti, cs0, false);
  if (act) {
    const double* xi = a.yc + 6 * (size_t)img + (ROT ? 0 : 3);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double xv = __ldg(xi + k);
      u0 += jc0[k] * xv;
      u1 += jc1[k] * xv;
    }
    if (a.intr >= 1) {
      const double* xk = a.yc + 6 * (size_t)tc.F;
      const double xf = __ldg(xk);
      u0 += jk[0] * xf; u1 += jk[1] * xf;
      if (a.intr == 3) { u0 += jk[2] * __ldg(xk + 1); u1 += jk[3] * __ldg(xk + 2); }
    }
  }
  double* sv = sm.sv + tid;
#pragma unroll
  for (int k = 0; k < 3; ++k) sv[k * TILE] = jp[0][k] * u0 + jp[1][k] * u1;
  __syncthreads();
  for (int pair = tid; pair < 3 * ti.np; pair += TILE) {
    const int k = pair / ti.np, l = pair - k * ti.np;
    const double* row = sm.sv + k * TILE;
    double acc = 0.0;
    for (int e = sm.pstart[l]; e < sm.pstart[l + 1]; ++e) acc += row[e];
    sm.sw[k * TILE + l] = acc;
  }
  __syncthreads();
  double dx2 = 0.0, xc2 = 0.0;
  if (tid < ti.np) {
    const size_t g = (size_t)ti.pt0 + tid;
    const double t0 = sm.sw[tid], t1 = sm.sw[TILE + tid], t2 = sm.sw[2 * TILE + tid];
    // y_p = hinv (E'r - E'F y_c) = w - hinv t ;  step_p = -y_p
    const double y0 = wp[0] - (hp[0] * t0 + hp[1] * t1 + hp[2] * t2);
    const double y1 = wp[1] - (hp[1] * t0 + hp[3] * t1 + hp[4] * t2);
    const double y2 = wp[2] - (hp[2] * t0 + hp[4] * t1 + hp[5] * t2);
    sm.sw[tid] = y0; sm.sw[TILE + tid] = y1; sm.sw[2 * TILE + tid] = y2;
    const double d0 = -y0 * spp[0], d1 = -y1 * spp[1], d2 = -y2 * spp[2];
    const double c0 = Xp[0] + d0, c1 = Xp[1] + d1, c2 = Xp[2] + d2;
    a.Xc[3 * g] = c0; a.Xc[3 * g + 1] = c1; a.Xc[3 * g + 2] = c2;
    const double e0 = Xp[0] - c0, e1 = Xp[1] - c1, e2 = Xp[2] - c2;
    dx2 = e0 * e0 + e1 * e1 + e2 * e2;
    xc2 = c0 * c0 + c1 * c1 + c2 * c2;
  }
  __syncthreads();
  // model residual of this observation: m = J * step = -(u + E y_p)
  double mm = 0.0;
  if (act) {
    const double y0 = sm.sw[lp], y1 = sm.sw[TILE + lp], y2 = sm.sw[2 * TILE + lp];
    const double m0 = -(u0 + jp[0][0] * y0 + jp[0][1] * y1 + jp[0][2] * y2);
    const double m1 = -(u1 + jp[1][0] * y0 + jp[1][1] * y1 + jp[1][2] * y2);
    mm = m0 * (r0 + 0.5 * m0) + m1 * (r1 + 0.5 * m1);
  }
  double v[3] = {mm, dx2, xc2};
  const double s = block_sum_multi<3>(v, sm.sred);
  if (tid < 3) atomicAdd(a.acc + tid, s);
}

// ------------------------------------------------------------------ K6: cost only

struct CostArgs {
  const double* pose;
  const double* X;
  const double* K;
  LossP loss;
  double* acc_cost;
};

__global__ void __launch_bounds__(256) k_cost(const TileCtx tc, const CostArgs a) {
  __shared__ double sred[32];
  double cost = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)tc.M; i += (size_t)gridDim.x * 256) {
    const int img = tc.obs_img[i], pt = tc.obs_pt[i], cam = tc.img_cam[img];
    const double2 xy = tc.obs_xy[i];
    double q[4], t[3];
    load_pose(a.pose, img, q, t);
    const double X[3] = {a.X[3 * (size_t)pt], a.X[3 * (size_t)pt + 1], a.X[3 * (size_t)pt + 2]};
    Proj pr;
    project(q, t, X, pr);
    const double f = a.K[3 * cam], cx = a.K[3 * cam + 1], cy = a.K[3 * cam + 2];
    const double e0 = f * pr.u + cx - xy.x, e1 = f * pr.v + cy - xy.y;
    double rho0, rho1;
    loss_eval(a.loss, e0 * e0 + e1 * e1, rho0, rho1);
    cost += 0.5 * rho0;
  }
  const double s = block_sum(cost, sred);
  if (threadIdx.x == 0) atomicAdd(a.acc_cost, s);
}

}  // namespace ba
}  // namespace psfm
