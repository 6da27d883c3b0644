// ba_schur_explicit.cuh — EXACT reduced-camera-system solve on the device.
//
// The reference solves the reduced camera system exactly for <= 1000 images (DENSE_SCHUR /
// SPARSE_SCHUR, bundle_adjustment.cc:276-286): Ceres' SchurEliminator forms
//     S = F'F + D_c^2 - sum_p (F'E)_p (E'E + D_p^2)^-1 (E'F)_p
// and factorises it.  This file does the same on the GPU:
//   k_schur_tile[_p] FUSED path: per observation the compact factors [D | w | Q = Jp H~ | Jp]
//                    stay in shared memory; per-image cross blocks / focal column / rhs
//                    correction; the tile's pair tasks sum (W_i H~) W_j' = Jc_i' (Q_i Jp_j') Jc_j
//                    into the band-block accumulator Sband[a][b - a][36]
//   pair structure   all (i, j) observation pairs of a point keyed by (tile, image a, image b)
//                    (built once per problem with radix sort / run-length encode)
//   k_schur_w, k_schur_pairs   unfused fallback through HBM (W, W H~ per observation, AoS),
//                    used when a tile's staging does not fit in shared memory
//   k_schur_assemble dense symmetric S (6F+3C)^2 with Jacobi scaling, LM diagonal, gauge
//   k_chol_blocked   cooperative blocked banded(+arrow) Cholesky and the triangular solves,
//                    band = 6 * (longest image span of a track) — video tracks make S banded
// Everything accumulates with the UNSCALED factored Jacobian (see ba_kernels.cuh); the
// scaling diag(s) is applied in k_schur_assemble.
#pragma once
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "ba_kernels.cuh"
#include "ba_tile_pipe.cuh"

namespace psfm {
namespace ba {

constexpr int NVX = 21;   // k_schur_w per-image sums: rot-t cross block (9) | F'G focal (6) | -(W H~) Wk' (6)

// ------------------------------------------------------------------ per-observation W, W H~

struct SwArgs {
  Lin L;
  const double* pose16;
  const double* X;
  const double* ht;     // [6][P]
  const double* wk;     // [9][P] G'E per point (focal row used)
  const double* K;
  double* W;            // [M][18]  rows: rot 0..2, t 0..2 ; 3 values per row
  double* WH;           // [M][18]
  double* acc_cam;      // [NREP][F][NVX]
  size_t rep_stride;
  int intr;
};

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_schur_w(const TileCtx tc, const SwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE> sm;
  sm.carve(smem_raw, NVX, 12, tc.cap_ns, tc.cap_np);
  const TileInfo ti = tile_header(tc);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const size_t M = tc.M;
  const size_t i = (size_t)ti.base + tid;
  int ls = 0, lp = 0;
  double a00 = 0, a02 = 0, a12 = 0;
  if (act) {
    ls = __ldg(tc.obs_lseg + i);
    lp = __ldg(tc.obs_lpt + i);
    a00 = a.L.a[i]; a02 = a.L.a[M + i]; a12 = a.L.a[2 * M + i];
  }
  const double inv_f = (a.intr >= 1) ? 1.0 / __ldg(a.K) : 0.0;
  // per point: X (0..2), H~ (3..8), focal row of G'E (9..11)
  tile_fill_smem<TILE>(tc, sm, ti, a.pose16, nullptr, a.X, a.ht, a.wk, true);
  double* sv = sm.sv + tid;
#pragma unroll
  for (int k = 0; k < NVX; ++k) sv[k * PSFM_SVS] = 0.0;
  if (act) {
    ObsGeom g;
    load_geom<TILE>(sm, ls, lp, g);
    const int cnp = sm.cap_np;
    double hv[6], wkp[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) hv[k] = sm.prow(3 + k)[lp];
#pragma unroll
    for (int k = 0; k < 3; ++k) wkp[k] = sm.prow(9 + k)[lp];
    double jp[2][3], jc[2][6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      jp[0][k] = a00 * g.R[k] + a02 * g.R[6 + k];
      jp[1][k] = a00 * g.R[3 + k] + a12 * g.R[6 + k];
    }
    if (ROT) {
      jc[0][0] = 2.0 * a02 * g.w[1]; jc[0][1] = 2.0 * (a00 * g.w[2] - a02 * g.w[0]); jc[0][2] = -2.0 * a00 * g.w[1];
      jc[1][0] = 2.0 * (a12 * g.w[1] - a00 * g.w[2]); jc[1][1] = -2.0 * a12 * g.w[0]; jc[1][2] = 2.0 * a00 * g.w[0];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { jc[0][k] = 0.0; jc[1][k] = 0.0; }
    }
    jc[0][3] = a00; jc[0][4] = 0.0; jc[0][5] = a02;
    jc[1][3] = 0.0; jc[1][4] = a00; jc[1][5] = a12;
    double W[6][3], WH[6][3];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) W[r][k] = jc[0][r] * jp[0][k] + jc[1][r] * jp[1][k];
      WH[r][0] = W[r][0] * hv[0] + W[r][1] * hv[1] + W[r][2] * hv[2];
      WH[r][1] = W[r][0] * hv[1] + W[r][1] * hv[3] + W[r][2] * hv[4];
      WH[r][2] = W[r][0] * hv[2] + W[r][1] * hv[4] + W[r][2] * hv[5];
    }
    double2* Wo = reinterpret_cast<double2*>(a.W + 18 * i);
    double2* WHo = reinterpret_cast<double2*>(a.WH + 18 * i);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      Wo[k] = make_double2(W[(2 * k) / 3][(2 * k) % 3], W[(2 * k + 1) / 3][(2 * k + 1) % 3]);
      WHo[k] = make_double2(WH[(2 * k) / 3][(2 * k) % 3], WH[(2 * k + 1) / 3][(2 * k + 1) % 3]);
    }
    // rot-t cross block of F'F: (Jr' Jt)[r][c]
    if (ROT) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) sv[(3 * r + c) * PSFM_SVS] = jc[0][r] * jc[0][3 + c] + jc[1][r] * jc[1][3 + c];
    }
    if (a.intr >= 1) {
      const double zf = (g.w[2] + g.tz) * inv_f;
      const double jf0 = -a02 * zf, jf1 = -a12 * zf;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        sv[(9 + r) * PSFM_SVS] = jc[0][r] * jf0 + jc[1][r] * jf1;                               // F'G
        sv[(15 + r) * PSFM_SVS] = -(WH[r][0] * wkp[0] + WH[r][1] * wkp[1] + WH[r][2] * wkp[2]);  // -(W H~) Wk'
      }
    }
  }
  __syncthreads();
  double* dst = a.acc_cam + (size_t)(blockIdx.x & (NREP - 1)) * a.rep_stride;
  tile_reduce_images<TILE>(sm, ti, NVX, [&](int k, int img, double acc) {
    if (acc != 0.0) atomicAdd(dst + (size_t)img * NVX + k, acc);
  });
}

// ------------------------------------------------------------------ pair structure

// entries started by observation j (sorted order): (j, j), (j, j+1) ... (j, end of its point)
// plus (j, j') for earlier observations j' of the same point IN THE SAME IMAGE (rare
// duplicates: both orders are needed inside a diagonal block)
__global__ void k_pair_count(const int* pt_ptr, const int* obs_pt, const int* obs_img, int M, int* cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int p = obs_pt[j];
  const int b = pt_ptr[p], e = pt_ptr[p + 1];
  int c = e - j;
  for (int k = j - 1; k >= b && obs_img[k] == obs_img[j]; --k) ++c;
  cnt[j] = c;
}

__global__ void k_pair_fill(const int* pt_ptr, const int* obs_pt, const int* obs_img, const int* ptr, int M, int F,
                            unsigned int* keys, unsigned long long* vals) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int p = obs_pt[j];
  const int b = pt_ptr[p], e = pt_ptr[p + 1];
  const int a = obs_img[j];
  size_t o = (size_t)ptr[j];
  for (int k = j; k < e; ++k, ++o) {
    keys[o] = (unsigned)a * (unsigned)F + (unsigned)obs_img[k];
    vals[o] = ((unsigned long long)(unsigned)j << 32) | (unsigned)k;
  }
  for (int k = j - 1; k >= b && obs_img[k] == a; --k, ++o) {
    keys[o] = (unsigned)a * (unsigned)F + (unsigned)a;
    vals[o] = ((unsigned long long)(unsigned)j << 32) | (unsigned)k;
  }
}

// ------------------------------------------------------------------ block products

struct PairArgs {
  const unsigned long long* entries;   // (i << 32) | j, sorted by image pair
  const int* chunk_blk;                // [nchunks] block id
  const long long* chunk_beg;          // [nchunks + 1] entry range of the chunk
  const double* W;                     // [M][18]
  const double* WH;                    // [M][18]
  double* Sblk;                        // [nblocks][36]  += sum (W_i H~) W_j'
};

__global__ void __launch_bounds__(128) k_schur_pairs(const PairArgs a) {
  __shared__ double sred[36 * 4];
  const int ch = blockIdx.x;
  const long long beg = a.chunk_beg[ch], end = a.chunk_beg[ch + 1];
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  for (long long e = beg + threadIdx.x; e < end; e += 128) {
    const unsigned long long ij = a.entries[e];
    const size_t i = (size_t)(ij >> 32), j = (size_t)(ij & 0xffffffffull);
    const double2* wh = reinterpret_cast<const double2*>(a.WH + 18 * i);
    const double2* wj = reinterpret_cast<const double2*>(a.W + 18 * j);
    double A[18], B[18];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const double2 u = __ldg(wh + k), v = __ldg(wj + k);
      A[2 * k] = u.x; A[2 * k + 1] = u.y;
      B[2 * k] = v.x; B[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c)
        acc[6 * r + c] += A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 36; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) sred[k * 4 + wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    const double s = sred[threadIdx.x * 4] + sred[threadIdx.x * 4 + 1] + sred[threadIdx.x * 4 + 2] + sred[threadIdx.x * 4 + 3];
    atomicAdd(a.Sblk + 36 * (size_t)a.chunk_blk[ch] + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ fused tile path: W in shared memory, pairs per tile
//
// A point never straddles a tile, so every (i, j) observation pair of the Schur complement
// lives inside one tile.  k_schur_tile keeps W_i and W_i H~ of its <= TILE observations in
// shared memory (nothing per-observation goes to HBM) and runs the tile's pair TASKS: a task
// is the run of pair entries of one image pair (a, b) inside the tile; two threads share a
// task (block rows 0-2 / 3-5), accumulate sum (W_i H~) W_j' in registers and issue one RED
// per element into the band-block accumulator  Sband[a][b - a][36]  (b - a <= span, the
// longest image span of a track, global over the ranks).  The per-image sums (rot-t cross
// block, focal column, rhs correction -W w^) are reduced per image segment as in the other
// tile kernels.

constexpr int NVX2 = 27;   // NVX (21) | -(W w^) rot (3) | t (3)
constexpr int PS = 18;     // shared-memory stride of one observation's record: 9 x 16 bytes, read with LDS.128 — eight
                           // consecutive records tile the 32 banks, so the records of one point (consecutive) do not collide

struct StArgs {
  Lin L;
  const double* pose16;
  const double* X;
  const double* ht;     // [6][P]
  const double* wt;     // [3][P] w^ = H~ g^
  const double* wk;     // [9][P] G'E per point (focal row used)
  const double* K;
  double* acc_cam;      // [NREP][F][NVX2]
  size_t rep_stride;
  int intr;
  const unsigned int* entries;   // (li << 16) | lj, tile-local indices, sorted by (tile, a, b)
  const int* task_slot;          // [ntasks] band-block slot a * (span + 1) + (b - a)
  const int2* task_rng;          // [ntasks] entry range (begin, end) of the task; tasks of a tile ordered longest first
  const int* tile_task;          // [T + 1] task range of the tile
  double* Sband;                 // [nrep][F * (span + 1) * 36]
  size_t band_stride;
  int nrep_mask;
  int dbg;      // PSFM_SCHUR_FLAGS (measurement only): 1 = no band REDs, 2 = no pair loop
};

// One tile of the fused Schur kernel; shared memory holds the staged inputs (see linearize_tile).
template <int TILE, bool ROT>
__device__ __forceinline__ void schur_tile_body(const TileCtx& tc, const StArgs& a, TileSmem<TILE>& sm, const TileInfo& ti,
                                                const bool act, const int ls, const int lp, const double a00, const double a02,
                                                const double a12, const int rep, const int t0, const int nt) {
  const int tid = threadIdx.x;
  const double inv_f = (a.intr >= 1) ? 1.0 / __ldg(a.K) : 0.0;
  // Pair tasks of this tile -> lanes: two lanes share a task (q = 2 task + part; entries e0 + part, stride 2).
  // Measured and dropped (round 2, profiles/r02_schur_tile_notes.md): runs cut into units of <= 8 entries so that
  // all 8 warps carry pairs (more REDs: 1.11 -> 1.88 ms), three lanes per task (7 trips instead of 11 on the
  // tile's critical path: 0.98 -> 1.05 ms) — the pair loop is bound by its fp64 instruction count, not by the
  // longest lane.
  constexpr int lpt = 2;
  const int nq = (a.dbg & 2) ? 0 : 2 * nt;
  auto lane_task = [&](int q, int& tq, int& par) -> bool {
    par = q & 1; tq = q >> 1;
    return q < nq;
  };
  // the first pass's task range / slot are fetched NOW: two dependent global loads (range, then entries)
  // would otherwise sit on the critical path of every tile right after the last barrier
  int tq_f, par_f;
  const bool valid_f = lane_task(tid, tq_f, par_f);
  int2 rg_f = make_int2(0, 0);
  int slot_f = 0;
  if (valid_f) { rg_f = __ldg(a.task_rng + t0 + tq_f); slot_f = __ldg(a.task_slot + t0 + tq_f); }
  double* sv = sm.sv + tid;
#pragma unroll
  for (int k = 0; k < NVX2; ++k) sv[k * PSFM_SVS] = 0.0;
  // compact per-observation factors kept for the pair phase (W = Jc' Jp and W H~ = Jc' Q are
  // never formed in memory):  rec = [a00 a02 a12 | w (3) | Q = Jp H~ (2x3) | Jp (2x3)]
  double rec[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) rec[k] = 0.0;
  if (act) {
    ObsGeom g;
    load_geom<TILE>(sm, ls, lp, g);
    const int cnp = sm.cap_np;
    double hv[6], wkp[3], wh[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) hv[k] = sm.prow(3 + k)[lp];
#pragma unroll
    for (int k = 0; k < 3; ++k) { wkp[k] = sm.prow(9 + k)[lp]; wh[k] = sm.prow(12 + k)[lp]; }
    double jp[2][3], jc[2][6], Q[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      jp[0][k] = a00 * g.R[k] + a02 * g.R[6 + k];
      jp[1][k] = a00 * g.R[3 + k] + a12 * g.R[6 + k];
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      Q[m][0] = jp[m][0] * hv[0] + jp[m][1] * hv[1] + jp[m][2] * hv[2];
      Q[m][1] = jp[m][0] * hv[1] + jp[m][1] * hv[3] + jp[m][2] * hv[4];
      Q[m][2] = jp[m][0] * hv[2] + jp[m][1] * hv[4] + jp[m][2] * hv[5];
    }
    if (ROT) {
      jc[0][0] = 2.0 * a02 * g.w[1]; jc[0][1] = 2.0 * (a00 * g.w[2] - a02 * g.w[0]); jc[0][2] = -2.0 * a00 * g.w[1];
      jc[1][0] = 2.0 * (a12 * g.w[1] - a00 * g.w[2]); jc[1][1] = -2.0 * a12 * g.w[0]; jc[1][2] = 2.0 * a00 * g.w[0];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { jc[0][k] = 0.0; jc[1][k] = 0.0; }
    }
    jc[0][3] = a00; jc[0][4] = 0.0; jc[0][5] = a02;
    jc[1][3] = 0.0; jc[1][4] = a00; jc[1][5] = a12;
    rec[0] = a00; rec[1] = a02; rec[2] = a12;
#pragma unroll
    for (int k = 0; k < 3; ++k) { rec[3 + k] = g.w[k]; rec[6 + k] = Q[0][k]; rec[9 + k] = Q[1][k]; rec[12 + k] = jp[0][k]; rec[15 + k] = jp[1][k]; }
    if (ROT) {   // rot-t cross block of F'F: (Jr' Jt)[r][c]
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) sv[(3 * r + c) * PSFM_SVS] = jc[0][r] * jc[0][3 + c] + jc[1][r] * jc[1][3 + c];
    }
    // per-image sums that need W = Jc' Jp and W H~ = Jc' Q contracted with a per-point 3-vector v:
    // (W v)[r] = jc0[r] (jp0.v) + jc1[r] (jp1.v),   (W H~ v)[r] = jc0[r] (Q0.v) + jc1[r] (Q1.v)
    if (a.intr >= 1) {
      const double zf = (g.w[2] + g.tz) * inv_f;
      const double jf0 = -a02 * zf, jf1 = -a12 * zf;
      const double qk0 = Q[0][0] * wkp[0] + Q[0][1] * wkp[1] + Q[0][2] * wkp[2];
      const double qk1 = Q[1][0] * wkp[0] + Q[1][1] * wkp[1] + Q[1][2] * wkp[2];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        sv[(9 + r) * PSFM_SVS] = jc[0][r] * jf0 + jc[1][r] * jf1;            // F'G
        sv[(15 + r) * PSFM_SVS] = -(jc[0][r] * qk0 + jc[1][r] * qk1);          // -(W H~) Wk'
      }
    }
    const double pw0 = jp[0][0] * wh[0] + jp[0][1] * wh[1] + jp[0][2] * wh[2];
    const double pw1 = jp[1][0] * wh[0] + jp[1][1] * wh[1] + jp[1][2] * wh[2];
#pragma unroll
    for (int r = (ROT ? 0 : 3); r < 6; ++r) sv[(21 + r) * PSFM_SVS] = -(jc[0][r] * pw0 + jc[1][r] * pw1);   // -(W w^)
  }
  __syncthreads();
  {
    double* dst = a.acc_cam + (size_t)(rep & (NREP - 1)) * a.rep_stride;
    tile_reduce_images<TILE>(sm, ti, NVX2, [&](int k, int img, double acc) {
      if (acc != 0.0) atomicAdd(dst + (size_t)img * NVX2 + k, acc);
    });
  }
  __syncthreads();
  // first entries of the first pass (three in flight per lane: one trip is shorter than an L2 round trip)
  unsigned int uf0 = 0u, uf1 = 0u, uf2 = 0u;
  {
    const int e = rg_f.x + par_f;
    if (e < rg_f.y) uf0 = __ldg(a.entries + e);
    if (e + lpt < rg_f.y) uf1 = __ldg(a.entries + e + lpt);
    if (e + 2 * lpt < rg_f.y) uf2 = __ldg(a.entries + e + 2 * lpt);
  }
  // the reduction rows are dead: the same shared memory now holds the per-observation records
  double* srec = sm.sv;
  if (act) {
    double2* o = reinterpret_cast<double2*>(srec + (size_t)tid * PS);
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
  }
  __syncthreads();
  double* band = a.Sband + (size_t)(rep & a.nrep_mask) * a.band_stride;
  // Pair tasks.  Block(i, j) = (W_i H~) W_j' = Jc_i' M Jc_j with the 2x2 M = Q_i Jp_j': per
  // entry 12 16-byte shared-memory loads (ncu, round 2: 64 % of the L1 data-pipe cycles with 24 8-byte loads at a
  // 19-double stride, the busiest unit of the kernel) and ~110 flops.  Two lanes per task (even | odd entries), the full 6x6 block in
  // registers, one shuffle per element to combine, 18 REDs per lane.  Uniform trip count.
  for (int q0 = 0; q0 < nq; q0 += TILE) {
    int tq = tq_f, par = par_f, e0 = rg_f.x, e1 = rg_f.y, slot = slot_f;
    bool valid = valid_f;
    unsigned int u0 = uf0, u1 = uf1, u2 = uf2;
    if (q0 > 0) {
      valid = lane_task(q0 + tid, tq, par);
      e0 = e1 = 0; u0 = u1 = u2 = 0u;
      if (valid) {
        const int2 rg = __ldg(a.task_rng + t0 + tq);
        slot = __ldg(a.task_slot + t0 + tq);
        e0 = rg.x; e1 = rg.y;
        const int e = e0 + par;
        if (e < e1) u0 = __ldg(a.entries + e);
        if (e + lpt < e1) u1 = __ldg(a.entries + e + lpt);
        if (e + 2 * lpt < e1) u2 = __ldg(a.entries + e + 2 * lpt);
      }
    }
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    for (int e = e0 + par; e < e1; e += lpt) {
      const unsigned int u = u0;
      u0 = u1; u1 = u2;
      u2 = (e + 3 * lpt < e1) ? __ldg(a.entries + e + 3 * lpt) : 0u;
      const double2* Pi = reinterpret_cast<const double2*>(srec + (size_t)(u >> 16) * PS);
      const double2* Pj = reinterpret_cast<const double2*>(srec + (size_t)(u & 0xffffu) * PS);
      // 12 LDS.128 per entry: [a00 a02 | a12 w0 | w1 w2 | Q (3 x 16 B)] of i, [a00 a02 | a12 w0 | w1 w2 | Jp (3 x 16 B)] of j
      double Ci[12], Cj[18];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double2 x = Pi[k]; Ci[2 * k] = x.x; Ci[2 * k + 1] = x.y; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double2 x = Pj[k]; Cj[2 * k] = x.x; Cj[2 * k + 1] = x.y; }
#pragma unroll
      for (int k = 6; k < 9; ++k) { const double2 x = Pj[k]; Cj[2 * k] = x.x; Cj[2 * k + 1] = x.y; }
      // M = Q_i Jp_j'
      double m00 = 0.0, m01 = 0.0, m10 = 0.0, m11 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double qa = Ci[6 + k], qb = Ci[9 + k], pa = Cj[12 + k], pb = Cj[15 + k];
        m00 = fma(qa, pa, m00); m01 = fma(qa, pb, m01); m10 = fma(qb, pa, m10); m11 = fma(qb, pb, m11);
      }
      // The camera Jacobian of an observation is Jc = [Jr | Jt] with Jt = [[a, 0, b], [0, a, c]] (a00, a02, a12)
      // and the rows of Jr twice the cross products w x (row of Jt), i.e. Jr' = 2 [w]x Jt'.  With
      // N = Jt_i' M Jt_j (3 x 3), u = w_i, v = w_j the 6 x 6 block is
      //     tt = N      rt = 2 [u]x N      tr = 2 B      rr = 4 [u]x B,     B[r, :] = v x N[r, :]
      // — 104 fp64 instructions per entry instead of 144 for the products with the explicit 2 x 6 Jacobians
      // (the fp64 pipe is what bounds this loop); the factors 2 and 4 are applied once per task.
      double N[3][3];
      {
        const double aj = Cj[0], bj = Cj[1], cj = Cj[2], ai = Ci[0], bi = Ci[1], ci = Ci[2];
        const double t00 = m00 * aj, t01 = m01 * aj, t02 = fma(m01, cj, m00 * bj);
        const double t10 = m10 * aj, t11 = m11 * aj, t12 = fma(m11, cj, m10 * bj);
        N[0][0] = ai * t00; N[0][1] = ai * t01; N[0][2] = ai * t02;
        N[1][0] = ai * t10; N[1][1] = ai * t11; N[1][2] = ai * t12;
        N[2][0] = fma(ci, t10, bi * t00); N[2][1] = fma(ci, t11, bi * t01); N[2][2] = fma(ci, t12, bi * t02);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[6 * (3 + r) + 3 + c] += N[r][c];
      if (ROT) {
        const double u0 = Ci[3], u1 = Ci[4], u2 = Ci[5], v0 = Cj[3], v1 = Cj[4], v2 = Cj[5];
        double B[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          B[r][0] = fma(v1, N[r][2], -(v2 * N[r][1]));
          B[r][1] = fma(v2, N[r][0], -(v0 * N[r][2]));
          B[r][2] = fma(v0, N[r][1], -(v1 * N[r][0]));
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[6 * (3 + r) + c] += B[r][c];                       // tr / 2
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          // rt / 2: column k of [u]x N;   rr / 4: column k of [u]x B
          acc[6 * 0 + 3 + k] = fma(u1, N[2][k], fma(-u2, N[1][k], acc[6 * 0 + 3 + k]));
          acc[6 * 1 + 3 + k] = fma(u2, N[0][k], fma(-u0, N[2][k], acc[6 * 1 + 3 + k]));
          acc[6 * 2 + 3 + k] = fma(u0, N[1][k], fma(-u1, N[0][k], acc[6 * 2 + 3 + k]));
          acc[6 * 0 + k] = fma(u1, B[2][k], fma(-u2, B[1][k], acc[6 * 0 + k]));
          acc[6 * 1 + k] = fma(u2, B[0][k], fma(-u0, B[2][k], acc[6 * 1 + k]));
          acc[6 * 2 + k] = fma(u0, B[1][k], fma(-u1, B[0][k], acc[6 * 2 + k]));
        }
      }
    }
    if (ROT) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { acc[6 * r + c] *= 4.0; acc[6 * r + 3 + c] *= 2.0; acc[6 * (3 + r) + c] *= 2.0; }
    }
    // combine the lanes of a task: lane `par` ends up with the sums of ITS share of the 36 elements (every
    // lane sends what its reader owns), then one RED per element
    double* dst = band + (size_t)slot * 36;
    const bool red = valid && !(a.dbg & 1);
    {
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        const double mine = par ? acc[18 + i] : acc[i];
        const double x = par ? acc[i] : acc[18 + i];
        const double v = mine + __shfl_xor_sync(0xffffffffu, x, 1);
        if (red && (ROT || v != 0.0)) atomicAdd(dst + 18 * par + i, v);
      }
    }
  }
}

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_schur_tile(const TileCtx tc, const StArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<TILE> sm;
  sm.carve(smem_raw, NVX2, 15, tc.cap_ns, tc.cap_np);
  const TileInfo ti = tile_header(tc);
  const int tid = threadIdx.x;
  const bool act = tid < ti.n;
  const size_t M = tc.M;
  const size_t i = (size_t)ti.base + tid;
  int ls = 0, lp = 0;
  double a00 = 0, a02 = 0, a12 = 0;
  if (act) {
    ls = __ldg(tc.obs_lseg + i);
    lp = __ldg(tc.obs_lpt + i);
    a00 = a.L.a[i]; a02 = a.L.a[M + i]; a12 = a.L.a[2 * M + i];
  }
  const int t0 = __ldg(a.tile_task + blockIdx.x), nt = __ldg(a.tile_task + blockIdx.x + 1) - t0;
  // per point: X (0..2), H~ (3..8), focal row of G'E (9..11), w^ (12..14)
  tile_fill_smem<TILE>(tc, sm, ti, a.pose16, nullptr, a.X, a.ht, a.wk, true, a.wt);
  schur_tile_body<TILE, ROT>(tc, a, sm, ti, act, ls, lp, a00, a02, a12, blockIdx.x, t0, nt);
}

// persistent, pipelined form (ba_tile_pipe.cuh): the inputs of the next tile are in flight
// (cp.async) while this tile's pair tasks run
template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 2 : 1)) k_schur_tile_p(const TileCtx tc, const PipeSrc ps, const StArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) int4 hdr_ring[4][2];
  __shared__ __align__(8) unsigned long long bars[2];
  const int cns = tc.cap_ns, cnp = tc.cap_np;
  const size_t sb = PipeStage<TILE>::bytes(false, true, 15, cns, cnp);
  TileSmem<TILE> sm;
  sm.cap_ns = cns; sm.cap_np = cnp;
  sm.sv = reinterpret_cast<double*>(smem_raw + 2 * sb);
  sm.sw = nullptr; sm.sred = nullptr; sm.sx = nullptr;
  pipe_run<TILE>(tc, ps, smem_raw, hdr_ring, bars, 15, cns, cnp, [&](const TileInfo& ti, const PipeStage<TILE>& s, int tile) {
    view_stage<TILE>(sm, s);
    const int tid = threadIdx.x;
    const bool act = tid < ti.n;
    int ls = 0, lp = 0;
    double a00 = 0, a02 = 0, a12 = 0;
    if (act) { ls = s.lseg[tid]; lp = s.lpt[tid]; a00 = s.a0[tid]; a02 = s.a1[tid]; a12 = s.a2[tid]; }
    const int t0 = __ldg(a.tile_task + tile), nt = __ldg(a.tile_task + tile + 1) - t0;
    schur_tile_body<TILE, ROT>(tc, a, sm, ti, act, ls, lp, a00, a02, a12, tile, t0, nt);
  });
}

template <int TILE>
inline size_t pipe_smem_schur_tile(int cns, int cnp) {
  return 2 * PipeStage<TILE>::bytes(false, true, 15, cns, cnp) + sizeof(double) * NVX2 * (TILE + 1);
}

// ---- tile-local pair structure (built once per problem)

// entries started by observation j (see k_pair_count); key = (tile, image a, image b), value =
// the two tile-local observation indices
__global__ void k_pair_fill_tile(const int* pt_ptr, const int* obs_pt, const int* obs_img, const int* ptr, int M,
                                 const int* tile_start, int T, int fbits, unsigned long long* keys, unsigned int* vals) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int lo = 0, hi = T;           // tile of j: last tile with tile_start <= j
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_start[mid] <= j) lo = mid; else hi = mid;
  }
  const int base = tile_start[lo];
  const int p = obs_pt[j];
  const int b = pt_ptr[p], e = pt_ptr[p + 1];
  const unsigned long long a = (unsigned long long)obs_img[j];
  const unsigned long long khi = ((unsigned long long)lo << (2 * fbits)) | (a << fbits);
  size_t o = (size_t)ptr[j];
  for (int k = j; k < e; ++k, ++o) {
    keys[o] = khi | (unsigned long long)obs_img[k];
    vals[o] = ((unsigned)(j - base) << 16) | (unsigned)(k - base);
  }
  for (int k = j - 1; k >= b && obs_img[k] == (int)a; --k, ++o) {
    keys[o] = khi | a;
    vals[o] = ((unsigned)(j - base) << 16) | (unsigned)(k - base);
  }
}

// first entry of every tile (entries are written in observation order, tiles are observation ranges)
__global__ void k_tile_entry_offsets(const int* tile_start, const int* ptr, int T, int* seg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= T) seg[t] = ptr[tile_start[t]];
}

// A run of equal keys (one image pair of one tile) is cut into UNITS of at most `chunk` entries, all
// of (nearly) the same length; a unit is what two lanes of k_schur_tile accumulate in registers.
// Without the cut the longest run of a tile (every point of the tile sees both images) sets the trip
// count of its warp while the other warps wait at the end-of-tile barrier.
__global__ void k_unit_count(const int* ucount, int nruns, int chunk, int* nunits) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nruns) return;
  nunits[t] = t < nruns ? (ucount[t] + chunk - 1) / chunk : 0;
}
// units of a tile are processed longest first (lanes of a warp then run similar trip counts and the
// short units fill the last pass): sort key (tile, ~count), payload = (slot, entry range)
__global__ void k_unit_fill(const unsigned long long* ukeys, const int* ucount, const int* beg, const int* ubeg, int nruns,
                            int fbits, int span, int by_pair, unsigned long long* key2, int* idx, int* slot, int2* rng) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nruns) return;
  const unsigned long long k = ukeys[t], mask = (1ull << fbits) - 1;
  const int b = (int)(k & mask), a = (int)((k >> fbits) & mask);
  const unsigned long long tile = k >> (2 * fbits);
  const int n = ubeg[t + 1] - ubeg[t], cnt = ucount[t], base = cnt / n, rem = cnt % n;
  int e = beg[t];
  for (int c = 0; c < n; ++c) {
    const int len = base + (c < rem), u = ubeg[t] + c;
    key2[u] = (tile << 32) | (by_pair ? (unsigned long long)(unsigned)u : (unsigned long long)(0xffffffffu - (unsigned)len));
    idx[u] = u;
    slot[u] = a * (span + 1) + (b - a);
    rng[u] = make_int2(e, e + len);
    e += len;
  }
}
__global__ void k_task_gather(const int* order, const int* slot_in, const int2* rng_in, int ntasks, int* slot_out, int2* rng_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntasks) return;
  const int o = order[t];
  slot_out[t] = slot_in[o];
  rng_out[t] = rng_in[o];
}

// tile -> first unit (lower bound over the sorted (tile, ~count) keys)
__global__ void k_tile_tasks(const unsigned long long* key2_sorted, int ntasks, int T, int* tile_task) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile > T) return;
  int lo = 0, hi = ntasks;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)(key2_sorted[mid] >> 32) < tile) lo = mid + 1; else hi = mid;
  }
  tile_task[tile] = lo;
}

// band blocks -> dense S:  S[6a+r][6b+c] -= s s' Sband[a][b-a][r][c]  (mirrored for a != b)
struct BandAsmArgs {
  const double* Sband;
  const double* scale_c;
  int F, span, lda;
  double* S;
};

__global__ void k_schur_assemble_band(const BandAsmArgs a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per_img = (size_t)(a.span + 1) * 36;
  if (t >= (size_t)a.F * per_img) return;
  const double x = a.Sband[t];
  if (x == 0.0) return;
  const int ia = (int)(t / per_img), rem = (int)(t % per_img), d = rem / 36, e = rem % 36, r = e / 6, c = e % 6;
  const int ib = ia + d;
  if (ib >= a.F) return;
  const size_t row = 6 * (size_t)ia + r, col = 6 * (size_t)ib + c;
  const double v = -a.scale_c[row] * a.scale_c[col] * x;
  a.S[row * a.lda + col] += v;
  if (d != 0) a.S[col * a.lda + row] += v;
}

// ------------------------------------------------------------------ assembly of the dense reduced system

struct AsmArgs {
  const double* Sblk;       // [nblocks][36]
  const int* blk_key;       // [nblocks] a * F + b  (a <= b)
  int nblocks;
  const double* lin_cam;    // [F][NVL]  rot F'F (6) | t F'F (6) | ...
  const double* lin_intr;   // [C][NVI]
  const double* prep_intr;  // [C][NVI]
  const double* xcam;       // [F][xstride] per-image sums of k_schur_w (NVX) / k_schur_tile (NVX2)
  int xstride;
  const double* scale_c;    // [NS]
  const double* Dc2;        // [NS]
  const unsigned char* active;
  const double* rhs;        // [NS]
  int F, C, NS;
  int lda;                  // leading dimension of S (NS + 1: the extra row carries the rhs)
  double* S;                // [lda][lda] row-major, zeroed
};

// off-diagonal / diagonal pair blocks: S[6a+r][6b+c] -= s s' Sblk (mirrored)
__global__ void k_schur_assemble_blocks(const AsmArgs a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)a.nblocks * 36) return;
  const int blk = (int)(t / 36), e = (int)(t % 36), r = e / 6, c = e % 6;
  const int ia = a.blk_key[blk] / a.F, ib = a.blk_key[blk] % a.F;
  const size_t row = 6 * (size_t)ia + r, col = 6 * (size_t)ib + c;
  const double v = -a.scale_c[row] * a.scale_c[col] * a.Sblk[t];
  atomicAdd(a.S + row * a.lda + col, v);
  if (ia != ib) atomicAdd(a.S + col * a.lda + row, v);
}

// rank-local per-image parts: rot-t cross block of F'F and the focal column
__global__ void k_schur_assemble_local(const AsmArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = a.lda, F = a.F;
  if (j >= F) return;
  const double* X = a.xcam + (size_t)j * a.xstride;
  const size_t s0 = 6 * (size_t)j;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double x = a.scale_c[s0 + r] * a.scale_c[s0 + 3 + c] * X[3 * r + c];
      if (x != 0.0) { atomicAdd(a.S + (s0 + r) * NS + s0 + 3 + c, x); atomicAdd(a.S + (s0 + 3 + c) * NS + s0 + r, x); }
    }
  const size_t sk = 6 * (size_t)F;   // single shared camera when intrinsics are free
  for (int r = 0; r < 6; ++r) {
    const double v = a.scale_c[s0 + r] * a.scale_c[sk] * (X[9 + r] + X[15 + r]);
    if (v != 0.0) { atomicAdd(a.S + (s0 + r) * NS + sk, v); atomicAdd(a.S + sk * NS + s0 + r, v); }
  }
}

// parts built from already all-reduced accumulators: F'F rot/t diagonal blocks, intrinsics block
__global__ void k_schur_assemble_global(const AsmArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = a.lda, F = a.F;
  const int ut[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  if (j < F) {
    const double* A = a.lin_cam + (size_t)j * NVL;
    const size_t s0 = 6 * (size_t)j;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        a.S[(s0 + r) * NS + s0 + c] += a.scale_c[s0 + r] * a.scale_c[s0 + c] * A[ut[r][c]];
        a.S[(s0 + 3 + r) * NS + s0 + 3 + c] += a.scale_c[s0 + 3 + r] * a.scale_c[s0 + 3 + c] * A[6 + ut[r][c]];
      }
  } else if (j < F + a.C) {
    const int c = j - F;
    const size_t sk = 6 * (size_t)F + 3 * c;
    const double* G = a.lin_intr + (size_t)c * NVI;
    const double* Gc = a.prep_intr + (size_t)c * NVI;
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc)
        a.S[(sk + r) * NS + sk + cc] += a.scale_c[sk + r] * a.scale_c[sk + cc] * (G[ut[r][cc]] + Gc[ut[r][cc]]);
  }
}

// LM diagonal on active slots, identity on inactive ones (their rows/cols are zero: scale 0)
__global__ void k_schur_assemble_finish(const AsmArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.NS) return;
  if (a.active[s]) a.S[(size_t)s * a.lda + s] += a.Dc2[s];
  else a.S[(size_t)s * a.lda + s] = 1.0;
  a.S[(size_t)a.NS * a.lda + s] = a.rhs[s];      // rhs as an extra (arrow) row: its factor row is L^-1 b
}

// ------------------------------------------------------------------ blocked, band-aware Cholesky (cooperative, multi-CTA)

// In-place lower Cholesky of the leading ns x ns part of A (row-major, leading dimension
// lda), where A[i][j] == 0 for |i - j| > bw among the first nb rows and rows nb..ns-1 are
// dense ("arrow": shared intrinsics).  Row ns of A holds the right-hand side b: it is
// carried through the factorisation as one more arrow row, so that its solved entries are
// L^-1 b; CTA 0 then solves L' x = L^-1 b.  One cooperative launch: per 32-column panel
// (1) every CTA factors the 32x32 diagonal block redundantly, (2) per trailing tile pair a CTA
// solves the two row tiles it needs against it and updates its tile, (3) ONE grid barrier.
// The factor lives in its own storage (Ld, Lp); A keeps the unsolved panels so that no CTA
// has to wait for another one's row solves.  fail[0] = 1 when a pivot is not positive.
constexpr int CB = 32;

struct CholArgs {
  double* A;
  int ns, lda, nb, bw;
  double* x;      // [ns]
  int* fail;
  unsigned int* bar;          // grid barrier counter, zeroed before the launch
  double* Lp;                 // [npanel][rmax][CB]  solved rows below each panel (the factor's off-diagonal part)
  double* Ld;                 // [npanel][CB][CB]    diagonal blocks of the factor
  int rmax;                   // rows reserved per panel in Lp
  unsigned long long* prof;   // optional [8]: SM cycles of CTA 0 per phase (PSFM_CHOL_PROFILE)
};

// Grid-wide barrier for the few (co-resident, cooperative launch) CTAs of k_chol_blocked: one
// atomic per CTA on a monotone counter and an acquire spin — about half the cost of
// cooperative_groups' grid.sync(), which this kernel pays twice per 32-column panel.
__device__ __forceinline__ void chol_grid_barrier(unsigned int* counter, unsigned int& target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ int chol_row_of(int pos, int c1, int nband, int arrow0) {
  return pos < nband ? c1 + pos : arrow0 + (pos - nband);
}

// Warp-level Cholesky of one w x w (w <= 32) diagonal block at A (leading dimension lda):
// lane r owns row r in registers, the block is padded to 32 x 32 with an identity tail.
// Writes L to sD and 1/L[k][k] to sDinv; returns true on a non-positive pivot.
__device__ __noinline__ bool chol_diag_warp(const double* __restrict__ A, int lda, int w,
                                            double (*sD)[CB + 1], double* sDinv) {
  const int lane = threadIdx.x & 31;
  double arow[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c)
    arow[c] = (lane < w && c <= lane) ? __ldcg(A + (size_t)lane * lda + c) : ((c == lane && lane >= w) ? 1.0 : 0.0);
  bool bad = false;
  double myinv = 1.0;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const double d = __shfl_sync(0xffffffffu, arow[j], j);
    if (!(d > 0.0) || isinf(d)) bad = true;
    const double idj = rsqrt(d);
    if (lane == j) { arow[j] = d * idj; myinv = idj; }
    else if (lane > j) arow[j] = arow[j] * idj;
    const double lr = arow[j];
#pragma unroll
    for (int c = j + 1; c < CB; ++c) {
      const double lc = __shfl_sync(0xffffffffu, lr, c);
      if (lane >= c) arow[c] -= lr * lc;
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) sD[lane][c] = arow[c];
  sDinv[lane] = myinv;
  return __any_sync(0xffffffffu, bad);
}

// One row of the panel below the diagonal block: x L11' = a (registers, right-looking so that the
// updates of one step are independent), from the (unsolved) global row into a shared-memory tile
// row and, when lprow != nullptr, into the factor storage.
__device__ __noinline__ void chol_row_solve_to(const double* __restrict__ grow, int w, const double (*sD)[CB + 1],
                                               const double* sDinv, double* __restrict__ srow, double* __restrict__ lprow) {
  double xr[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) xr[k] = (k < w) ? __ldcg(grow + k) : 0.0;
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    xr[k] *= sDinv[k];
#pragma unroll
    for (int m = k + 1; m < CB; ++m) xr[m] -= xr[k] * sD[m][k];
  }
#pragma unroll
  for (int k = 0; k < CB; ++k) srow[k] = xr[k];
  if (lprow) {
#pragma unroll
    for (int k = 0; k < CB; ++k) lprow[k] = xr[k];
  }
}

__global__ void __launch_bounds__(256) k_chol_blocked(const CholArgs a) {
  unsigned int bar_target = 0;
  __shared__ double sD[CB][CB + 1];
  __shared__ double sLi[CB][CB + 1];
  __shared__ double sLj[CB][CB + 1];
  __shared__ double sDinv[CB];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int ns = a.ns, lda = a.lda, nrows = a.ns + 1;     // rows incl. the rhs row
  double* A = a.A;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long tk = prof ? clock64() : 0;
#define PSFM_CHOL_TICK(slot)                                   \
  if (prof) {                                                  \
    const long long now_ = clock64();                          \
    a.prof[slot] += (unsigned long long)(now_ - tk);           \
    tk = now_;                                                 \
  }
  for (int c0 = 0; c0 < ns; c0 += CB) {
    const int w = min(CB, ns - c0), c1 = c0 + w;
    // ---- (1) diagonal block, redundantly per CTA: warp 0 factors it in registers
    if (tid < CB && chol_diag_warp(A + (size_t)c0 * lda + c0, lda, w, sD, sDinv)) s_bad = 1;
    __syncthreads();
    PSFM_CHOL_TICK(0);
    if (s_bad) break;     // uniform across the grid: every CTA factors the same block
    // ---- rows below the panel that can be non-zero
    const int rb = (c1 < a.nb) ? min(a.nb, c1 + a.bw) : c1;
    const int nband = max(0, rb - c1);
    const int arrow0 = max(a.nb, c1);
    const int npos = nband + (nrows - arrow0);
    const int p = c0 / CB;
    // the factor's diagonal block goes to its own storage (the unfactored block in A is still
    // being read by slower CTAs)
    if (blockIdx.x == 0)
      for (int t = tid; t < CB * CB; t += 256) a.Ld[(size_t)p * CB * CB + t] = sD[t / CB][t % CB];
    // ---- (2+3) per tile pair of the rows below the panel: this CTA solves the two row tiles
    //      it needs itself (X L11' = A21, from the UNSOLVED panel in A: nobody waits for a
    //      row-solve phase of another CTA), then updates its trailing tile.  Solved rows are
    //      stored to Lp by the CTA of pair (ti, 0).  ONE grid barrier per panel.
    const int ntile = (npos + CB - 1) / CB;
    const int npair = ntile * (ntile + 1) / 2;
    for (int pr = blockIdx.x; pr < npair; pr += gridDim.x) {
      int ti = 0;
      while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;          // pr -> (ti, tj), tj <= ti; a handful of tiles
      const int tj = pr - ti * (ti + 1) / 2;
      __syncthreads();
      if (tid < 2 * CB) {
        const int which = tid >> 5, lr = tid & 31;
        const int pos = (which ? tj : ti) * CB + lr;
        double* srow = which ? sLj[lr] : sLi[lr];
        if (!(which && ti == tj)) {
          if (pos < npos) {
            const int g = chol_row_of(pos, c1, nband, arrow0);
            double* lprow = (!which && tj == 0) ? a.Lp + ((size_t)p * a.rmax + pos) * CB : nullptr;
            chol_row_solve_to(A + (size_t)g * lda + c0, w, sD, sDinv, srow, lprow);
          } else {
#pragma unroll
            for (int k = 0; k < CB; ++k) srow[k] = 0.0;
          }
        }
      }
      double* dst[4];
      double old[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {          // this thread's four outputs: issue the loads early
        const int r = (tid + 256 * u) / CB, c = (tid + 256 * u) % CB;
        const int pi = ti * CB + r, pj = tj * CB + c;
        dst[u] = nullptr;
        old[u] = 0.0;
        if (pi < npos && pj < npos && pj <= pi) {
          const int gi = chol_row_of(pi, c1, nband, arrow0), gj = chol_row_of(pj, c1, nband, arrow0);
          if (gj < ns) {       // the rhs row has no column
            dst[u] = A + (size_t)gi * lda + gj;
            old[u] = __ldcg(dst[u]);
          }
        }
      }
      __syncthreads();
      if (ti == tj) {
        for (int t = tid; t < CB * CB; t += 256) sLj[t / CB][t % CB] = sLi[t / CB][t % CB];
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (dst[u]) {
          const int r = (tid + 256 * u) / CB, c = (tid + 256 * u) % CB;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four chains: a dependent fp64 op costs ~25-30 cycles
#pragma unroll
          for (int k = 0; k < CB; k += 4) {
            s0 = fma(sLi[r][k], sLj[c][k], s0);
            s1 = fma(sLi[r][k + 1], sLj[c][k + 1], s1);
            s2 = fma(sLi[r][k + 2], sLj[c][k + 2], s2);
            s3 = fma(sLi[r][k + 3], sLj[c][k + 3], s3);
          }
          *dst[u] = old[u] - ((s0 + s1) + (s2 + s3));
        }
      }
    }
    PSFM_CHOL_TICK(3);
    chol_grid_barrier(a.bar, bar_target);
    PSFM_CHOL_TICK(4);
  }
  if (blockIdx.x != 0) return;
  if (tid == 0) *a.fail = s_bad;
  if (s_bad) return;
  // ---- back substitution L' x = y (CTA 0, panel by panel from the end); y = L^-1 b is the
  //      solved rhs row, the last row below every panel in Lp
  __shared__ double sacc[8][CB];
  const int npanel = (ns + CB - 1) / CB;
  for (int t = tid; t < npanel * CB; t += 256) {
    const int p = t / CB, c = t % CB, c0 = p * CB, w = min(CB, ns - c0), c1 = c0 + w;
    if (c < w) {
      const int rb = (c1 < a.nb) ? min(a.nb, c1 + a.bw) : c1;
      const int npos_all = max(0, rb - c1) + (nrows - max(a.nb, c1));
      a.x[c0 + c] = __ldcg(a.Lp + ((size_t)p * a.rmax + (npos_all - 1)) * CB + c);
    }
  }
  __syncthreads();
  for (int p = npanel - 1; p >= 0; --p) {
    const int c0 = p * CB, w = min(CB, ns - c0), c1 = c0 + w;
    const int rb = (c1 < a.nb) ? min(a.nb, c1 + a.bw) : c1;
    const int nband = max(0, rb - c1);
    const int arrow0 = max(a.nb, c1);
    const int npos = nband + (ns - arrow0);          // rhs row excluded
    const double* Lp = a.Lp + (size_t)p * a.rmax * CB;
    // partial sums: column c of the panel, rows strided over the 8 row-groups
    const int c = tid % CB, g = tid / CB;
    for (int t = tid; t < CB * CB; t += 256) sD[t / CB][t % CB] = __ldcg(a.Ld + (size_t)p * CB * CB + t);
    double s = 0.0;
    if (c < w) {
      int pos = g;
      for (; pos + 24 < npos; pos += 32) {       // four independent loads in flight
        double av[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          av[u] = __ldcg(Lp + (size_t)(pos + 8 * u) * CB + c);
          xv[u] = a.x[chol_row_of(pos + 8 * u, c1, nband, arrow0)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s += av[u] * xv[u];
      }
      for (; pos < npos; pos += 8) s += __ldcg(Lp + (size_t)pos * CB + c) * a.x[chol_row_of(pos, c1, nband, arrow0)];
    }
    sacc[g][c] = s;
    __syncthreads();
    if (tid < w) {
      double t = 0.0;
      for (int gg = 0; gg < 8; ++gg) t += sacc[gg][tid];
      sacc[0][tid] = a.x[c0 + tid] - t;
    }
    __syncthreads();
    // in-panel backward substitution (warp 0; L11 staged in shared memory above)
    if (tid < CB) {
      double mine = (tid < w) ? sacc[0][tid] : 0.0;
      const double dinv = (tid < w) ? 1.0 / sD[tid][tid] : 0.0;
      for (int j = w - 1; j >= 0; --j) {
        const double v = __shfl_sync(0xffffffffu, mine * dinv, j);
        if (tid == j) a.x[c0 + j] = v;
        if (tid < j) mine -= sD[j][tid] * v;
      }
    }
    __syncthreads();
  }
  PSFM_CHOL_TICK(5);
#undef PSFM_CHOL_TICK
}

}  // namespace ba
}  // namespace psfm
