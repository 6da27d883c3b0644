// ba_band_chol.cuh — the reduced camera system of a VIDEO (banded + arrow) solved by ONE CTA.
//
// Exact-Schur mode (reference rule for <= 1000 images, bundle_adjustment.cc:276-286; Ceres
// SchurComplementSolver semantics): S y = rhs with S = 6F x 6F banded (half
// bandwidth bw = 6 * track span + 5) plus an "arrow" of the shared camera's 3 intrinsics slots.
// The factorisation is a chain of 6F dependent pivots: what bounds it is the latency of one
// pivot step, not flops (~4 MFLOP) — grid barriers through L2 (the round-1 kernel,
// k_chol_blocked: 38 panels x ~1.5 us) or cluster barriers (~380 cycles) are the wrong tool.
// Here the whole active window lives in the REGISTERS of one CTA:
//
//   k_band_assemble  band blocks + per-image sums -> compact scaled band matrix Ab (each entry
//                    written exactly once; the dense (6F+3)^2 S is never formed or zeroed)
//   k_band_chol      right-looking Cholesky on a sliding W x W window (W >= bw + 1, multiple
//                    of 4).  Index i lives at circular position i mod W; the symmetric window
//                    is held as unordered pairs of positions {p, q}, one 4 x 4 block of pairs
//                    per thread.  Per pivot j: the owners of column j publish it to shared
//                    memory, ONE __syncthreads, every thread applies the rank-1 update to its
//                    block (1/d recomputed redundantly: no second barrier), the freed slots are
//                    refilled with row j + W, which a streaming warp copies 4 pivots ahead with
//                    cp.async into an 8-row ring (already permuted to window positions); an
//                    output warp writes the finished column of L.  Each role runs its OWN small
//                    loop (bar.sync from three program counters).  The 4 arrow rows (3 intrinsics + the rhs, so
//                    that L^-1 b falls out of the same sweep) are one more block row; their 4x4
//                    corner one more thread.  Then L' x = y by warp 0 in axpy form (per pivot:
//                    one multiply, one shuffle, one fma on the chain), the rows of L staged
//                    chunk-wise into shared memory by the other warps.
//                    TWO-SIDED form (gridDim.x = 2, matrices of >= 4 windows): the chain of nb dependent
//                    pivots is cut in two.  CTA 0 factors the leading k0 pivots of A top-down, CTA 1
//                    the trailing n1 pivots bottom-up (the same code on the index-reversed matrix);
//                    the W rows in between (W > bw, so no top pivot touches a bottom pivot: the two
//                    eliminations commute) collect both Schur updates — CTA 1 hands its update of
//                    that W x W block, of the arrow and of the corner over through global memory,
//                    CTA 0 adds it to its register window and finishes the last W pivots and the
//                    corner.  The back substitution mirrors it: CTA 0 solves the middle rows first
//                    and releases them, then both CTAs walk outwards.  Same arithmetic per pivot as
//                    the one-sided form, different (but fixed) elimination order.
// tools/emu_band_chol.py is a thread-level numpy emulation of the one-sided window logic.
#pragma once
#include "ba_schur_explicit.cuh"

namespace psfm {
namespace ba {

constexpr int BC_RING = 8;        // ring of upcoming rows (streamed 4 pivots ahead)
constexpr int BC_MAXW = 152;      // window limit: (W/4 + 1)(W/4 + 2)/2 workers + W + 4 helper lanes <= 1024 threads
constexpr int BC_CSM = 184;       // FIXED shared-memory row stride of colbuf / ring (addresses become immediates):
                                  // bc_idx(BC_MAXW + 7) < 184
// Window position p lives at double index p + 2 (p / 16): 16 bytes of padding after every 128.
// A warp's lanes read 32 consecutive bytes each (their block of 4 positions) with two LDS.128; at a
// plain 32-byte stride the lanes q and q + 4 of a quarter warp hit the same banks (2-way conflict:
// measured, the workers were bound by shared-memory wavefronts), with the padding they do not.
__host__ __device__ __forceinline__ constexpr int bc_idx(int p) { return p + ((p >> 4) << 1); }

constexpr int BC_MAXSLOT = 7;     // 1 + ceil(bw / 32) register slots of the back substitution

// Split of the pivot chain (band_chol_plan): one-sided (two = 0) or two-sided.
struct BandPlan {
  int nb, bw, W, RS;
  int two;                  // 1: two CTAs
  int blk6;                 // 1: block-6 kernel k_band_chol6 (nb and bw + 1 multiples of 6, 3 <= W / 6 <= 25)
  int nbp;                  // nb rounded up to 8 (identity padding rows)
  int k0, n1;               // pivots factored top-down before the hand-over | bottom-up; k0 + W + n1 = nbp
  int nbs[2];               // rows of the matrix each side sees: k0 + W | n1 + W   (one-sided: nb | 0)
  int npiv[2];              // pivots each side factors:           k0 + W | n1       (one-sided: nb | 0)
  int rows[2];              // rows of Ab each side may touch (band_chol_rows)
};

struct BandAsmArgs2 {
  const double* Sband;      // [F][span + 1][36] all-reduced pair-block sums
  const double* lin_cam;    // [F][NVL]  F'F rot (6) | t (6) | ...
  const double* lin_intr;   // [C][NVI]
  const double* prep_intr;  // [C][NVI]
  const double* xcam;       // [F][xstride] rot-t cross (9) | F'G (6) | -(W H~) Wk' (6) | ...
  int xstride;
  const double* scale_c;    // [NS]
  const double* Dc2;        // [NS]
  const double* rhs;        // [NS]
  const unsigned char* active;
  int F, span;
  BandPlan pl;
  double* Ab;               // side 0 [rows[0]][RS]: Ab[r][k] = A[r][r - k] (k <= bw), Ab[r][W + a] = A[nb + a][r], a < 3; Ab[r][W + 3] = rhs[r]
  double* Ab1;              // side 1 [rows[1]][RS]: the same of the index-reversed matrix, middle block and middle arrow zero
  double* C4;               // [2][4][4] arrow corner: intrinsics block (3 x 3) | rhs entries in row / column 3; side 1: zero
  int* fail;                // cleared here, set by k_band_chol
};

// packed index of element (r, c) of a symmetric 3 x 3 stored as 00 01 02 11 12 22
__device__ __forceinline__ int sym3(int r, int c) {
  const int lo = min(r, c), hi = max(r, c);
  return lo * (5 - lo) / 2 + hi;
}

// scaled reduced-system element A[r][r - e] of the band part (0 <= e <= bw, r - e >= 0), identity below nb
__device__ __forceinline__ double band_entry(const BandAsmArgs2& a, int r, int e) {
  if (r >= a.pl.nb) return e == 0 ? 1.0 : 0.0;
  const int c = r - e;
  double v = 0.0;
  const int ia = r / 6, rr = r % 6, ib = c / 6, cc = c % 6, d = ia - ib;
  const double ss = a.scale_c[r] * a.scale_c[c];
  if (d <= a.span)     // lower element (r, c): mirror of the stored upper block (ib, ib + d)
    v = -ss * a.Sband[((size_t)ib * (a.span + 1) + d) * 36 + (d ? 6 * cc + rr : 6 * rr + cc)];
  if (d == 0) {
    const double* A = a.lin_cam + (size_t)ia * NVL;
    if (rr < 3) v += ss * A[sym3(rr, cc)];                                   // cc <= rr < 3
    else if (cc >= 3) v += ss * A[6 + sym3(rr - 3, cc - 3)];
    else v += ss * a.xcam[(size_t)ia * a.xstride + 3 * cc + (rr - 3)];     // (Jr' Jt)[cc][rr - 3]
    if (e == 0) { if (a.active[r]) v += a.Dc2[r]; else v = 1.0; }
  }
  return v;
}
// arrow entry aa of column r: A[nb + aa][r] (aa < 3) | rhs[r] (aa = 3); zero below nb
__device__ __forceinline__ double arrow_entry(const BandAsmArgs2& a, int r, int aa) {
  if (r >= a.pl.nb) return 0.0;
  if (aa == 3) return a.rhs[r];
  if (aa == 0) {
    const double* X = a.xcam + (size_t)(r / 6) * a.xstride;
    return a.scale_c[r] * a.scale_c[6 * (size_t)a.F] * (X[9 + r % 6] + X[15 + r % 6]);
  }
  return 0.0;
}

__global__ void k_band_assemble(const BandAsmArgs2 a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const BandPlan& pl = a.pl;
  if (t == 0) *a.fail = 0;
  if (t < 32) {   // corner (camera 0: intrinsics are only ever free for a single shared camera); side 1 starts from zero
    const int r = ((int)t & 15) / 4, c = (int)t % 4;
    const size_t sk = 6 * (size_t)a.F;
    double v = 0.0;
    if (t < 16) {
      if (r < 3 && c < 3) {
        v = a.scale_c[sk + r] * a.scale_c[sk + c] * (a.lin_intr[sym3(r, c)] + a.prep_intr[sym3(r, c)]);
        if (r == c) { if (a.active[sk + r]) v += a.Dc2[sk + r]; else v = 1.0; }
      } else if (r == 3 && c < 3) v = a.rhs[sk + c];
      else if (c == 3 && r < 3) v = a.rhs[sk + r];
    }
    a.C4[t] = v;
  }
  const int RS = pl.RS, W = pl.W;
  const size_t n0 = (size_t)pl.rows[0] * RS, n1 = (size_t)pl.rows[1] * RS;
  if (t < n0) {
    const int r = (int)(t / RS), e = (int)(t % RS);
    double v = 0.0;
    if (r >= min(pl.nb, pl.nbs[0])) {
      v = (e == 0) ? 1.0 : 0.0;                       // identity padding below the part this side factors
    } else if (e < W) {
      if (e <= pl.bw && r - e >= 0) v = band_entry(a, r, e);
    } else {
      v = arrow_entry(a, r, e - W);
    }
    a.Ab[t] = v;
  } else if (t < n0 + n1) {
    // index-reversed matrix B[r'][c'] = A[nbp-1-r'][nbp-1-c']: its lower element (r', r' - e) is the
    // lower element (R, R - e) of A with R = nbp - 1 - (r' - e).  Rows / columns >= n1 are the middle
    // block, which side 0 owns: zero here, so that what is left in the window is the pure update.
    const size_t u = t - n0;
    const int r = (int)(u / RS), e = (int)(u % RS);
    double v = 0.0;
    if (r >= pl.nbs[1]) {
      v = (e == 0) ? 1.0 : 0.0;
    } else if (e < W) {
      const int c = r - e;
      if (e <= pl.bw && c >= 0 && !(r >= pl.n1 && c >= pl.n1)) v = band_entry(a, pl.nbp - 1 - c, e);
    } else if (r < pl.n1) {
      v = arrow_entry(a, pl.nbp - 1 - r, e - W);
    }
    a.Ab1[u] = v;
  }
}

struct BandSide {
  const double* Ab;
  const double* C4;
  int nb, npiv;             // rows of this side's matrix | pivots it factors (multiple of 8 in the two-sided form)
  double* Lr;               // [nb][bw + 1]  UNNORMALISED columns: Lr[r][k] = A~[r][r - k] (= L[r][r - k] sqrt(d[r - k]))
  double* La;               // [4][nb]       unnormalised arrow rows (row 3 = rhs)
  double* dinv;             // [nb]          pivots d[j], replaced by 1 / sqrt(d[j]) after the factorisation
};

struct BandCholArgs {
  BandSide s[2];            // blockIdx.x = side
  int two, nbg, nbp, k0;    // two-sided | rows of the whole matrix | padded to 8 | pivots of side 0 before the hand-over
  int bw, W, RS, ns;        // ns: length of x (slots past nbg + 3 — other cameras — are zeroed)
  double* x;                // [ns]
  int* fail;                // OR of both sides (cleared by k_band_assemble)
  double* D;                // hand-over of side 1: [W][W] update of the middle block (A orientation) | [4][W] arrow | [16] corner
  int* sync;                // [2] = epoch once: 0 the hand-over is written, 1 the middle x and the intrinsics are written
  int epoch;
  int blk6;                 // k_band_chol6 (plan)
  int flags;                // timing experiments (PSFM_CHOL_FLAGS; results invalid): 1 no output of L, 2 no reciprocal, 4 no publish, 8 barrier + pivot load only, 16 no row streaming
  long long* prof;          // optional [8]: SM cycles of factorisation | corner + staging | back substitution, pivots (side 0)
};

__device__ __forceinline__ void bc_post(int* flag, int v) {
  __threadfence();
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ void bc_wait(const int* flag, int v) {
  int x;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(x) : "l"(flag) : "memory");
  } while (x != v);
}

// bar.sync 0 from role-specific loops: every thread of the CTA executes the same NUMBER of
// barriers, from different program counters.  Measured on B200: one warp runs dependent scalar
// code at ~5 cycles per instruction, so what a role does per pivot is counted in instructions —
// the first version (all roles interleaved in one unrolled body, index arithmetic per pivot)
// took 1900 cycles per pivot.  Hence: per-role loops, compile-time shared-memory offsets,
// one element per helper lane, no early exit.
__device__ __forceinline__ void bc_bar() { asm volatile("bar.sync 0;\n" ::: "memory"); }

// 1 / d to full double precision (not correctly rounded): MUFU seed + two Newton steps — about
// half the dependent latency of the IEEE division, which sits on the pivot-to-pivot chain
__device__ __forceinline__ double bc_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// Everything after the factorisation, shared by k_band_chol and k_band_chol6: normalisation, the 3 x 3 arrow corner
// (s_c4: the 4 x 4 corner block of the window, written to shared memory by its owner before the caller's barrier),
// hand-shake of the two-sided form, back substitution.  Called by every thread of the CTA.
template <int RP>
__device__ __forceinline__ void bc_tail(const BandCholArgs& a, const BandSide& sd, const int side, double* stage, int& s_fail,
                                        double* s_xI, const double* s_c4, const long long tk0, const long long tk1) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int nb = sd.nb, npiv = sd.npiv, bw = a.bw, LS = bw + 1;
  const int nsteps = npiv;
  // 1 / sqrt(d): normalisation of the stored columns, applied while staging the back substitution
  for (int j = tid; j < npiv; j += blockDim.x) sd.dinv[j] = rsqrt(__ldcg(sd.dinv + j));
  // ---- arrow corner: 3 x 3 intrinsics block and its right-hand side (thread of block {Wb, Wb})
  if (side == 0 && tid == 0) {
    bool cbad = false;
    const double m00 = s_c4[0], m10 = s_c4[4], m20 = s_c4[8], m11 = s_c4[5], m21 = s_c4[9], m22 = s_c4[10];
    cbad |= !(m00 > 0.0);
    const double l00 = sqrt(m00), l10 = m10 / l00, l20 = m20 / l00;
    double t = m11 - l10 * l10;
    cbad |= !(t > 0.0);
    const double l11 = sqrt(t), l21 = (m21 - l20 * l10) / l11;
    t = m22 - l20 * l20 - l21 * l21;
    cbad |= !(t > 0.0);
    const double l22 = sqrt(t);
    const double z0 = s_c4[12] / l00, z1 = (s_c4[13] - l10 * z0) / l11, z2 = (s_c4[14] - l20 * z0 - l21 * z1) / l22;
    const double x2 = z2 / l22, x1 = (z1 - l21 * x2) / l11, x0 = (z0 - l10 * x1 - l20 * x2) / l00;
    s_xI[0] = x0; s_xI[1] = x1; s_xI[2] = x2;
    if (cbad || !isfinite(x0 + x1 + x2)) s_fail = 1;
  }
  __syncthreads();
  if (tid == 0 && s_fail) atomicOr(a.fail, 1);
  if (side == 0) {
    if (s_fail) {
      if (a.two && tid == 0) bc_post(a.sync + 1, a.epoch);     // side 1 must not wait for ever
      return;
    }
    for (int s = a.nbg + tid; s < a.ns; s += blockDim.x) a.x[s] = (s < a.nbg + 3) ? s_xI[s - a.nbg] : 0.0;
  } else {
    // the middle x and the intrinsics come from side 0
    if (tid == 0) {
      bc_wait(a.sync + 1, a.epoch);
      for (int k = 0; k < 3; ++k) s_xI[k] = __ldcg(a.x + a.nbg + k);
    }
    __syncthreads();
  }

  // ---- back substitution L' x = y - La' x_I in axpy form (warp 0; the other warps stage).
  //      Lane l holds the running right-hand side of positions 32 (c - m) + l, m = 0 .. msv-1, of the
  //      current 32-column chunk c.  stage[jj][32 + k] = -L[r][r - k] for 1 <= k <= bw (r = 32 c + jj),
  //      zero elsewhere, so the inner step is one shared load and one fma per slot — no predicates;
  //      per pivot one fma, one shuffle and one fma are on the dependent chain.  Rows >= npiv (side 1:
  //      the middle rows) are not solved for: their x is known and only propagated.
  const int ms = 1 + (bw + 31) / 32;
  const int msv = ms <= 4 ? 4 : BC_MAXSLOT;
  const int LSP = 32 * (msv + 1);
  const int ctop = (nb + 31) / 32 - 1;
  const int cpost = (a.two && side == 0) ? a.k0 / 32 : -1;     // after this chunk every middle x is written
  const int CH = 32 * LSP;
  // one row per warp and pass; all loads of a row (<= 7 x 2 per lane) are issued before the first use — with a
  // dependent load pair per element the staging, not the substitution chain, set the pace (measured: 7.5 k cycles
  // per 32-row chunk against 1.3 k for the chain)
  auto stage_chunk = [&](int c, double* buf, int w0, int nw) {
    // RP rows per pass (2 in the block-6 kernel): with 8 warps (the block-6 kernel) a warp stages ~5 rows of a chunk, one global-memory
    // latency each would outlast the substitution chain of the chunk
    for (int j0 = w0; j0 < 32; j0 += RP * nw) {
      double lv[RP][BC_MAXSLOT], dv[RP][BC_MAXSLOT];
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int jj = j0 + h * nw, r = 32 * c + jj;
#pragma unroll
        for (int t = 0; t < BC_MAXSLOT; ++t) {
          const int k = lane + 32 * t;                       // element q = 32 (t + 1) + lane of the padded row
          const bool ok = jj < 32 && t < msv && k >= 1 && k <= bw && r < nb && r - k >= 0 && r - k < npiv;
          lv[h][t] = ok ? __ldcg(sd.Lr + (size_t)r * LS + k) : 0.0;
          dv[h][t] = ok ? __ldcg(sd.dinv + r - k) : 0.0;
        }
      }
#pragma unroll
      for (int h = 0; h < RP; ++h) {
        const int jj = j0 + h * nw;
        if (jj < 32) {
          buf[jj * LSP + lane] = 0.0;
#pragma unroll
          for (int t = 0; t < BC_MAXSLOT; ++t)
            if (t < msv) buf[jj * LSP + 32 * (t + 1) + lane] = -lv[h][t] * dv[h][t];
        }
      }
    }
  };
  const double xi0 = s_xI[0], xi1 = s_xI[1], xi2 = s_xI[2];
  auto y0 = [&](int i) -> double {
    if (i < 0 || i >= npiv) return 0.0;
    return __ldcg(sd.dinv + i) * (__ldcg(sd.La + 3 * (size_t)nb + i) -
           (xi0 * __ldcg(sd.La + i) + xi1 * __ldcg(sd.La + (size_t)nb + i) + xi2 * __ldcg(sd.La + 2 * (size_t)nb + i)));
  };
  const int wid = tid >> 5, nwarp = blockDim.x >> 5;
  stage_chunk(ctop, stage, wid, nwarp);
  double yy[BC_MAXSLOT];
#pragma unroll
  for (int m = 0; m < BC_MAXSLOT; ++m) yy[m] = (tid < 32 && m < ms) ? y0(32 * (ctop - m) + lane) : 0.0;
  __syncthreads();
  const long long tk2 = a.prof ? clock64() : 0;
  for (int c = ctop, n = 0; c >= 0; --c, ++n) {
    const double* buf = stage + (n & 1) * CH;
    if (tid >= 32) {
      if (c > 0) stage_chunk(c - 1, stage + ((n + 1) & 1) * CH, wid - 1, nwarp - 1);
    } else {
      const double fresh = (c > 0) ? y0(32 * (c - ms) + lane) : 0.0;     // slot ms - 1 of the next chunk
      const int jl = 32 * c + lane;
      const double dl = (jl < npiv) ? __ldcg(sd.dinv + jl) : 0.0;
      // global index of local row jl: side 0 jl, side 1 nbp - 1 - jl; rows past nbg are identity padding (x = 0)
      const int gl = side ? a.nbp - 1 - jl : jl;
      const bool inx = jl < nb && gl >= 0 && gl < a.nbg;
      const double xk = (jl >= npiv && inx) ? __ldcg(a.x + gl) : 0.0;   // known x (side 1: middle rows)
      const bool wr = jl < npiv && inx;
      double* xw = a.x + (inx ? gl : 0);
      const double* rp = buf + 31 * LSP + 32 + (31 - lane);             // &stage[jj][32 + jj - lane], jj = 31
      const int step = LSP + 1;
      if (msv == 4) {
#pragma unroll 8
        for (int jj = 31; jj >= 0; --jj) {
          const double xj = __shfl_sync(0xffffffffu, fma(yy[0], dl, xk), jj);
#pragma unroll
          for (int m = 0; m < 4; ++m) yy[m] = fma(rp[32 * m], xj, yy[m]);
          if (lane == jj && wr) *xw = xj;
          rp -= step;
        }
      } else {
#pragma unroll 4
        for (int jj = 31; jj >= 0; --jj) {
          const double xj = __shfl_sync(0xffffffffu, fma(yy[0], dl, xk), jj);
#pragma unroll
          for (int m = 0; m < BC_MAXSLOT; ++m) yy[m] = fma(rp[32 * m], xj, yy[m]);
          if (lane == jj && wr) *xw = xj;
          rp -= step;
        }
      }
#pragma unroll
      for (int m = 0; m < BC_MAXSLOT; ++m) yy[m] = (m == ms - 1) ? fresh : ((m + 1 < BC_MAXSLOT) ? yy[m + 1] : 0.0);
      if (c == cpost) __threadfence();
    }
    __syncthreads();
    if (c == cpost && tid == 0) bc_post(a.sync + 1, a.epoch);
  }
  if (a.prof && tid == 0 && side == 0) {
    const long long tk3 = clock64();
    a.prof[0] = tk1 - tk0; a.prof[1] = tk2 - tk1; a.prof[2] = tk3 - tk2; a.prof[3] = nsteps;
  }
}

// Roles: workers (one BS x BS block of window slots each), helper lanes (one element each:
// stream entry e of the upcoming rows into the ring with cp.async, 4 pivots ahead, and write
// entry e of the finished column to global memory).
template <int BS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_band_chol(const BandCholArgs a) {
  extern __shared__ __align__(16) double bc_smem[];
  __shared__ int s_fail;
  __shared__ double s_xI[4];
  __shared__ double s_c4[16];
  constexpr int UN = 8;                     // pivots per unrolled body: ring slot and colbuf parity are compile-time
  const int side = blockIdx.x;
  // this side's pointers in registers (indexing the parameter block with blockIdx.x would turn every use into a
  // dependent constant-bank load inside the pivot loop)
  BandSide sd;
  sd.Ab = side ? a.s[1].Ab : a.s[0].Ab; sd.C4 = side ? a.s[1].C4 : a.s[0].C4;
  sd.nb = side ? a.s[1].nb : a.s[0].nb; sd.npiv = side ? a.s[1].npiv : a.s[0].npiv;
  sd.Lr = side ? a.s[1].Lr : a.s[0].Lr; sd.La = side ? a.s[1].La : a.s[0].La; sd.dinv = side ? a.s[1].dinv : a.s[0].dinv;
  const int W = a.W, Wb = W / BS, RS = a.RS, nb = sd.nb, npiv = sd.npiv, bw = a.bw, LS = bw + 1;
  const int jswitch = (a.two && side == 0) ? a.k0 : -1;
  double* colbuf = bc_smem;                 // [2][BC_CSM]   pivot column by window position (double buffered)
  double* ring = colbuf + 2 * BC_CSM;       // [BC_RING][BC_CSM] upcoming rows, permuted to window positions
  double* stage = ring + BC_RING * BC_CSM;  // [2][32][LSP] coefficients of the back substitution
  const int tid = threadIdx.x;
  const int NT = (Wb + 1) * (Wb + 2) / 2;
  const int ldr0 = (NT + 31) & ~31;         // first helper thread; blockDim.x = ldr0 + 32 * ceil((W + 4) / 32)
  const bool worker = tid < NT, helper = tid >= ldr0;
  int P = 0, Q = 0;
  if (worker) {
    P = (int)((sqrtf(8.f * (float)tid + 1.f) - 1.f) * 0.5f);
    while ((P + 1) * (P + 2) / 2 <= tid) ++P;
    while (P * (P + 1) / 2 > tid) --P;
    Q = tid - P * (P + 1) / 2;
  }
  if (tid == 0) s_fail = 0;
  for (int t = tid; t < (2 + BC_RING) * BC_CSM; t += blockDim.x) colbuf[t] = 0.0;
  // initial window: indices 0 .. W-1
  double v[BS][BS];
#pragma unroll
  for (int i = 0; i < BS; ++i)
#pragma unroll
    for (int k = 0; k < BS; ++k) {
      double x = 0.0;
      if (worker) {
        if (P < Wb) {
          const int rp = BS * P + i, rq = BS * Q + k, hi = max(rp, rq), lo = min(rp, rq);
          x = __ldg(sd.Ab + (size_t)hi * RS + (hi - lo));
        } else if (Q < Wb) { if (i < 4) x = __ldg(sd.Ab + (size_t)(BS * Q + k) * RS + W + i); }
        else if (i < 4 && k < 4) x = __ldg(sd.C4 + 4 * i + k);
      }
      v[i][k] = x;
    }
  __syncthreads();
  // column 0
  if (worker && Q == 0) {
#pragma unroll
    for (int i = 0; i < BS; ++i) colbuf[bc_idx(BS * P) + i] = v[i][0];
  }
  const long long tk0 = a.prof ? clock64() : 0;
  const int nsteps = ((npiv + UN - 1) / UN) * UN;
  const int jsplit = jswitch >= 0 ? jswitch : nsteps;      // end of phase 0
  bool bad = false;

  if (helper) {
    // ---- element e: entry e of row r of Ab (column r - e) goes to window position (r - e) mod W of ring
    //      slot r mod 8; entries W .. W+3 are the arrow and keep their position.  After the barrier of
    //      pivot j, position e of the pivot column is entry (e - pj) mod W of column j of L (unnormalised).
    const int e = tid - ldr0;
    const bool band = e < W, live = e < W + 4;
    const double* src = sd.Ab + (size_t)W * RS + e;       // row W
    int pos = band ? (e == 0 ? 0 : W - e) : e;            // (W - e) mod W: position of entry e of row W
    // rows travel global -> register (4 pivots ahead) -> ring: plain loads, NOT cp.async — a pending
    // LDGSTS is a pending shared-memory write, and bar.sync drains those (measured: ~600 cycles / pivot)
    double rg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { rg[i] = live ? __ldg(src) : 0.0; src += RS; }
    int kk = e;                                            // (e - pj) mod W for band entries
    const bool no_out = (a.flags & 1) != 0, no_ring = (a.flags & 16) != 0;
    const int ce = bc_idx(e);                              // where position e of the pivot column lives
    // entry (j + kk, kk) of Lr: one element back per pivot, W (LS + 1) forward when kk wraps
    double* lp = sd.Lr + (size_t)e * LS + e;
    double* la = sd.La + (size_t)(live && !band ? e - W : 0) * nb;
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0;
    // two phases (before | after the hand-over of the two-sided form) around ONE copy of the pivot loop:
    // the hand-over code stays out of the loop body (instruction cache), the one-sided form has an empty phase 1
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1 && jswitch >= 0) { bc_bar(); bc_bar(); }    // hand-over of side 1 (workers, below)
    const int jend = phase == 0 ? jsplit : nsteps;
#pragma unroll 1
    for (int j0 = phase == 0 ? 0 : jsplit; j0 < jend; j0 += UN) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u;
        if (!no_ring) {
          if (live) ring[u * BC_CSM + bc_idx(pos)] = rg[u & 3];   // row j + W -> slot u
          rg[u & 3] = live ? __ldg(src) : 0.0;                    // row j + W + 4
        }
        src += RS;
        if (band && ++pos == W) pos = 0;
        if (a.prof) { const long long t = clock64(); p_own += t - tlast; tlast = t; }
        bc_bar();
        if (a.prof) {
          const double dd = colbuf[(u & 1) * BC_CSM];        // first use after the barrier: the wait shows up here
          const long long t = clock64();
          p_wait += t - tlast + (dd == 1.25e-300 ? 1 : 0); tlast = t;
        }
        if (j < nb && !no_out) {
          const double val = colbuf[(u & 1) * BC_CSM + ce];
          if (band) {
            if (kk <= bw && j + kk < nb) *lp = val;
            if (kk == 0) sd.dinv[j] = val;                 // the pivot itself
          } else if (live) {
            la[j] = val;
          }
        }
        if (band) {
          if (--kk < 0) { kk = W - 1; lp += (size_t)W * (LS + 1); }
          lp -= 1;
        }
      }
    }
    }
    if (a.prof && tid == ldr0) { a.prof[4] = p_own; a.prof[5] = p_wait; }
  } else {
    // ---- workers (and idle threads of the last worker warp: barriers only)
    const double* sP = colbuf + bc_idx(BS * P);
    const double* sQ = colbuf + bc_idx(BS * Q);
    int pj0 = 0, pjp = 0;                                  // pivot position of step u = 0 of the body, and its padded index
    int Pj = 0;
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 1 && jswitch >= 0) {
        // two-sided form: the window now holds rows k0 .. k0 + W - 1 with the updates of the pivots above;
        // add side 1's updates of the same rows (pivots below), then publish column k0 again
        if (tid == 0) bc_wait(a.sync, a.epoch);
        bc_bar();
        if (worker) {
          const int kw = a.k0 % W;
          const double* DA = a.D + (size_t)W * W;
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int k = 0; k < BS; ++k) {
              const int mk = (BS * Q + k - kw + W) % W;      // position -> row of the middle block
              if (P < Wb) v[i][k] += __ldcg(a.D + (size_t)((BS * P + i - kw + W) % W) * W + mk);
              else if (Q < Wb) { if (i < 4) v[i][k] += __ldcg(DA + (size_t)i * W + mk); }
              else if (i < 4 && k < 4) v[i][k] += __ldcg(DA + 4 * (size_t)W + 4 * i + k);
            }
          if (Q == Pj) {
#pragma unroll
            for (int i = 0; i < BS; ++i) colbuf[bc_idx(BS * P) + i] = v[i][0];
          } else if (P == Pj) {
#pragma unroll
            for (int k = 0; k < BS; ++k) colbuf[bc_idx(BS * Q) + k] = v[0][k];
          }
        }
        bc_bar();
      }
    const int jend = phase == 0 ? jsplit : nsteps;
#pragma unroll 1
    for (int j0 = phase == 0 ? 0 : jsplit; j0 < jend; j0 += UN) {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int ij = u % BS;
        const int par = (u & 1) * BC_CSM, parn = ((u + 1) & 1) * BC_CSM, slot = u * BC_CSM;
        if (a.prof) { const long long t = clock64(); p_own += t - tlast; tlast = t; }
        bc_bar();
        const double d = colbuf[par + pjp + u];            // 8 consecutive positions never straddle a padding gap
        bad |= !(d > 0.0 && d <= 1.7976931348623157e308);
        if (a.prof) { const long long t = clock64(); p_wait += t - tlast + (bad ? 0 : 0); tlast = t; }
        if (worker && !(a.flags & 8)) {
          const double invd = (a.flags & 2) ? 1.0 - 1e-3 * d : bc_rcp(d);
          double cp[BS], tq[BS];
#pragma unroll
          for (int i = 0; i < BS; i += 2) {
            const double2 x = *reinterpret_cast<const double2*>(sP + par + i);
            const double2 y = *reinterpret_cast<const double2*>(sQ + par + i);
            cp[i] = x.x; cp[i + 1] = x.y;
            tq[i] = y.x * invd; tq[i + 1] = y.y * invd;
          }
          // rank-1 update
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int k = 0; k < BS; ++k) v[i][k] = fma(-cp[i], tq[k], v[i][k]);
          // the slots of position pj are free: they take row j + W (only their ~W/4 owners touch the ring)
          if (Q == Pj) {
#pragma unroll
            for (int i = 0; i < BS; i += 2) {
              const double2 z = *reinterpret_cast<const double2*>(sP + 2 * BC_CSM + slot + i);
              v[i][ij] = z.x; v[i + 1][ij] = z.y;
            }
          }
          if (P == Pj) {
#pragma unroll
            for (int k = 0; k < BS; k += 2) {
              const double2 w = *reinterpret_cast<const double2*>(sQ + 2 * BC_CSM + slot + k);
              v[ij][k] = w.x; v[ij][k + 1] = w.y;
            }
          }
          // publish column j + 1
          const int ijn = (ij + 1) % BS;
          int Pjn = Pj;
          if (ij == BS - 1) { Pjn = Pj + 1; if (Pjn == Wb) Pjn = 0; }
          double* cbn = colbuf + parn;
          if (a.flags & 4) {
          } else if (Q == Pjn) {
#pragma unroll
            for (int i = 0; i < BS; i += 2) *reinterpret_cast<double2*>(cbn + bc_idx(BS * P) + i) = make_double2(v[i][ijn], v[i + 1][ijn]);
          } else if (P == Pjn) {
#pragma unroll
            for (int k = 0; k < BS; k += 2) *reinterpret_cast<double2*>(cbn + bc_idx(BS * Q) + k) = make_double2(v[ijn][k], v[ijn][k + 1]);
          }
          if (ij == BS - 1) Pj = Pjn;
        } else if (ij == BS - 1) {
          if (++Pj == Wb) Pj = 0;
        }
      }
      pj0 += UN;
      if (pj0 == W) pj0 = 0;
      pjp = bc_idx(pj0);
    }
    }
    if (a.prof && tid == 0) { a.prof[6] = p_own; a.prof[7] = p_wait; }
  }
  __syncthreads();
  const long long tk1 = a.prof ? clock64() : 0;
  if (a.prof && tid == 0 && side == 0) { a.prof[0] = tk1 - tk0; a.prof[3] = nsteps; }
  if (bad) s_fail = 1;
  if (a.two && side == 1) {
    // hand-over: what is left in the window (rows n1 .. n1 + W - 1 of the reversed matrix, zero on input)
    // is the update of the middle block by this side's pivots; reversed row n1 + m is middle row W - 1 - m
    if (worker) {
      const int kw = npiv % W;
      double* DA = a.D + (size_t)W * W;
#pragma unroll
      for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          const int mk = W - 1 - (BS * Q + k - kw + W) % W;
          if (P < Wb) {
            const int mi = W - 1 - (BS * P + i - kw + W) % W;
            a.D[(size_t)mi * W + mk] = v[i][k];
            if (P != Q) a.D[(size_t)mk * W + mi] = v[i][k];
          } else if (Q < Wb) { if (i < 4) DA[(size_t)i * W + mk] = v[i][k]; }
          else if (i < 4 && k < 4) DA[4 * (size_t)W + 4 * i + k] = v[i][k];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) bc_post(a.sync, a.epoch);
  }
  if (worker && tid == NT - 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) s_c4[4 * i + k] = v[i][k];
  }
  __syncthreads();
  bc_tail<1>(a, sd, side, stage, s_fail, s_xI, s_c4, tk0, tk1);
}

// ------------------------------------------------------------------ block-6 form with look-ahead (k_band_chol6)
//
// The rank-1 kernel above pays ~450 cycles per pivot: ~65 instructions per lone worker warp and one CTA barrier
// for 16 FMAs per thread.  The reduced camera system is made of 6 x 6 image blocks, so the natural unit is a
// BLOCK pivot: per step k one 6 x 6 diagonal block is factored, the 6-column panel below it solved and the
// trailing window updated with a rank-6 product — one barrier per six pivots, and the two dependent chains
// (factor + solve of the next panel | rank-6 update of the window) run side by side in different warps:
//   workers   one thread per unordered pair {P, Q} of ring POSITIONS (block index I lives at position I mod Wb,
//             position Wb = the arrow rows), the 6 x 6 block T{P,Q}[a][b] = S[6 I_P + a][6 I_Q + b] in registers.
//             Step k (pivot position c): every block that touches neither c nor cn = c + 1 takes the rank-6 update
//             with panel k (shared memory); blocks touching c are dead (their content was panel k) and reload the
//             incoming block row k + Wb straight from global memory — a full step ahead of their next use;
//             blocks touching cn are dormant: the panel group carries that column.  At the end of the step the
//             blocks touching cn2 = c + 2 (updated through pivot k) are published for the panel group.
//   panel     one thread per row (position, a) of the next pivot column: takes the published row (or, for a freshly
//             recycled position, the raw row it prefetched from global memory), applies update k itself (its own
//             row of L_k is still in its registers), the 6 rows of the diagonal block are shared, EVERY panel
//             thread factors the 6 x 6 block redundantly (LDL' recurrence: reciprocal, not square root, on the
//             chain) and solves its own row on the fly; the normalised row goes to shared memory for the next
//             step, the unnormalised one to global memory in the layout bc_tail expects.
// One barrier of the 6 - 8 participating warps per step (bar.sync 1), one among the 3 - 5 panel warps (bar.sync 2);
// the remaining warps of the CTA only exist for the back substitution's staging and sleep at the final barrier.
// Phases end with a FLUSH step (no next panel: the workers update the dormant blocks too), after which the whole
// window is current: that is where the two-sided form hands over and where a phase restarts.
// Dataflow emulated block-wise against numpy before it was written; tests/test_gpu_band_chol.py.
constexpr int B6_PBS = 38;        // doubles per 6 x 6 panel block in shared memory (304 bytes: the blocks of eight
                                  // consecutive positions start in eight different groups of four banks)

__device__ __forceinline__ void bc_gbar(int n) { asm volatile("bar.sync 1, %0;\n" ::"r"(n) : "memory"); }
__device__ __forceinline__ void bc_pbar(int n) { asm volatile("bar.sync 2, %0;\n" ::"r"(n) : "memory"); }

template <int MAXT, int TR, int TC>
__global__ void __launch_bounds__(MAXT, 1) k_band_chol6(const BandCholArgs a) {
  // TR x TC: tile of a worker thread inside a 6 x 6 block (6 x 6: one thread per block; 3 x 6: two; 3 x 3: four).  Measured on B200
  // (PSFM_CHOL_PROFILE): a warp executes ~1 instruction per 4.6 cycles whatever the dependences, so a step costs
  // what its LONGEST warp executes — many thin threads beat few fat ones as long as the CTA has room for them.
  extern __shared__ __align__(16) double bc_smem[];
  __shared__ int s_fail;
  __shared__ double s_xI[4];
  __shared__ double s_c4[16];
  constexpr int TPB = (6 / TR) * (6 / TC);
  const int side = blockIdx.x;
  BandSide sd;
  sd.Ab = side ? a.s[1].Ab : a.s[0].Ab; sd.C4 = side ? a.s[1].C4 : a.s[0].C4;
  sd.nb = side ? a.s[1].nb : a.s[0].nb; sd.npiv = side ? a.s[1].npiv : a.s[0].npiv;
  sd.Lr = side ? a.s[1].Lr : a.s[0].Lr; sd.La = side ? a.s[1].La : a.s[0].La; sd.dinv = side ? a.s[1].dinv : a.s[0].dinv;
  const int W = a.W, Wb = W / 6, RS = a.RS, nb = sd.nb, LS = a.bw + 1;
  const int nsteps = sd.npiv / 6;
  const int jsw = (a.two && side == 0) ? a.k0 / 6 : -1;
  const int tid = threadIdx.x;
  const int nblk = (Wb + 1) * (Wb + 2) / 2;
  const int NWT = nblk * TPB;                   // worker threads
  const int NW = (NWT + 31) & ~31;
  const int NPR = 6 * (Wb + 1), NP = (NPR + 31) & ~31;
  const int NG = NW + NP;
  const int NG2 = (int)blockDim.x;       // the step barrier includes the loader warps (all remaining warps of the CTA)
  const int PB = (Wb + 1) * B6_PBS;
  // panel of step k in buffer k & 1, UNNORMALISED: Y = rows of the reduced pivot column (y), Z = y / d — the update is
  // T -= Z_P Y_Q', no square root anywhere in the loop (bc_tail normalises what goes to the back substitution)
  double* Ypan = bc_smem;               // [2][PB]
  double* Zpan = Ypan + 2 * PB;         // [2][PB]
  double* Raw = Zpan + 2 * PB;          // [2][PB] rows of the next pivot column, published at the end of step k in buffer k & 1
  double* Raw0 = Raw + 2 * PB;          // [PB]    rows of the pivot column a phase starts with
  double* Dsh = Raw0 + PB;              // [36]    diagonal block being factored
  double* ring = Dsh + 36;              // [4][6 RS] incoming block rows: slot k & 3 holds block row k + Wb (raw rows of Ab) during step k
  const int RB = 6 * RS;
  const double* __restrict__ Ab = sd.Ab;
  if (tid == 0) s_fail = 0;
  const long long tk0 = a.prof ? clock64() : 0;
  bool bad = false;
  const int kb0 = 0, ke0 = jsw >= 0 ? jsw : nsteps;     // phase 0; phase 1 (two-sided, side 0): jsw .. nsteps

  if (tid < NW) {
    // ------------------------------------------------------------ workers
    const bool act = tid < NWT;
    const int blk = tid / TPB, sub = tid % TPB;
    const int sr = TR * (sub / (6 / TC)), sc = TC * (sub % (6 / TC));   // tile origin inside the 6 x 6 block
    int P = 0, Q = 0;
    if (act) {
      P = (int)((sqrtf(8.f * (float)blk + 1.f) - 1.f) * 0.5f);
      while ((P + 1) * (P + 2) / 2 <= blk) ++P;
      while (P * (P + 1) / 2 > blk) --P;
      Q = blk - P * (P + 1) / 2;
    }
    double T[TR][TC];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int k = 0; k < TC; ++k) {
        double x = 0.0;
        const int bi = sr + i, bk = sc + k;
        if (act) {
          if (P < Wb) {
            const int r = 6 * P + bi, c = 6 * Q + bk, hi = max(r, c), lo = min(r, c);
            x = __ldg(Ab + (size_t)hi * RS + (hi - lo));
          } else if (Q < Wb) { if (bi < 4) x = __ldg(Ab + (size_t)(6 * Q + bk) * RS + W + bi); }
          else if (bi < 4 && bk < 4) x = __ldg(sd.C4 + 4 * bi + bk);
        }
        T[i][k] = x;
      }
    // rows of pivot column `pos` held by this tile -> dst[row position][a][0..5]
    auto publish = [&](double* dst, int pos) {
      if (Q == pos) {
        double* o = dst + P * B6_PBS + 6 * sr + sc;
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int k = 0; k < TC; ++k) o[6 * i + k] = T[i][k];
      } else if (P == pos) {
        double* o = dst + Q * B6_PBS + 6 * sc + sr;
#pragma unroll
        for (int k = 0; k < TC; ++k)
#pragma unroll
          for (int i = 0; i < TR; ++i) o[6 * k + i] = T[i][k];
      }
    };
    // T -= Z_P Y_Q' with the panel in shared memory (rows of 6 doubles, 16-byte aligned)
    auto update = [&](const double* Yb, const double* Zb) {
      const double2* lq = reinterpret_cast<const double2*>(Yb + Q * B6_PBS + 6 * sc);
      const double2* lp = reinterpret_cast<const double2*>(Zb + P * B6_PBS + 6 * sr);
      double q[TC][6];
#pragma unroll
      for (int k = 0; k < TC; ++k)
#pragma unroll
        for (int m = 0; m < 3; ++m) { const double2 z = lq[3 * k + m]; q[k][2 * m] = z.x; q[k][2 * m + 1] = z.y; }
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        double pr[6];
#pragma unroll
        for (int m = 0; m < 3; ++m) { const double2 z = lp[3 * i + m]; pr[2 * m] = z.x; pr[2 * m + 1] = z.y; }
#pragma unroll
        for (int k = 0; k < TC; ++k)
#pragma unroll
          for (int m = 0; m < 6; ++m) T[i][k] = fma(-pr[m], q[k][m], T[i][k]);
      }
    };
    // position c takes block row k + Wb (raw entries: no pivot <= k reaches it), staged in ring slot k & 3 by the
    // loader warp; cn = (k + 1) mod Wb holds block k + 1
    auto recycle = [&](int k, int c, int cn) {
      const int rn = 6 * (k + Wb);
      const double* rg = ring + (k & 3) * RB;            // rg[i * RS + e] = A[rn + i][rn + i - e]
      if (P == c && Q == c) {
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int kk = 0; kk < TC; ++kk) {
            const int bi = sr + i, bk = sc + kk;
            T[i][kk] = rg[max(bi, bk) * RS + (bi > bk ? bi - bk : bk - bi)];
          }
      } else if (P == Wb) {
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int kk = 0; kk < TC; ++kk) T[i][kk] = (sr + i < 4) ? rg[(sc + kk) * RS + W + sr + i] : 0.0;
      } else {
        const int o = (P == c) ? Q : P;
        int dI = o - cn; if (dI < 0) dI += Wb;
        const int ro = 6 * (k + 1 + dI);
        if (P == c) {
#pragma unroll
          for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int kk = 0; kk < TC; ++kk) T[i][kk] = rg[(sr + i) * RS + (rn + sr + i - ro - sc - kk)];
        } else {
#pragma unroll
          for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int kk = 0; kk < TC; ++kk) T[i][kk] = rg[(sc + kk) * RS + (rn + sc + kk - ro - sr - i)];
        }
      }
    };
    // column block k of the factor (panel k, unnormalised) -> global memory in the layout bc_tail reads; one
    // element per worker thread and pass (the workers have slack: the panel group is the longer chain)
    auto output = [&](int k, int c) {
      const double* Yb = Ypan + (k & 1) * PB;
      const int jb = 6 * k;
      for (int v = tid; v < 36 * (Wb + 1); v += NW) {
        const int p = v / 36, e = v - 36 * p, ar = e / 6, m = e - 6 * ar;
        const double y = Yb[p * B6_PBS + e];
        if (p == c) {
          if (m <= ar) sd.Lr[(size_t)(jb + ar) * LS + (ar - m)] = y;
          if (m == ar) sd.dinv[jb + ar] = y;
        } else if (p < Wb) {
          int dI = p - c; if (dI < 0) dI += Wb;
          const int rr = 6 * (k + dI) + ar, kk = rr - jb - m;
          if (rr < nb && kk < LS) sd.Lr[(size_t)rr * LS + kk] = y;          // beyond bw: structurally zero
        } else if (ar < 4) {
          sd.La[(size_t)ar * nb + jb + m] = y;
        }
      }
    };
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0, q_upd = 0, q_rec = 0, q_pub = 0;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      const int kb = phase == 0 ? kb0 : jsw, ke = phase == 0 ? ke0 : nsteps;
      if (phase == 1) {
        if (jsw < 0) break;
        // two-sided form: the window is current through pivot k0 - 1; add side 1's update of the same rows
        if (tid == 0) bc_wait(a.sync, a.epoch);
        bc_gbar(NG2);
        if (act) {
          const int cb = jsw % Wb;
          const double* DA = a.D + (size_t)W * W;
          int dP = P - cb; if (dP < 0) dP += Wb;
          int dQ = Q - cb; if (dQ < 0) dQ += Wb;
#pragma unroll
          for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int k = 0; k < TC; ++k) {
              const int bi = sr + i, bk = sc + k;
              if (P < Wb) T[i][k] += __ldcg(a.D + (size_t)(6 * dP + bi) * W + 6 * dQ + bk);
              else if (Q < Wb) { if (bi < 4) T[i][k] += __ldcg(DA + (size_t)bi * W + 6 * dQ + bk); }
              else if (bi < 4 && bk < 4) T[i][k] += __ldcg(DA + 4 * (size_t)W + 4 * bi + bk);
            }
        }
      }
      // (re)start: the whole window is current; hand the pivot column and the next one to the panel group
      int c = kb % Wb;
      {
        const int cn = c + 1 == Wb ? 0 : c + 1;
        if (act) {
          publish(Raw0, c);
          if (P != c && Q != c) publish(Raw + ((kb - 1) & 1) * PB, cn);
        }
        bc_gbar(NG2);
        bc_gbar(NG2);
      }
#pragma unroll 1
      for (int k = kb; k < ke; ++k) {
        const bool last = k == ke - 1;
        const int cn = c + 1 == Wb ? 0 : c + 1, cn2 = cn + 1 == Wb ? 0 : cn + 1;
        if (act) {
          const bool ic = (P == c || Q == c), icn = (P == cn || Q == cn);
          long long t0 = a.prof ? clock64() : 0;
          if (ic) { if (!icn || last) recycle(k, c, cn); }
          else if (!icn || last) update(Ypan + (k & 1) * PB, Zpan + (k & 1) * PB);
          if (a.prof) { const long long t1 = clock64(); if (ic) q_rec += t1 - t0; else q_upd += t1 - t0; t0 = t1; }
          if (!last && !ic && !icn && (P == cn2 || Q == cn2)) publish(Raw + (k & 1) * PB, cn2);
          if (a.prof) q_pub += clock64() - t0;
        }
        output(k, c);
        if (a.prof) { const long long t = clock64(); p_own += t - tlast; tlast = t; }
        bc_gbar(NG2);
        if (a.prof) { const long long t = clock64(); p_wait += t - tlast; tlast = t; }
        c = cn;
      }
    }
    if (a.prof && tid == 0 && side == 0) { a.prof[6] = p_own; a.prof[7] = p_wait; a.prof[8] = q_upd; a.prof[9] = q_rec; a.prof[10] = q_pub; }
    if (act && a.two && side == 1) {
      // hand-over: what is left in the window (rows n1 .. n1 + W - 1 of the reversed matrix, zero on input) is the
      // update of the middle block by this side's pivots; reversed row n1 + m is middle row W - 1 - m
      const int cb = nsteps % Wb;
      double* DA = a.D + (size_t)W * W;
      int dP = P - cb; if (dP < 0) dP += Wb;
      int dQ = Q - cb; if (dQ < 0) dQ += Wb;
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int k = 0; k < TC; ++k) {
          const int bi = sr + i, bk = sc + k;
          const int mk = W - 1 - (6 * dQ + bk);
          if (P < Wb) {
            const int mi = W - 1 - (6 * dP + bi);
            a.D[(size_t)mi * W + mk] = T[i][k];
            if (P != Q) a.D[(size_t)mk * W + mi] = T[i][k];
          } else if (Q < Wb) { if (bi < 4) DA[(size_t)bi * W + mk] = T[i][k]; }
          else if (bi < 4 && bk < 4) DA[4 * (size_t)W + 4 * bi + bk] = T[i][k];
        }
    }
    if (act && blk == nblk - 1) {
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int k = 0; k < TC; ++k)
          if (sr + i < 4 && sc + k < 4) s_c4[4 * (sr + i) + sc + k] = T[i][k];
    }
  } else if (tid < NG) {
    // ------------------------------------------------------------ panel group: thread = row (position p, a)
    const int r = tid - NW;
    const bool pact = r < NPR;
    const int p = pact ? r / 6 : 0, ar = pact ? r % 6 : 0;
    double yo[6] = {0, 0, 0, 0, 0, 0};     // own row of the current panel (unnormalised)
    int kre = -1;                         // step at which this thread's position was recycled last (in this phase)
    // raw row ar of block row kr + Wb (ring slot kr & 3) against column block j: A[r][6 j + m], r = 6 (kr + Wb) + ar
    auto ring_row = [&](int kr, int j, double (&dst)[6]) {
      const double* rg = ring + (kr & 3) * RB + ar * RS + (6 * (kr + Wb) + ar - 6 * j);
#pragma unroll
      for (int m = 0; m < 6; ++m) dst[m] = rg[-m];
    };
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0, q_pbar = 0, q_chol = 0, q_out = 0;
    // col = row (p, ar) of the fully updated pivot column block at position cp: factor the diagonal block (LDL'
    // recurrence, every thread redundantly: one reciprocal per pivot on the chain), solve the row on the fly,
    // publish y and z = y / d
    auto finish = [&](int cp, double (&col)[6], double* Ydst, double* Zdst) {
      if (pact && p == cp) {
        double2* o = reinterpret_cast<double2*>(Dsh + 6 * ar);
#pragma unroll
        for (int m = 0; m < 6; m += 2) o[m >> 1] = make_double2(col[m], col[m + 1]);
      }
      long long tq = a.prof ? clock64() : 0;
      bc_pbar(NP);
      if (a.prof) { const long long t = clock64(); q_pbar += t - tq; tq = t; }
      double Dl[6][6], w[6][6], invd[6];
      {
        double dfull[36];
        const double2* dd = reinterpret_cast<const double2*>(Dsh);
#pragma unroll
        for (int t = 0; t < 18; ++t) { const double2 z = dd[t]; dfull[2 * t] = z.x; dfull[2 * t + 1] = z.y; }
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) Dl[i][j] = dfull[6 * i + j];
      }
      const bool diag = p == cp;
      double z[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double dj = Dl[j][j];
        bad |= !(dj > 0.0 && dj <= 1.7976931348623157e308);
        invd[j] = bc_rcp(dj);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) w[i][j] = Dl[i][j] * invd[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
          for (int i2 = j + 1; i2 <= i; ++i2) Dl[i][i2] = fma(-Dl[i][j], w[i2][j], Dl[i][i2]);
        // row solve, column j: y_j = col_j - sum_{n < j} y_n w[j][n]
        double t = col[j];
#pragma unroll
        for (int n = 0; n < j; ++n) t = fma(-yo[n], w[j][n], t);
        if (diag && j > ar) t = 0.0;       // upper part of the diagonal block
        yo[j] = t;
        z[j] = t * invd[j];
      }
      if (a.prof) { const long long t = clock64(); q_chol += t - tq + (z[5] == 1.25e-300 ? 1 : 0); tq = t; }
      if (pact) {
        double2* oy = reinterpret_cast<double2*>(Ydst + p * B6_PBS + 6 * ar);
        double2* oz = reinterpret_cast<double2*>(Zdst + p * B6_PBS + 6 * ar);
#pragma unroll
        for (int m = 0; m < 6; m += 2) { oy[m >> 1] = make_double2(yo[m], yo[m + 1]); oz[m >> 1] = make_double2(z[m], z[m + 1]); }
      }
      if (a.prof) q_out += clock64() - tq;
    };
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      const int kb = phase == 0 ? kb0 : jsw, ke = phase == 0 ? ke0 : nsteps;
      if (phase == 1) {
        if (jsw < 0) break;
        bc_gbar(NG2);
      }
      int c = kb % Wb;
      kre = -1;
      bc_gbar(NG2);
      {
        double col[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) col[m] = pact ? Raw0[p * B6_PBS + 6 * ar + m] : 0.0;
        finish(c, col, Ypan + (kb & 1) * PB, Zpan + (kb & 1) * PB);
      }
      bc_gbar(NG2);
#pragma unroll 1
      for (int k = kb; k < ke; ++k) {
        const bool last = k == ke - 1;
        const int cn = c + 1 == Wb ? 0 : c + 1;
        if (!last) {
          double col[6];
          const bool useA = pact && p == c && p < Wb;          // recycled in this step: row of block k + Wb, no update
          if (useA) { ring_row(k, k + 1, col); kre = k; }
          else if (pact && kre == k - 1 && kre >= 0) ring_row(k - 1, k + 1, col);   // recycled in the previous step
          else {
            const double2* rawp = reinterpret_cast<const double2*>(Raw + ((k - 1) & 1) * PB + p * B6_PBS + 6 * ar);
#pragma unroll
            for (int m = 0; m < 3; ++m) { const double2 v = rawp[m]; col[2 * m] = pact ? v.x : 0.0; col[2 * m + 1] = pact ? v.y : 0.0; }
          }
          if (!useA) {
            const double2* Zcn = reinterpret_cast<const double2*>(Zpan + (k & 1) * PB + cn * B6_PBS);
#pragma unroll
            for (int m = 0; m < 6; ++m) {
              double t = col[m];
#pragma unroll
              for (int n = 0; n < 3; ++n) { const double2 v = Zcn[3 * m + n]; t = fma(-yo[2 * n], v.x, t); t = fma(-yo[2 * n + 1], v.y, t); }
              col[m] = t;
            }
          }
          finish(cn, col, Ypan + ((k + 1) & 1) * PB, Zpan + ((k + 1) & 1) * PB);
        }
        if (a.prof) { const long long t = clock64(); p_own += t - tlast; tlast = t; }
        bc_gbar(NG2);
        if (a.prof) { const long long t = clock64(); p_wait += t - tlast; tlast = t; }
        c = cn;
      }
    }
    if (a.prof && tid == NW && side == 0) { a.prof[4] = p_own; a.prof[5] = p_wait; a.prof[11] = q_pbar; a.prof[12] = q_chol; a.prof[13] = q_out; }
  } else {
    // ------------------------------------------------------------ loader warps: block row k + Wb of Ab -> ring slot k & 3,
    //      one step before it is used, loaded from global memory one step before that (registers in between)
    const int lt = tid - NG, nl = (int)blockDim.x - NG;
    const int maxblk = nsteps - 1 + Wb;                  // last block row anybody reads
    constexpr int LPT = 12;                              // double2 per loader thread: 32 loaders x 12 x 2 = 768 >= 6 RS up to Wb = 20, 64 loaders for all
    double2 rg[LPT];
    auto ld = [&](int blk) {
      const double2* src = reinterpret_cast<const double2*>(Ab + (size_t)6 * blk * RS);
#pragma unroll
      for (int u = 0; u < LPT; ++u) {
        const int e = lt + u * nl;
        rg[u] = (blk <= maxblk && 2 * e < RB) ? __ldg(src + e) : make_double2(0.0, 0.0);
      }
    };
    auto st_ = [&](int slot) {
      double2* dst = reinterpret_cast<double2*>(ring + slot * RB);
#pragma unroll
      for (int u = 0; u < LPT; ++u) {
        const int e = lt + u * nl;
        if (2 * e < RB) dst[e] = rg[u];
      }
    };
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      const int kb = phase == 0 ? kb0 : jsw, ke = phase == 0 ? ke0 : nsteps;
      if (phase == 1) {
        if (jsw < 0) break;
        bc_gbar(NG2);
      }
      ld(kb + Wb); st_(kb & 3);
      ld(kb + 1 + Wb);
      bc_gbar(NG2);
      bc_gbar(NG2);
#pragma unroll 1
      for (int k = kb; k < ke; ++k) {
        st_((k + 1) & 3);
        ld(k + 2 + Wb);
        bc_gbar(NG2);
      }
    }
  }
  if (a.two && side == 1) __threadfence();
  __syncthreads();
  const long long tk1 = a.prof ? clock64() : 0;
  if (bad) s_fail = 1;
  if (a.two && side == 1 && tid == 0) bc_post(a.sync, a.epoch);
  __syncthreads();
  bc_tail<2>(a, sd, side, bc_smem, s_fail, s_xI, s_c4, tk0, tk1);
}

// threads of the kernel for window W with BS x BS blocks
inline int band_chol_threads(int W, int BS) {
  const int Wb = W / BS, NT = (Wb + 1) * (Wb + 2) / 2;
  return ((NT + 31) & ~31) + 32 * ((W + 4 + 31) / 32);
}

inline size_t band_chol_smem(int bw) {
  const int ms = 1 + (bw + 31) / 32, msv = ms <= 4 ? 4 : BC_MAXSLOT;
  return sizeof(double) * ((size_t)(2 + BC_RING) * BC_CSM + 2 * 32 * (size_t)(32 * (msv + 1)));
}

// window for half bandwidth bw: the next multiple of 8 above bw (0 when it exceeds the register window)
inline int band_chol_window(int bw) {
  const int W = ((bw + 1 + 7) / 8) * 8;
  return W <= BC_MAXW ? W : 0;
}

// rows of Ab the kernel may touch (padding rows below the band part are identity rows)
inline int band_chol_rows(int nb, int W) { return ((nb + 7) & ~7) + W + 8; }

// How the pivot chain is cut.  Two-sided from 4 windows on (below that the hand-over costs more than
// the shorter chain saves); PSFM_CHOL_ONE_SIDED forces the one-CTA form.
// blk_span >= 0: the matrix is made of 6 x 6 blocks and blocks further apart than blk_span are zero (the reduced
// camera system: blk_span = longest image span of a track); -1: only the scalar half bandwidth bw is known.
inline BandPlan band_chol_plan(int nb, int bw, int blk_span = -1) {
  BandPlan p{};
  p.nb = nb; p.bw = std::min(bw, nb - 1);
  static const bool one = getenv("PSFM_CHOL_ONE_SIDED") != nullptr;
  static const bool rank1 = getenv("PSFM_CHOL_RANK1") != nullptr;
  // window of the block-6 kernel, in blocks: every block at distance >= Wb from the pivot block must be zero
  int Wb6 = blk_span >= 0 ? blk_span + 1 : (p.bw + 5) / 6 + 1;
  if (nb % 6 == 0) Wb6 = std::max(3, std::min(Wb6, nb / 6));
  if (!rank1 && nb % 6 == 0 && nb / 6 >= 3 && Wb6 <= 25 && 6 * Wb6 > p.bw) {
    // block-6 form: W = 6 Wb, everything in units of image blocks
    const int F = nb / 6, Wb = Wb6;
    p.blk6 = 1;
    p.W = 6 * Wb; p.RS = p.W + 4; p.nbp = nb;
    p.two = (!one && F >= 4 * Wb) ? 1 : 0;
    if (p.two) {
      p.k0 = 6 * ((F - Wb) / 2);
      p.n1 = p.nbp - p.W - p.k0;
      p.nbs[0] = p.k0 + p.W; p.npiv[0] = p.k0 + p.W;
      p.nbs[1] = p.n1 + p.W; p.npiv[1] = p.n1;
      p.rows[0] = band_chol_rows(p.nbs[0], p.W); p.rows[1] = band_chol_rows(p.nbs[1], p.W);
    } else {
      p.nbs[0] = nb; p.npiv[0] = nb; p.rows[0] = band_chol_rows(nb, p.W);
    }
    return p;
  }
  p.W = band_chol_window(p.bw); p.RS = p.W + 4;
  p.nbp = (nb + 7) & ~7;
  p.two = (p.W > 0 && !one && p.nbp >= 4 * p.W) ? 1 : 0;
  if (p.two) {
    p.k0 = (((p.nbp - p.W) / 2 + 7) / 8) * 8;
    p.n1 = p.nbp - p.W - p.k0;
    p.nbs[0] = p.k0 + p.W; p.npiv[0] = p.k0 + p.W;
    p.nbs[1] = p.n1 + p.W; p.npiv[1] = p.n1;
    p.rows[0] = band_chol_rows(p.nbs[0], p.W); p.rows[1] = band_chol_rows(p.nbs[1], p.W);
  } else {
    p.nbs[0] = nb; p.npiv[0] = nb; p.rows[0] = p.W ? band_chol_rows(nb, p.W) : 0;
  }
  return p;
}

// device buffers of one plan
struct BandWork {
  BandPlan pl{};
  DBuf<double> Ab, C4, Lr, La, dinv, D;
  DBuf<int> sync;
  int epoch = 0;
  void alloc(const BandPlan& p, cudaStream_t st) {
    pl = p;
    const size_t LS = p.bw + 1;
    Ab.alloc((size_t)(p.rows[0] + p.rows[1]) * p.RS, st);
    C4.alloc(32, st);
    Lr.alloc((size_t)(p.nbs[0] + p.nbs[1]) * LS, st); Lr.zero(st);
    La.alloc(4 * (size_t)(p.nbs[0] + p.nbs[1]), st);
    dinv.alloc((size_t)p.nbs[0] + p.nbs[1], st);
    D.alloc((size_t)p.W * p.W + 4 * (size_t)p.W + 16, st);
    sync.alloc(2, st); sync.zero(st);
    epoch = 0;
  }
  double* ab(int side) { return Ab.p + (side ? (size_t)pl.rows[0] * pl.RS : 0); }
  BandCholArgs args(double* x, int ns, int* fail) {
    BandCholArgs c{};
    const size_t LS = pl.bw + 1;
    for (int s = 0; s < 2; ++s) {
      const size_t o = s ? (size_t)pl.nbs[0] : 0;
      c.s[s].Ab = ab(s); c.s[s].C4 = C4.p + 16 * s; c.s[s].nb = pl.nbs[s]; c.s[s].npiv = pl.npiv[s];
      c.s[s].Lr = Lr.p + o * LS; c.s[s].La = La.p + 4 * o; c.s[s].dinv = dinv.p + o;
    }
    c.two = pl.two; c.nbg = pl.nb; c.nbp = pl.nbp; c.k0 = pl.k0;
    c.bw = pl.bw; c.W = pl.W; c.RS = pl.RS; c.ns = ns; c.blk6 = pl.blk6;
    c.x = x; c.fail = fail; c.D = D.p; c.sync = sync.p; c.epoch = ++epoch;
    c.prof = nullptr;
    return c;
  }
};

// one launch: one CTA, or two for the two-sided form
inline void band_chol_launch(BandCholArgs c, cudaStream_t st) {
  static const int flags = getenv("PSFM_CHOL_FLAGS") ? atoi(getenv("PSFM_CHOL_FLAGS")) : 0;
  c.flags = flags;
  const int grid = c.two ? 2 : 1;
  if (c.blk6) {
    const int Wb = c.W / 6, nblk = (Wb + 1) * (Wb + 2) / 2;
    const int np = (6 * (Wb + 1) + 31) & ~31;
    const int ng4 = ((4 * nblk + 31) & ~31) + np, ng1 = ((nblk + 31) & ~31) + np;
    const size_t fact = sizeof(double) * (7 * (size_t)(Wb + 1) * B6_PBS + 36 + 4 * 6 * (size_t)c.RS);
    const size_t smem = std::max(fact, band_chol_smem(c.bw));
    static const bool fat = getenv("PSFM_CHOL_FAT") != nullptr;     // measurement: one thread per block everywhere
#define PSFM_BC6_GO(MT, TRV, TCV)                                                                                      \
  do {                                                                                                             \
    static size_t attr = 0;                                                                                        \
    if (smem > attr) {                                                                                             \
      PSFM_CUDA(cudaFuncSetAttribute(k_band_chol6<MT, TRV, TCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = smem;                                                                                                 \
    }                                                                                                              \
    k_band_chol6<MT, TRV, TCV><<<grid, MT, smem, st>>>(c);                                                               \
  } while (0)
    // thin threads while the CTA has room for them (measured, profiles/r02_band_chol_notes.md), else one fat thread
    // per block; the threads after the workers and the panel group are the loader (>= 32)
    static const int tile = getenv("PSFM_CHOL_TILE") ? atoi(getenv("PSFM_CHOL_TILE")) : 0;     // measurement: 33 | 36 | 66
    const int ng2 = ((2 * nblk + 31) & ~31) + np;
    if (!fat && tile != 36 && tile != 66 && ng4 + 32 <= 512) PSFM_BC6_GO(512, 3, 3);
    else if (!fat && tile != 66 && ng2 + 32 <= 384) PSFM_BC6_GO(384, 3, 6);
    else if (ng1 + 64 <= 256) PSFM_BC6_GO(256, 6, 6);
    else if (ng1 + 64 <= 512) PSFM_BC6_GO(512, 6, 6);
    else PSFM_BC6_GO(640, 6, 6);
#undef PSFM_BC6_GO
    PSFM_LAUNCH_CHECK();
    return;
  }
  const int threads = band_chol_threads(c.W, 4);
  const size_t smem = band_chol_smem(c.bw);
#define PSFM_BC_GO(BSV, MT)                                                                                        \
  do {                                                                                                             \
    static size_t attr = 0;                                                                                        \
    if (smem > attr) {                                                                                             \
      PSFM_CUDA(cudaFuncSetAttribute(k_band_chol<BSV, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = smem;                                                                                                 \
    }                                                                                                              \
    k_band_chol<BSV, MT><<<grid, threads, smem, st>>>(c);                                                           \
  } while (0)
  if (threads <= 512) PSFM_BC_GO(4, 512);
  else PSFM_BC_GO(4, 1024);
#undef PSFM_BC_GO
  PSFM_LAUNCH_CHECK();
}

}  // namespace ba
}  // namespace psfm
