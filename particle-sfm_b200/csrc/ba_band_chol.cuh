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
// tools/emu_band_chol.py is a thread-level numpy emulation of exactly this index logic.
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
  int F, span, nb, bw, W, RS, nrows;
  double* Ab;               // [nrows][RS]: Ab[r][k] = A[r][r - k] (k <= bw), Ab[r][W + a] = A[nb + a][r], a < 3; Ab[r][W + 3] = rhs[r]
  double* C4;               // [4][4] arrow corner: intrinsics block (3 x 3) | rhs entries in row / column 3
};

// packed index of element (r, c) of a symmetric 3 x 3 stored as 00 01 02 11 12 22
__device__ __forceinline__ int sym3(int r, int c) {
  const int lo = min(r, c), hi = max(r, c);
  return lo * (5 - lo) / 2 + hi;
}

__global__ void k_band_assemble(const BandAsmArgs2 a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 16) {   // corner (camera 0: intrinsics are only ever free for a single shared camera)
    const int r = (int)t / 4, c = (int)t % 4;
    const size_t sk = 6 * (size_t)a.F;
    double v = 0.0;
    if (r < 3 && c < 3) {
      v = a.scale_c[sk + r] * a.scale_c[sk + c] * (a.lin_intr[sym3(r, c)] + a.prep_intr[sym3(r, c)]);
      if (r == c) { if (a.active[sk + r]) v += a.Dc2[sk + r]; else v = 1.0; }
    } else if (r == 3 && c < 3) v = a.rhs[sk + c];
    else if (c == 3 && r < 3) v = a.rhs[sk + r];
    a.C4[t] = v;
  }
  if (t >= (size_t)a.nrows * a.RS) return;
  const int r = (int)(t / a.RS), e = (int)(t % a.RS);
  double v = 0.0;
  if (r >= a.nb) {
    v = (e == 0) ? 1.0 : 0.0;                       // identity padding below the band part
  } else if (e < a.W) {
    const int c = r - e;
    if (e <= a.bw && c >= 0) {
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
    }
  } else {
    const int aa = e - a.W;
    if (aa == 3) v = a.rhs[r];
    else if (aa == 0) {
      const double* X = a.xcam + (size_t)(r / 6) * a.xstride;
      v = a.scale_c[r] * a.scale_c[6 * (size_t)a.F] * (X[9 + r % 6] + X[15 + r % 6]);
    }
  }
  a.Ab[t] = v;
}

struct BandCholArgs {
  const double* Ab;
  const double* C4;
  int nb, bw, W, RS, ns;    // ns: length of x (slots past nb + 3 — other cameras — are zeroed)
  double* Lr;               // [nb][bw + 1]  UNNORMALISED columns: Lr[r][k] = A~[r][r - k] (= L[r][r - k] sqrt(d[r - k]))
  double* La;               // [4][nb]       unnormalised arrow rows (row 3 = rhs)
  double* dinv;             // [nb]          pivots d[j], replaced by 1 / sqrt(d[j]) after the factorisation
  double* x;                // [ns]
  int* fail;
  int flags;                // timing experiments (PSFM_CHOL_FLAGS; results invalid): 1 no output of L, 2 no reciprocal, 4 no publish, 8 barrier + pivot load only, 16 no row streaming
  long long* prof;          // optional [8]: SM cycles of factorisation | corner + staging | back substitution, pivots
};

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

// Roles: workers (one BS x BS block of window slots each), helper lanes (one element each:
// stream entry e of the upcoming rows into the ring with cp.async, 4 pivots ahead, and write
// entry e of the finished column to global memory).
template <int BS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_band_chol(const BandCholArgs a) {
  extern __shared__ __align__(16) double bc_smem[];
  __shared__ int s_fail;
  __shared__ double s_xI[4];
  constexpr int UN = 8;                     // pivots per unrolled body: ring slot and colbuf parity are compile-time
  const int W = a.W, Wb = W / BS, RS = a.RS, nb = a.nb, bw = a.bw, LS = bw + 1;
  double* colbuf = bc_smem;                 // [2][BC_CSM]   pivot column by window position (double buffered)
  double* ring = colbuf + 2 * BC_CSM;       // [BC_RING][BC_CSM] upcoming rows, permuted to window positions
  double* stage = ring + BC_RING * BC_CSM;  // [2][32][LSP] coefficients of the back substitution
  const int tid = threadIdx.x, lane = tid & 31;
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
          x = __ldg(a.Ab + (size_t)hi * RS + (hi - lo));
        } else if (Q < Wb) { if (i < 4) x = __ldg(a.Ab + (size_t)(BS * Q + k) * RS + W + i); }
        else if (i < 4 && k < 4) x = __ldg(a.C4 + 4 * i + k);
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
  const int nsteps = ((nb + UN - 1) / UN) * UN;
  bool bad = false;

  if (helper) {
    // ---- element e: entry e of row r of Ab (column r - e) goes to window position (r - e) mod W of ring
    //      slot r mod 8; entries W .. W+3 are the arrow and keep their position.  After the barrier of
    //      pivot j, position e of the pivot column is entry (e - pj) mod W of column j of L (unnormalised).
    const int e = tid - ldr0;
    const bool band = e < W, live = e < W + 4;
    const double* src = a.Ab + (size_t)W * RS + e;        // row W
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
    double* lp = a.Lr + (size_t)e * LS + e;
    double* la = a.La + (size_t)(live && !band ? e - W : 0) * nb;
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0;
    for (int j0 = 0; j0 < nsteps; j0 += UN) {
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
            if (kk == 0) a.dinv[j] = val;                  // the pivot itself
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
    if (a.prof && tid == ldr0) { a.prof[4] = p_own; a.prof[5] = p_wait; }
  } else {
    // ---- workers (and idle threads of the last worker warp: barriers only)
    const double* sP = colbuf + bc_idx(BS * P);
    const double* sQ = colbuf + bc_idx(BS * Q);
    int pj0 = 0, pjp = 0;                                  // pivot position of step u = 0 of the body, and its padded index
    int Pj = 0;
    long long p_own = 0, p_wait = 0, tlast = a.prof ? clock64() : 0;
    for (int j0 = 0; j0 < nsteps; j0 += UN) {
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
    if (a.prof && tid == 0) { a.prof[6] = p_own; a.prof[7] = p_wait; }
  }
  __syncthreads();
  const long long tk1 = a.prof ? clock64() : 0;
  if (a.prof && tid == 0) { a.prof[0] = tk1 - tk0; a.prof[3] = nsteps; }
  if (bad) s_fail = 1;
  // 1 / sqrt(d): normalisation of the stored columns, applied while staging the back substitution
  for (int j = tid; j < nb; j += blockDim.x) a.dinv[j] = rsqrt(__ldcg(a.dinv + j));
  // ---- arrow corner: 3 x 3 intrinsics block and its right-hand side (thread of block {Wb, Wb})
  if (tid == NT - 1) {
    bool cbad = false;
    const double m00 = v[0][0], m10 = v[1][0], m20 = v[2][0], m11 = v[1][1], m21 = v[2][1], m22 = v[2][2];
    cbad |= !(m00 > 0.0);
    const double l00 = sqrt(m00), l10 = m10 / l00, l20 = m20 / l00;
    double t = m11 - l10 * l10;
    cbad |= !(t > 0.0);
    const double l11 = sqrt(t), l21 = (m21 - l20 * l10) / l11;
    t = m22 - l20 * l20 - l21 * l21;
    cbad |= !(t > 0.0);
    const double l22 = sqrt(t);
    const double z0 = v[3][0] / l00, z1 = (v[3][1] - l10 * z0) / l11, z2 = (v[3][2] - l20 * z0 - l21 * z1) / l22;
    const double x2 = z2 / l22, x1 = (z1 - l21 * x2) / l11, x0 = (z0 - l10 * x1 - l20 * x2) / l00;
    s_xI[0] = x0; s_xI[1] = x1; s_xI[2] = x2;
    if (cbad || !isfinite(x0 + x1 + x2)) s_fail = 1;
  }
  __syncthreads();
  if (tid == 0) *a.fail = s_fail;
  if (s_fail) return;
  for (int s = nb + tid; s < a.ns; s += blockDim.x) a.x[s] = (s < nb + 3) ? s_xI[s - nb] : 0.0;

  // ---- back substitution L' x = y - La' x_I in axpy form (warp 0; the other warps stage).
  //      Lane l holds the running right-hand side of positions 32 (c - m) + l, m = 0 .. msv-1, of the
  //      current 32-column chunk c.  stage[jj][32 + k] = -L[r][r - k] for 1 <= k <= bw (r = 32 c + jj),
  //      zero elsewhere, so the inner step is one shared load and one fma per slot — no predicates;
  //      per pivot one multiply, one shuffle and one fma are on the dependent chain.
  const int ms = 1 + (bw + 31) / 32;
  const int msv = ms <= 4 ? 4 : BC_MAXSLOT;
  const int LSP = 32 * (msv + 1);
  const int ctop = (nb + 31) / 32 - 1;
  const int CH = 32 * LSP;
  auto stage_chunk = [&](int c, double* buf, int w0, int nw) {
    for (int jj = w0; jj < 32; jj += nw) {
      const int r = 32 * c + jj;
      for (int q = lane; q < LSP; q += 32) {
        const int k = q - 32;
        double val = 0.0;
        if (k >= 1 && k <= bw && r < nb && r - k >= 0) val = -__ldcg(a.Lr + (size_t)r * LS + k) * __ldcg(a.dinv + r - k);
        buf[jj * LSP + q] = val;
      }
    }
  };
  const double xi0 = s_xI[0], xi1 = s_xI[1], xi2 = s_xI[2];
  auto y0 = [&](int i) -> double {
    if (i < 0 || i >= nb) return 0.0;
    return __ldcg(a.dinv + i) * (__ldcg(a.La + 3 * (size_t)nb + i) -
           (xi0 * __ldcg(a.La + i) + xi1 * __ldcg(a.La + (size_t)nb + i) + xi2 * __ldcg(a.La + 2 * (size_t)nb + i)));
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
      const double dl = (jl < nb) ? __ldcg(a.dinv + jl) : 0.0;
      const double* rp = buf + 31 * LSP + 32 + (31 - lane);             // &stage[jj][32 + jj - lane], jj = 31
      const int step = LSP + 1;
      double* xo = a.x + 32 * c;
      if (msv == 4) {
#pragma unroll 8
        for (int jj = 31; jj >= 0; --jj) {
          const double xj = __shfl_sync(0xffffffffu, yy[0] * dl, jj);
#pragma unroll
          for (int m = 0; m < 4; ++m) yy[m] = fma(rp[32 * m], xj, yy[m]);
          if (lane == jj && jl < nb) xo[jj] = xj;
          rp -= step;
        }
      } else {
#pragma unroll 4
        for (int jj = 31; jj >= 0; --jj) {
          const double xj = __shfl_sync(0xffffffffu, yy[0] * dl, jj);
#pragma unroll
          for (int m = 0; m < BC_MAXSLOT; ++m) yy[m] = fma(rp[32 * m], xj, yy[m]);
          if (lane == jj && jl < nb) xo[jj] = xj;
          rp -= step;
        }
      }
#pragma unroll
      for (int m = 0; m < BC_MAXSLOT; ++m) yy[m] = (m == ms - 1) ? fresh : ((m + 1 < BC_MAXSLOT) ? yy[m + 1] : 0.0);
    }
    __syncthreads();
  }
  if (a.prof && tid == 0) {
    const long long tk3 = clock64();
    a.prof[0] = tk1 - tk0; a.prof[1] = tk2 - tk1; a.prof[2] = tk3 - tk2; a.prof[3] = nsteps;
  }
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

// one launch (4 x 4 register blocks)
inline void band_chol_launch(BandCholArgs c, cudaStream_t st) {
  static const int flags = getenv("PSFM_CHOL_FLAGS") ? atoi(getenv("PSFM_CHOL_FLAGS")) : 0;
  c.flags = flags;
  const int threads = band_chol_threads(c.W, 4);
  const size_t smem = band_chol_smem(c.bw);
#define PSFM_BC_GO(BSV, MT)                                                                                        \
  do {                                                                                                             \
    static size_t attr = 0;                                                                                        \
    if (smem > attr) {                                                                                             \
      PSFM_CUDA(cudaFuncSetAttribute(k_band_chol<BSV, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = smem;                                                                                                 \
    }                                                                                                              \
    k_band_chol<BSV, MT><<<1, threads, smem, st>>>(c);                                                              \
  } while (0)
  if (threads <= 512) PSFM_BC_GO(4, 512);
  else PSFM_BC_GO(4, 1024);
#undef PSFM_BC_GO
  PSFM_LAUNCH_CHECK();
}

}  // namespace ba
}  // namespace psfm
