// ba_tile_pipe.cuh — persistent, software-pipelined form of the tile kernels.
//
// The one-CTA-per-tile kernels of ba_kernels.cuh pay three dependent global-memory latencies
// per tile (tile header -> segment / point lists -> poses) before any arithmetic starts, and
// only 2-3 CTAs fit on an SM to hide them.  Here a CTA is PERSISTENT (grid = SMs x resident
// CTAs) and walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; every input of tile k+1 is
// copied global -> shared with cp.async (LDGSTS, no register staging) into the second stage
// of a two-stage shared-memory buffer while tile k is being computed, and the 32-byte header
// of tile k+2 rides along.  All inputs are single-level (contiguous ranges addressed from
// the header alone): the per-segment pose rows are gathered once per linearisation into
// seg_pose by k_seg_pose, the point ranges / segment offsets are stored tile-relative.
#pragma once
#include "ba_kernels.cuh"

namespace psfm {
namespace ba {

__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Global arrays in pipeline form (built once per problem, seg_pose once per linearisation)
struct PipeSrc {
  const int4* tile_hdr;             // [T][2]  {base, n, pt0, np} {cs0, ns, 0, 0}
  const unsigned short* obs_lseg;   // [M + 2]
  const unsigned short* obs_lpt;    // [M + 2]
  const unsigned short* tile_perm;  // [M + 2]
  const int* pstart_rel;            // [P]     first observation of the point, relative to its tile
  const int* cseg_img;              // [nseg]
  const int* cseg_off32;            // [nseg]  start of the segment in the tile's image order
  const double* seg_pose;           // [nseg][12]  R (9) | t (3) of the segment's image
  // what the kernel wants staged besides the structure
  const double2* obs_xy;            // [M] or null
  const double* obs_a;              // [3][M] or null
  const double* X;                  // [3P]
  const double* p6;                 // [6][P] or null   -> spt rows 3..8
  const double* p3a;                // [3][P] or null   -> spt rows 9..11
  const double* p3b;                // [3][P] or null   -> spt rows 12..14
};

// gather pose16 rows per tile segment (after k_pose_table)
__global__ void k_seg_pose(const int* __restrict__ cseg_img, const double* __restrict__ pose16, int nseg,
                           double* __restrict__ seg_pose) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nseg * 12) return;
  const int s = (int)(t / 12), k = (int)(t % 12);
  seg_pose[t] = pose16[16 * (size_t)cseg_img[s] + k];
}

// tile-relative structure arrays
__global__ void k_pipe_headers(const int* tile_start, const int* tile_pt, const int* cseg_ptr, int T, int4* hdr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  hdr[2 * t] = make_int4(tile_start[t], tile_start[t + 1] - tile_start[t], tile_pt[t], tile_pt[t + 1] - tile_pt[t]);
  hdr[2 * t + 1] = make_int4(cseg_ptr[t], cseg_ptr[t + 1] - cseg_ptr[t], 0, 0);
}
__global__ void k_pipe_pstart(const int* tile_start, const int* tile_pt, const int* pt_ptr, int T, int* pstart_rel) {
  const int t = blockIdx.x;
  const int base = tile_start[t], p0 = tile_pt[t], p1 = tile_pt[t + 1];
  for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) pstart_rel[p] = pt_ptr[p] - base;
}
__global__ void k_pipe_off32(const unsigned short* cseg_off, int nseg, int* off32) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) off32[s] = cseg_off[s];
}

// One shared-memory stage: everything a tile reads that comes from global memory.
template <int TILE>
struct PipeStage {
  double2* xy;            // [TILE]
  double* a;              // [3][TILE]
  double* simg;           // [12][cap_ns]
  double* spt;            // [npt][cap_np]
  int *pstart, *coff, *cimg;
  unsigned short *lseg, *lpt, *perm;   // [TILE + 2] each (4-byte copies from an even index)
  static __host__ __device__ size_t bytes(bool has_xy, bool has_a, int npt, int cns, int cnp) {
    size_t b = 0;
    if (has_xy) b += sizeof(double2) * TILE;
    if (has_a) b += sizeof(double) * 3 * TILE;
    b += sizeof(double) * (12 * (size_t)cns + (size_t)npt * cnp);
    b += sizeof(int) * ((size_t)cnp + 1 + 2 * ((size_t)cns + 1) + 1);
    b += sizeof(unsigned short) * 3 * (TILE + 4);
    return (b + 15) & ~(size_t)15;
  }
  __device__ __forceinline__ void carve(unsigned char* base, bool has_xy, bool has_a, int npt, int cns, int cnp) {
    unsigned char* p = base;
    xy = reinterpret_cast<double2*>(p); if (has_xy) p += sizeof(double2) * TILE;
    a = reinterpret_cast<double*>(p); if (has_a) p += sizeof(double) * 3 * TILE;
    simg = reinterpret_cast<double*>(p); p += sizeof(double) * 12 * (size_t)cns;
    spt = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)npt * cnp;
    pstart = reinterpret_cast<int*>(p); p += sizeof(int) * ((size_t)cnp + 1);
    coff = reinterpret_cast<int*>(p); p += sizeof(int) * ((size_t)cns + 1);
    cimg = reinterpret_cast<int*>(p); p += sizeof(int) * ((size_t)cns + 1);
    p += sizeof(int);                                   // keep the 2-byte arrays 4-byte aligned whatever the parity above
    p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 3) & ~(uintptr_t)3);
    lseg = reinterpret_cast<unsigned short*>(p); p += sizeof(unsigned short) * (TILE + 4);
    lpt = reinterpret_cast<unsigned short*>(p); p += sizeof(unsigned short) * (TILE + 4);
    perm = reinterpret_cast<unsigned short*>(p);
  }
};

// issue the asynchronous copies of one tile (all threads take part; no wait, no commit)
template <int TILE>
__device__ __forceinline__ void pipe_issue(const TileCtx& tc, const PipeSrc& ps, const PipeStage<TILE>& st, const int4 h0,
                                           const int4 h1, int cns, int cnp) {
  const int tid = threadIdx.x;
  const int base = h0.x, n = h0.y, pt0 = h0.z, np = h0.w, cs0 = h1.x, ns = h1.y;
  if (ps.obs_xy) {
    if (tid < n) cp_async16(st.xy + tid, ps.obs_xy + (size_t)base + tid);
  }
  if (ps.obs_a) {
    if (tid < n) {
#pragma unroll
      for (int k = 0; k < 3; ++k) cp_async8(st.a + k * TILE + tid, ps.obs_a + (size_t)k * tc.M + base + tid);
    }
  }
  {
    const int off = base & 1, words = (off + n + 1) >> 1;
    const size_t w0 = ((size_t)base - off) >> 1;
    if (tid < words) {
      cp_async4(reinterpret_cast<unsigned int*>(st.lseg) + tid, reinterpret_cast<const unsigned int*>(ps.obs_lseg) + w0 + tid);
      cp_async4(reinterpret_cast<unsigned int*>(st.lpt) + tid, reinterpret_cast<const unsigned int*>(ps.obs_lpt) + w0 + tid);
      cp_async4(reinterpret_cast<unsigned int*>(st.perm) + tid, reinterpret_cast<const unsigned int*>(ps.tile_perm) + w0 + tid);
    }
  }
  for (int j = tid; j < np; j += TILE) cp_async4(st.pstart + j, ps.pstart_rel + pt0 + j);
  for (int j = tid; j < ns; j += TILE) {
    cp_async4(st.coff + j, ps.cseg_off32 + cs0 + j);
    cp_async4(st.cimg + j, ps.cseg_img + cs0 + j);
  }
  for (int j = tid; j < ns * 12; j += TILE) {
    const int s = j / 12, k = j - 12 * s;
    cp_async8(st.simg + k * cns + s, ps.seg_pose + 12 * (size_t)cs0 + j);
  }
  for (int j = tid; j < np * 3; j += TILE) {
    const int l = j / 3, k = j - 3 * l;
    cp_async8(st.spt + k * cnp + l, ps.X + 3 * (size_t)pt0 + j);
  }
  if (ps.p6) {
    for (int j = tid; j < np * 6; j += TILE) {
      const int k = j / np, l = j - k * np;
      cp_async8(st.spt + (3 + k) * cnp + l, ps.p6 + (size_t)k * tc.P + pt0 + l);
    }
  }
  if (ps.p3a) {
    for (int j = tid; j < np * 3; j += TILE) {
      const int k = j / np, l = j - k * np;
      cp_async8(st.spt + (9 + k) * cnp + l, ps.p3a + (size_t)k * tc.P + pt0 + l);
    }
  }
  if (ps.p3b) {
    for (int j = tid; j < np * 3; j += TILE) {
      const int k = j / np, l = j - k * np;
      cp_async8(st.spt + (12 + k) * cnp + l, ps.p3b + (size_t)k * tc.P + pt0 + l);
    }
  }
}

// Drives the pipeline.  body(ti, stage, tile) computes one tile from a ready stage; it may
// use barriers and must be called by every thread.  sm_common: shared memory that is not
// staged (reduction scratch), carved by the caller after the two stages.
template <int TILE, typename Body>
__device__ __forceinline__ void pipe_run(const TileCtx& tc, const PipeSrc& ps, unsigned char* stage_base, int4 (*hdr_ring)[2],
                                         const int npt, const int cns, const int cnp, Body body) {
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  if (tile >= tc.T) return;
  const bool has_xy = ps.obs_xy != nullptr, has_a = ps.obs_a != nullptr;
  const size_t sb = PipeStage<TILE>::bytes(has_xy, has_a, npt, cns, cnp);
  PipeStage<TILE> st;
  st.carve(stage_base, has_xy, has_a, npt, cns, cnp);
  int4 h0 = __ldg(ps.tile_hdr + 2 * (size_t)tile), h1 = __ldg(ps.tile_hdr + 2 * (size_t)tile + 1);
  pipe_issue<TILE>(tc, ps, st, h0, h1, cns, cnp);
  {
    const int next = tile + gridDim.x;
    if (next < tc.T && tid < 2) cp_async16(&hdr_ring[1][tid], ps.tile_hdr + 2 * (size_t)next + tid);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (int it = 0;; ++it) {
    const int cur = it & 1;
    const int ntile = tile + gridDim.x, nntile = ntile + gridDim.x;
    int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
    if (ntile < tc.T) {
      n0 = hdr_ring[(it + 1) & 3][0];
      n1 = hdr_ring[(it + 1) & 3][1];
      st.carve(stage_base + (cur ^ 1) * sb, has_xy, has_a, npt, cns, cnp);
      pipe_issue<TILE>(tc, ps, st, n0, n1, cns, cnp);
      if (nntile < tc.T && tid < 2) cp_async16(&hdr_ring[(it + 2) & 3][tid], ps.tile_hdr + 2 * (size_t)nntile + tid);
    }
    cp_async_commit();
    TileInfo ti;
    ti.base = h0.x; ti.n = h0.y; ti.pt0 = h0.z; ti.np = h0.w; ti.cs0 = h1.x; ti.ns = h1.y;
    st.carve(stage_base + cur * sb, has_xy, has_a, npt, cns, cnp);
    if (tid == 0) { st.pstart[ti.np] = ti.n; st.coff[ti.ns] = ti.n; }
    body(ti, st, tile);
    cp_async_wait_all();
    __syncthreads();
    if (ntile >= tc.T) break;
    tile = ntile; h0 = n0; h1 = n1;
  }
}

// point a TileSmem view at a stage
template <int TILE>
__device__ __forceinline__ void view_stage(TileSmem<TILE>& sm, const PipeStage<TILE>& st, int base) {
  sm.simg = st.simg; sm.spt = st.spt; sm.pstart = st.pstart; sm.coff = st.coff; sm.cimg = st.cimg;
  sm.perm = st.perm + (base & 1);
}

// ------------------------------------------------------------------ pipelined Jacobian sweep

template <int TILE, bool ROT>
__global__ void __launch_bounds__(TILE, (TILE == 256 ? 3 : 1)) k_linearize_p(const TileCtx tc, const PipeSrc ps, const LinArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int4 hdr_ring[4][2];
  const int cns = tc.cap_ns, cnp = tc.cap_np;
  const size_t sb = PipeStage<TILE>::bytes(true, false, 3, cns, cnp);
  TileSmem<TILE> sm;
  sm.cap_ns = cns; sm.cap_np = cnp;
  sm.sv = reinterpret_cast<double*>(smem_raw + 2 * sb);
  sm.sw = nullptr;
  sm.sred = sm.sv + 18 * PSFM_SVS;
  sm.sx = nullptr;
  pipe_run<TILE>(tc, ps, smem_raw, hdr_ring, 3, cns, cnp, [&](const TileInfo& ti, const PipeStage<TILE>& s, int tile) {
    view_stage<TILE>(sm, s, ti.base);
    const int tid = threadIdx.x;
    const bool act = tid < ti.n;
    const int off = ti.base & 1;
    int ls = 0, lp = 0;
    double2 xy = make_double2(0.0, 0.0);
    if (act) { ls = s.lseg[off + tid]; lp = s.lpt[off + tid]; xy = s.xy[tid]; }
    linearize_tile<TILE, ROT>(tc, a, sm, ti, act, ls, lp, xy, tile);
  });
}

template <int TILE>
inline size_t pipe_smem_linearize(int cns, int cnp) {
  return 2 * PipeStage<TILE>::bytes(true, false, 3, cns, cnp) + sizeof(double) * (18 * (TILE + 1) + 9 * 32);
}

}  // namespace ba
}  // namespace psfm
