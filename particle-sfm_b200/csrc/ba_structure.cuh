// ba_structure.cuh — device-side flattening of a bundle-adjustment problem into the tiled
// layout of ba_kernels.cuh (what BundleAdjuster::SetUp enumerates on the host in the
// reference, bundle_adjustment.cc:326-447, done here with radix sorts and scans on the GPU
// because it is inside the end-to-end time of psfm_ba_solve).
//   observed points ordered by (first image, id)  ->  observations sorted by (point, image)
//   ->  tiles of whole points  ->  per tile: image order, image segments, local indices.
// CUB (shipped with the CUDA toolkit) provides the sorts/scans; this is set-up plumbing, not
// a hot kernel.
#pragma once
#include <cub/cub.cuh>

#include "psfm_common.cuh"

namespace psfm {
namespace ba {

__global__ void k_st_init(int* cnt, int* min_img, int Pt, int F) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < Pt) { cnt[p] = 0; min_img[p] = F; }
}

__global__ void k_st_count(const int* img, const int* pt, int M, int F, int Pt, int* cnt, int* min_img,
                           unsigned char* img_has_obs, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int im = img[i], p = pt[i];
  if (im < 0 || im >= F || p < 0 || p >= Pt) { *bad = 1; return; }
  atomicAdd(cnt + p, 1);
  atomicMin(min_img + p, im);
  img_has_obs[im] = 1;
}

__global__ void k_st_iota(int* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// order[j] = caller's id of internal point j ; writes pt_new and the counts in internal order
__global__ void k_st_rank(const int* order, const int* cnt, int Pt, int* pt_new, int* cnt_sorted) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Pt) return;
  const int p = order[j];
  const int c = cnt[p];
  pt_new[p] = c > 0 ? j : -1;      // unobserved points sort last (key = F), so j < P iff observed
  cnt_sorted[j] = c;
}

__global__ void k_st_keys(const int* img, const int* pt, const int* pt_new, int M, unsigned long long* keys, int* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  keys[i] = ((unsigned long long)(unsigned)pt_new[pt[i]] << 32) | (unsigned)img[i];
  vals[i] = i;
}

__global__ void k_st_gather(const unsigned long long* keys, const int* idx, const double2* xy, int M, int* obs_img,
                            int* obs_pt, double2* obs_xy) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  obs_img[j] = (int)(keys[j] & 0xffffffffull);
  obs_pt[j] = (int)(keys[j] >> 32);
  obs_xy[j] = xy[idx[j]];
}

// One CTA per tile: rank every observation by (image, position) -> image order; local
// segment / point indices; number of image segments of the tile.
template <int TILE>
__global__ void __launch_bounds__(TILE) k_st_tile_order(const int* tile_start, const int* tile_pt, const int* obs_img,
                                                        const int* obs_pt, unsigned short* tile_perm,
                                                        unsigned short* obs_lseg, unsigned short* obs_lpt, int* tile_ns) {
  __shared__ int s_img[TILE];
  __shared__ int s_sorted[TILE];
  __shared__ int s_cnt;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int base = tile_start[tile], n = tile_start[tile + 1] - base, pt0 = tile_pt[tile];
  if (tid == 0) s_cnt = 0;
  int im = 0x7fffffff;
  if (tid < n) im = obs_img[base + tid];
  s_img[tid] = im;
  __syncthreads();
  int rank = 0;
  if (tid < n) {
    for (int e = 0; e < n; ++e) {
      const int o = s_img[e];
      rank += (o < im) || (o == im && e < tid);
    }
    tile_perm[base + rank] = (unsigned short)tid;
    s_sorted[rank] = im;
    obs_lpt[base + tid] = (unsigned short)(obs_pt[base + tid] - pt0);
  }
  __syncthreads();
  if (tid < n) {
    // segment index = number of distinct images smaller than mine
    int seg = 0;
    for (int e = 1; e < n; ++e) {
      const int o = s_sorted[e];
      if (o > im) break;
      seg += (o != s_sorted[e - 1]);
    }
    obs_lseg[base + tid] = (unsigned short)seg;
    if (tid == 0 || s_sorted[tid] != s_sorted[tid - 1]) atomicAdd(&s_cnt, 1);
  }
  __syncthreads();
  if (tid == 0) tile_ns[tile] = s_cnt;
}

template <int TILE>
__global__ void __launch_bounds__(TILE) k_st_tile_segments(const int* tile_start, const int* obs_img,
                                                           const unsigned short* tile_perm, const unsigned short* obs_lseg,
                                                           const int* cseg_ptr, int* cseg_img, unsigned short* cseg_off) {
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int base = tile_start[tile], n = tile_start[tile + 1] - base;
  if (tid >= n) return;
  const int e = tile_perm[base + tid];                  // tid-th observation in image order
  const int im = obs_img[base + e];
  const int prev = tid > 0 ? obs_img[base + tile_perm[base + tid - 1]] : -1;
  if (im != prev) {
    const int s = cseg_ptr[tile] + obs_lseg[base + e];
    cseg_img[s] = im;
    cseg_off[s] = (unsigned short)tid;
  }
}

}  // namespace ba
}  // namespace psfm
