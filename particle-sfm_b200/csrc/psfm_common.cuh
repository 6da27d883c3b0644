// psfm_common.cuh — shared helpers of the sm_100a library (error handling, launch
// accounting, block reductions).  Product code: no CPU path lives here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/psfm_b200.h"

namespace psfm {

void set_error(const std::string& msg);
extern std::atomic<long long> g_launch_count;

struct CudaFail {
  int code;
};

#define PSFM_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::psfm::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
      throw ::psfm::CudaFail{PSFM_ERR_CUDA};                                              \
    }                                                                                     \
  } while (0)

#define PSFM_LAUNCH_CHECK()                  \
  do {                                       \
    ::psfm::g_launch_count.fetch_add(1);     \
    PSFM_CUDA(cudaGetLastError());           \
  } while (0)

// RAII device buffer.  With a stream: stream-ordered allocation from the default memory
// pool (cudaMallocAsync), whose release threshold psfm::keep_pool_memory() raises so that
// create/destroy cycles of a solver re-use the cached blocks instead of paying cudaMalloc.
void keep_pool_memory();

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  bool async = false;
  DBuf() {}
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) {
      if (async) cudaFreeAsync(p, st);
      else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, cudaStream_t stream = nullptr) {
    release();
    n = count;
    if (count == 0) count = 1;
    if (stream) {
      keep_pool_memory();
      PSFM_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T) + 64, stream));   // slack: bulk copies read whole 16-byte windows
      st = stream;
      async = true;
    } else {
      PSFM_CUDA(cudaMalloc((void**)&p, count * sizeof(T) + 64));
      async = false;
    }
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count) PSFM_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) { PSFM_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), s)); }
};

#ifdef __CUDACC__
// ---- warp / block reductions (sum and max), result valid in thread 0 ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sbuf: at least 32 doubles of shared memory; all threads of the block must call.
__device__ __forceinline__ double block_sum(double v, double* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sbuf[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? sbuf[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}
__device__ __forceinline__ double block_max(double v, double* sbuf) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sbuf[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? sbuf[lane] : 0.0;
    v = warp_max(v);
  }
  return v;
}
// max of non-negative doubles through their bit pattern (monotone for x >= 0)
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
#endif

}  // namespace psfm
