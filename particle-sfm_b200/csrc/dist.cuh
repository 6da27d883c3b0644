// dist.cuh — multi-GPU plumbing of HP2: one process per GPU, points sharded, the
// camera-side vectors replicated and summed with one NCCL all-reduce.
// NCCL is dlopen()ed lazily so that the library loads on machines without it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace psfm {
namespace dist {
int world_size();
int rank();
// in-place sum / max over ranks on `stream`; no-ops when world_size() == 1
void allreduce_sum(double* buf, size_t n, cudaStream_t stream);
void allreduce_max(double* buf, size_t n, cudaStream_t stream);
}  // namespace dist
}  // namespace psfm
