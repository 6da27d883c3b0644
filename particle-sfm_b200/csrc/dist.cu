// dist.cu — see dist.cuh.  The reference has no multi-GPU code at all;
// this is new design: psfm_dist_get_unique_id / psfm_dist_init build a private NCCL
// communicator (the id bytes travel through the caller's torch.distributed group).
#include "dist.cuh"

#include <dlfcn.h>

#include "psfm_common.cuh"

namespace psfm {
namespace dist {

// minimal NCCL ABI (nccl.h 2.27): opaque comm, 128-byte id, enums
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
enum { ncclFloat64 = 8 };

static struct {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
} g;

static bool load_nccl() {
  if (g.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g.handle) break;
  }
  if (!g.handle) {
    set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    return false;
  }
  g.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(g.handle, "ncclGetUniqueId");
  g.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(g.handle, "ncclCommInitRank");
  g.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g.handle, "ncclAllReduce");
  g.CommDestroy = (int (*)(ncclComm_t))dlsym(g.handle, "ncclCommDestroy");
  g.GetErrorString = (const char* (*)(int))dlsym(g.handle, "ncclGetErrorString");
  if (!g.GetUniqueId || !g.CommInitRank || !g.AllReduce || !g.CommDestroy) {
    set_error("libnccl.so.2 lacks a required symbol");
    return false;
  }
  return true;
}

int world_size() { return g.world; }
int rank() { return g.rank; }

static void check(int rc, const char* what) {
  if (rc != ncclSuccess) {
    set_error(std::string(what) + ": " + (g.GetErrorString ? g.GetErrorString(rc) : "nccl error"));
    throw CudaFail{PSFM_ERR_NCCL};
  }
}

void allreduce_sum(double* buf, size_t n, cudaStream_t stream) {
  if (g.world <= 1 || n == 0) return;
  check(g.AllReduce(buf, buf, n, ncclFloat64, ncclSum, g.comm, stream), "ncclAllReduce(sum)");
}
void allreduce_max(double* buf, size_t n, cudaStream_t stream) {
  if (g.world <= 1 || n == 0) return;
  check(g.AllReduce(buf, buf, n, ncclFloat64, ncclMax, g.comm, stream), "ncclAllReduce(max)");
}

}  // namespace dist
}  // namespace psfm

using namespace psfm;

extern "C" int psfm_dist_get_unique_id(uint8_t id[PSFM_NCCL_UNIQUE_ID_BYTES]) {
  if (!dist::load_nccl()) return PSFM_ERR_NCCL;
  dist::ncclUniqueId u;
  if (dist::g.GetUniqueId(&u) != dist::ncclSuccess) { set_error("ncclGetUniqueId failed"); return PSFM_ERR_NCCL; }
  memcpy(id, u.internal, PSFM_NCCL_UNIQUE_ID_BYTES);
  return PSFM_OK;
}

extern "C" int psfm_dist_init(const uint8_t id[PSFM_NCCL_UNIQUE_ID_BYTES], int32_t rank, int32_t world_size) {
  if (world_size < 1 || rank < 0 || rank >= world_size) { set_error("psfm_dist_init: bad rank/world"); return PSFM_ERR_INVALID; }
  if (dist::g.comm) { set_error("psfm_dist_init: already initialised"); return PSFM_ERR_INVALID; }
  if (world_size == 1) { dist::g.world = 1; dist::g.rank = 0; return PSFM_OK; }
  if (!dist::load_nccl()) return PSFM_ERR_NCCL;
  dist::ncclUniqueId u;
  memcpy(u.internal, id, PSFM_NCCL_UNIQUE_ID_BYTES);
  const int rc = dist::g.CommInitRank(&dist::g.comm, world_size, u, rank);
  if (rc != dist::ncclSuccess) {
    set_error(std::string("ncclCommInitRank: ") + (dist::g.GetErrorString ? dist::g.GetErrorString(rc) : "error"));
    dist::g.comm = nullptr;
    return PSFM_ERR_NCCL;
  }
  dist::g.world = world_size;
  dist::g.rank = rank;
  return PSFM_OK;
}

extern "C" int psfm_dist_world_size(void) { return dist::g.world; }
extern "C" int psfm_dist_rank(void) { return dist::g.rank; }
extern "C" void psfm_dist_finalize(void) {
  if (dist::g.comm) dist::g.CommDestroy(dist::g.comm);
  dist::g.comm = nullptr;
  dist::g.world = 1;
  dist::g.rank = 0;
}
