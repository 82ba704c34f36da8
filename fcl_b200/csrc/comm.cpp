// comm.cpp — fclgpu_comm_*: the multi-GPU wrapper of the C ABI.
//
// The path shards over poses with no exchange inside the traversal: BVHs are replicated, rank r owns the contiguous
// block fclgpu_shard_range() names, and the per-rank result records are gathered with ncclAllGather over NVLink.
// This file gives a C / C++ caller of libfclgpu.so that collective without torch: NCCL is opened at run time
// (dlopen("libnccl.so.2"): the library has no link-time dependency on it and reports FCLGPU_ERR_UNSUPPORTED_FUNCTION
// where NCCL is not installed), the 128-byte NCCL unique id is created on rank 0 and handed to the other ranks by
// whatever the application already uses to start its processes (MPI, a file, torch.distributed's store, ...).
// Collectives are asynchronous on the caller's stream: to overlap the gather of batch k with the traversal of batch
// k + 1, enqueue them on two streams ordered by an event (bench.py does exactly that).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/fclgpu.h"

namespace {

struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
typedef int NcclResult;  // ncclSuccess == 0
constexpr int kNcclUint8 = 1;

struct NcclApi {
  void* handle = nullptr;
  NcclResult (*GetUniqueId)(NcclId*) = nullptr;
  NcclResult (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  NcclResult (*CommDestroy)(NcclComm) = nullptr;
  NcclResult (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  NcclResult (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  NcclResult (*GroupStart)() = nullptr;
  NcclResult (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(NcclResult) = nullptr;
  bool ok = false;
};

thread_local char g_comm_err[256] = "";
int cfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_comm_err, sizeof g_comm_err, fmt, ap);
  va_end(ap);
  return code;
}

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define LOAD(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllGather, "ncclAllGather");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.GroupStart &&
             api.GroupEnd;
  });
  return api;
}

int nccl_status(NcclResult r, const char* what) {
  if (r == 0) return FCLGPU_OK;
  NcclApi& a = nccl();
  return cfail(FCLGPU_ERR_COMM, "%s failed: %s", what, a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
}

}  // namespace

struct fclgpu_comm {
  NcclComm comm = nullptr;
  int device = -1, rank = 0, world = 1;
};

extern "C" const char* fclgpu_comm_last_error(void) { return g_comm_err; }

extern "C" void fclgpu_shard_range(int64_t n, int rank, int world, int64_t* start, int64_t* count) {
  if (world < 1) world = 1;
  const int64_t base = n / world, rem = n % world;
  const int64_t s = rank * base + (rank < rem ? rank : rem);
  if (start) *start = s;
  if (count) *count = base + (rank < rem ? 1 : 0);
}

extern "C" int fclgpu_comm_unique_id(char id[FCLGPU_COMM_ID_BYTES]) {
  if (!id) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "id is NULL");
  NcclApi& a = nccl();
  if (!a.ok) return cfail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "NCCL (libnccl.so.2) is not available: %s", dlerror() ? dlerror() : "symbols missing");
  NcclId u;
  const int rc = nccl_status(a.GetUniqueId(&u), "ncclGetUniqueId");
  if (rc) return rc;
  std::memcpy(id, u.internal, sizeof u.internal);
  return FCLGPU_OK;
}

extern "C" int fclgpu_comm_init(int device, int rank, int world, const char id[FCLGPU_COMM_ID_BYTES], fclgpu_comm** out) {
  if (!out) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!id || world < 1 || rank < 0 || rank >= world) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "bad rank / world / id");
  NcclApi& a = nccl();
  if (!a.ok) return cfail(FCLGPU_ERR_UNSUPPORTED_FUNCTION, "NCCL (libnccl.so.2) is not available");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return cfail(FCLGPU_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  if (cudaSetDevice(device) != cudaSuccess) return cfail(FCLGPU_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  NcclId u;
  std::memcpy(u.internal, id, sizeof u.internal);
  fclgpu_comm* c = new fclgpu_comm;
  c->device = device;
  c->rank = rank;
  c->world = world;
  const int rc = nccl_status(a.CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank");
  if (rc) {
    delete c;
    return rc;
  }
  *out = c;
  return FCLGPU_OK;
}

extern "C" int fclgpu_comm_rank(const fclgpu_comm* c) { return c ? c->rank : -1; }
extern "C" int fclgpu_comm_world(const fclgpu_comm* c) { return c ? c->world : 0; }

extern "C" int fclgpu_comm_allgather(fclgpu_comm* c, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  if (!c || !send || !recv) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL comm / buffer");
  if (bytes_per_rank == 0) return FCLGPU_OK;
  if (cudaSetDevice(c->device) != cudaSuccess) return cfail(FCLGPU_ERR_CUDA, "cudaSetDevice failed");
  return nccl_status(nccl().AllGather(send, recv, bytes_per_rank, kNcclUint8, c->comm, (cudaStream_t)stream), "ncclAllGather");
}

extern "C" int fclgpu_comm_allgather_ragged(fclgpu_comm* c, const void* send, void* recv, const int64_t* bytes_of_rank,
                                            void* stream) {
  if (!c || !recv || !bytes_of_rank) return cfail(FCLGPU_ERR_INVALID_ARGUMENT, "NULL comm / buffer / sizes");
  if (cudaSetDevice(c->device) != cudaSuccess) return cfail(FCLGPU_ERR_CUDA, "cudaSetDevice failed");
  NcclApi& a = nccl();
  int rc = nccl_status(a.GroupStart(), "ncclGroupStart");
  if (rc) return rc;
  int64_t off = 0;
  for (int r = 0; r < c->world && rc == 0; ++r) {
    if (bytes_of_rank[r] < 0) rc = cfail(FCLGPU_ERR_INVALID_ARGUMENT, "negative size for rank %d", r);
    else if (bytes_of_rank[r] > 0)
      rc = nccl_status(a.Broadcast(r == c->rank ? send : (const void*)((char*)recv + off), (char*)recv + off, (size_t)bytes_of_rank[r], kNcclUint8, r, c->comm,
                                   (cudaStream_t)stream), "ncclBroadcast");
    off += bytes_of_rank[r] > 0 ? bytes_of_rank[r] : 0;
  }
  const int rc2 = nccl_status(a.GroupEnd(), "ncclGroupEnd");
  return rc ? rc : rc2;
}

extern "C" int fclgpu_comm_destroy(fclgpu_comm* c) {
  if (!c) return FCLGPU_OK;
  int rc = FCLGPU_OK;
  if (c->comm) {
    cudaSetDevice(c->device);
    rc = nccl_status(nccl().CommDestroy(c->comm), "ncclCommDestroy");
  }
  delete c;
  return rc;
}
