// nccl_dl.cuh -- the four NCCL entry points the sharded database search needs, bound at run time.
//
// liborbx_b200.so does not link libnccl: a single-GPU user (the operator() drop-in) must not need it, and a
// process that already carries an NCCL (torch bundles its own) must keep using THAT one, because a
// communicator is only meaningful to the library that created it.  The first use looks for an already
// loaded libnccl.so.2 (RTLD_NOLOAD), then loads it by soname; ORBX_NCCL_LIB overrides the path.
// Declarations follow nccl.h (NCCL 2.x ABI: ncclResult_t and ncclDataType_t are ints, ncclUniqueId is a
// 128-byte struct passed by value).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace orbx {

struct NcclUniqueId {
  char internal[128];
};

struct NcclApi {
  void* lib = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(void** comm, int nranks, NcclUniqueId id, int rank) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*CommCount)(void* comm, int* count) = nullptr;
  int (*CommUserRank)(void* comm, int* rank) = nullptr;
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t st) = nullptr;
  int (*GetVersion)(int* v) = nullptr;
  char err[200] = "";
};
constexpr int kNcclUint64 = 5;  // ncclDataType_t::ncclUint64

// nullptr when no NCCL can be loaded; `why` (may be NULL) then receives the reason
const NcclApi* nccl_api(const char** why);

}  // namespace orbx
