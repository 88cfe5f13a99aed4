// comm.h -- NCCL, loaded at run time.  The library must load (and the CPU tests must run) on
// machines without NCCL, and inside a process where torch has already loaded its own copy
// (same SONAME libnccl.so.2, so dlopen hands back that copy instead of a second one).
#ifndef QCC_B200_CSRC_COMM_H_
#define QCC_B200_CSRC_COMM_H_

#include <cuda_runtime.h>
#include <nccl.h>

#include <string>

namespace qb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

// nullptr + *err on failure.
const NcclApi *nccl_api(std::string *err);

}  // namespace qb

#endif  // QCC_B200_CSRC_COMM_H_
