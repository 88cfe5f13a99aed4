// comm.cc -- see comm.h.
#include "comm.h"

#include <dlfcn.h>

namespace qb {

const NcclApi *nccl_api(std::string *err) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string why;
  if (!tried) {
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      why = std::string("cannot load libnccl.so.2: ") + dlerror();
    } else {
      ok = true;
#define QB_SYM(field, name)                                           \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
  if (!api.field) {                                                   \
    ok = false;                                                       \
    why = std::string("libnccl lacks ") + name;                       \
  }
      QB_SYM(GetUniqueId, "ncclGetUniqueId")
      QB_SYM(CommInitRank, "ncclCommInitRank")
      QB_SYM(CommDestroy, "ncclCommDestroy")
      QB_SYM(GroupStart, "ncclGroupStart")
      QB_SYM(GroupEnd, "ncclGroupEnd")
      QB_SYM(Send, "ncclSend")
      QB_SYM(Recv, "ncclRecv")
      QB_SYM(AllReduce, "ncclAllReduce")
      QB_SYM(AllGather, "ncclAllGather")
      QB_SYM(Broadcast, "ncclBroadcast")
      QB_SYM(GetErrorString, "ncclGetErrorString")
#undef QB_SYM
    }
  }
  if (!ok) {
    if (err) *err = why;
    return nullptr;
  }
  return &api;
}

}  // namespace qb
