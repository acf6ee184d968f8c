#include "nccl_dl.h"

#include <dlfcn.h>

#include <mutex>

namespace cip {
void set_error(const char* fmt, ...);

const NcclApi* nccl_api() {
  static NcclApi api;
  static std::mutex mu;                       // several device threads of one process may get here together
  std::lock_guard<std::mutex> lock(mu);
  if (api.loaded) return &api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
    return nullptr;
  }
#define CIP_SYM(field, name)                                       \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
  if (!api.field) {                                                \
    set_error("dlsym(%s) failed", name);                           \
    return nullptr;                                                \
  }
  CIP_SYM(GetUniqueId, "ncclGetUniqueId")
  CIP_SYM(CommInitRank, "ncclCommInitRank")
  CIP_SYM(AllReduce, "ncclAllReduce")
  CIP_SYM(Broadcast, "ncclBroadcast")
  CIP_SYM(CommInitAll, "ncclCommInitAll")
  CIP_SYM(Reduce, "ncclReduce")
  CIP_SYM(GroupStart, "ncclGroupStart")
  CIP_SYM(GroupEnd, "ncclGroupEnd")
  CIP_SYM(CommDestroy, "ncclCommDestroy")
  CIP_SYM(GetErrorString, "ncclGetErrorString")
#undef CIP_SYM
  api.loaded = true;
  return &api;
}

}  // namespace cip
