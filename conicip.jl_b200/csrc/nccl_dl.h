// Minimal NCCL binding resolved with dlopen("libnccl.so.2") at run time, so the library
// neither links a second NCCL next to the one the host process (PyTorch, Julia) already
// loaded nor needs NCCL at all for single-GPU use.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace cip {

struct NcclId {
  char internal[128];
};

struct NcclApi {
  bool loaded = false;
  int (*GetUniqueId)(NcclId* id) = nullptr;
  int (*CommInitRank)(void** comm, int nranks, NcclId id, int rank) = nullptr;
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm,
                   cudaStream_t stream) = nullptr;
  int (*Broadcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm,
                   cudaStream_t stream) = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*Reduce)(const void* send, void* recv, size_t count, int dtype, int op, int root, void* comm,
                cudaStream_t stream) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

// returns nullptr (and sets the error string) if libnccl cannot be loaded
const NcclApi* nccl_api();

constexpr int kNcclInt32 = 2;
constexpr int kNcclFloat64 = 8;
constexpr int kNcclSum = 0;
constexpr int kNcclMax = 2;

}  // namespace cip
