// NCCL entry points resolved at run time (dlopen of libnccl.so.2), so the library links and loads on a box without
// NCCL and a single-GPU user never touches it.  Only the handful of calls the panel exchange needs (tmm_dist.cu).
// Declarations restate the public NCCL 2.x ABI (nccl.h): opaque communicator, 128-byte unique id passed by value.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace tmm {
namespace nccl {

struct UniqueId { char internal[128]; };
typedef struct ncclComm* Comm;
enum DataType : int { Int8 = 0, Uint8 = 1, Int32 = 2, Int64 = 4, Uint64 = 5, Float64 = 8 };
enum RedOp : int { Sum = 0, Prod = 1, Max = 2, Min = 3 };

struct Api {
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
    const char* why = "";
};

// Loads libnccl.so.2 once (the copy already mapped into the process, e.g. torch's, wins); api().ok == false if absent.
const Api& api();

}  // namespace nccl
}  // namespace tmm
