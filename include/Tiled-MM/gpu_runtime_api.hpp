// Drop-in counterpart of reference src/Tiled-MM/gpu_runtime_api.hpp (CUDA half only: there is no ROCm back end).
// Only the names the public path and its callers touch are kept; everything forwards to the CUDA runtime.
#pragma once
#include <cuda_runtime_api.h>
#include <utility>

#ifndef TILED_MM_CUDA
#define TILED_MM_CUDA
#endif

namespace gpu {
namespace runtime_api {

using StatusType = cudaError_t;
using StreamType = cudaStream_t;
using EventType = cudaEvent_t;

namespace status {
constexpr StatusType Success = cudaSuccess;
constexpr StatusType ErrorMemoryAllocation = cudaErrorMemoryAllocation;
constexpr StatusType ErrorInvalidValue = cudaErrorInvalidValue;
}  // namespace status

namespace flag {
constexpr auto HostRegisterDefault = cudaHostRegisterDefault;
constexpr auto StreamNonBlocking = cudaStreamNonBlocking;
constexpr auto MemcpyHostToDevice = cudaMemcpyHostToDevice;
constexpr auto MemcpyDeviceToHost = cudaMemcpyDeviceToHost;
constexpr auto EventDisableTiming = cudaEventDisableTiming;
}  // namespace flag

template <typename... A> inline StatusType host_register(A... a) { return cudaHostRegister(std::forward<A>(a)...); }
template <typename... A> inline StatusType host_unregister(A... a) { return cudaHostUnregister(std::forward<A>(a)...); }
template <typename... A> inline StatusType malloc(A... a) { return cudaMalloc(std::forward<A>(a)...); }
template <typename... A> inline StatusType free(A... a) { return cudaFree(std::forward<A>(a)...); }
template <typename... A> inline StatusType host_alloc(A... a) { return cudaHostAlloc(std::forward<A>(a)...); }
template <typename... A> inline StatusType memcpy(A... a) { return cudaMemcpy(std::forward<A>(a)...); }
template <typename... A> inline StatusType memcpy_async(A... a) { return cudaMemcpyAsync(std::forward<A>(a)...); }
template <typename... A> inline StatusType memcpy_2d_async(A... a) { return cudaMemcpy2DAsync(std::forward<A>(a)...); }
template <typename... A> inline StatusType get_device(A... a) { return cudaGetDevice(std::forward<A>(a)...); }
template <typename... A> inline StatusType set_device(A... a) { return cudaSetDevice(std::forward<A>(a)...); }
template <typename... A> inline StatusType mem_get_info(A... a) { return cudaMemGetInfo(std::forward<A>(a)...); }
template <typename... A> inline StatusType stream_synchronize(A... a) { return cudaStreamSynchronize(std::forward<A>(a)...); }
template <typename... A> inline StatusType stream_create_with_flags(A... a) { return cudaStreamCreateWithFlags(std::forward<A>(a)...); }
template <typename... A> inline StatusType stream_destroy(A... a) { return cudaStreamDestroy(std::forward<A>(a)...); }
template <typename... A> inline StatusType stream_wait_event(A... a) { return cudaStreamWaitEvent(std::forward<A>(a)...); }
template <typename... A> inline StatusType event_create_with_flags(A... a) { return cudaEventCreateWithFlags(std::forward<A>(a)...); }
template <typename... A> inline StatusType event_destroy(A... a) { return cudaEventDestroy(std::forward<A>(a)...); }
template <typename... A> inline StatusType event_record(A... a) { return cudaEventRecord(std::forward<A>(a)...); }
template <typename... A> inline StatusType event_synchronize(A... a) { return cudaEventSynchronize(std::forward<A>(a)...); }
template <typename... A> inline StatusType event_elapsed_time(A... a) { return cudaEventElapsedTime(std::forward<A>(a)...); }
inline const char* get_error_string(StatusType s) { return cudaGetErrorString(s); }
inline StatusType get_last_error() { return cudaGetLastError(); }
inline StatusType device_synchronize() { return cudaDeviceSynchronize(); }

}  // namespace runtime_api
}  // namespace gpu
