// Drop-in counterpart of reference src/Tiled-MM/gpu_blas_handle.hpp.  The reference wraps a cublasHandle_t
// (gpu_blas_handle.hpp:11-17); this library calls no vendor BLAS, and the only state a handle carried that matters to the
// hand-written kernels is the stream they run on - so a handle IS a stream (blas_api::HandleType), the default stream
// unless set_stream() is called.  Callers written against the reference, such as tests/test-multiply.cpp:40-51
// (gpu_blas_handle handle; blas_api::dgemm(handle.handle(), ...)), compile and run unchanged.
#pragma once
#include "gpu_blas_api.hpp"
#include "gpu_runtime_api.hpp"

namespace gpu {

class gpu_blas_handle {
public:
    gpu_blas_handle() = default;
    gpu_blas_handle(gpu_blas_handle&& other) noexcept : stream_(other.stream_) { other.stream_ = nullptr; }
    gpu_blas_handle& operator=(gpu_blas_handle&& other) noexcept {
        stream_ = other.stream_;
        other.stream_ = nullptr;
        return *this;
    }
    gpu_blas_handle(gpu_blas_handle&) = delete;
    gpu_blas_handle& operator=(gpu_blas_handle&) = delete;

    blas_api::HandleType handle() const { return stream_; }
    // counterpart of blas_api::set_stream(handle, stream) (gpu_context.cpp:12-14)
    void set_stream(runtime_api::StreamType s) { stream_ = s; }

private:
    void* stream_ = nullptr;  // cudaStream_t; nullptr = the default stream
};

}  // namespace gpu
