// Drop-in counterpart of reference src/Tiled-MM/util.hpp: error checking, malloc_pinned / malloc_device,
// copy_to_device / copy_to_host (+ async).  Same names, arguments and error behaviour
// (message on stderr + std::runtime_error("GPU ERROR"), reference util.hpp:13-27).
#pragma once
#include "../tiled_mm_b200.h"
#include "gpu_blas_api.hpp"
#include "gpu_runtime_api.hpp"

#include <algorithm>
#include <cstddef>
#include <iostream>
#include <stdexcept>
#include <string>

namespace gpu {

static inline void check_runtime_status(runtime_api::StatusType status) {
    if (status != runtime_api::status::Success) {
        std::cerr << "error: GPU API call : " << runtime_api::get_error_string(status) << std::endl;
        throw(std::runtime_error("GPU ERROR"));
    }
}

static inline void check_blas_status(blas_api::StatusType status) {
    if (status != blas_api::status::Success) {
        std::cerr << "error: BLAS API call: " << blas_api::status::get_string(status) << std::endl;
        throw(std::runtime_error("GPU ERROR"));
    }
}

// a negative tiled_mm_b200 code -> the reference's exception
static inline void check_tmm_status(int status) {
    if (status != TMM_OK) {
        std::cerr << "error: GPU API call : " << tmm_last_error() << std::endl;
        throw(std::runtime_error("GPU ERROR"));
    }
}

static inline void check_last_device_kernel(std::string const& errstr) {
    auto status = runtime_api::get_last_error();
    if (status != runtime_api::status::Success) {
        std::cout << "error: GPU kernel launch : " << errstr << " : " << runtime_api::get_error_string(status) << std::endl;
        throw(std::runtime_error("GPU ERROR"));
    }
}

inline std::size_t gpu_allocated_memory() {
    runtime_api::device_synchronize();
    check_runtime_status(runtime_api::get_last_error());
    std::size_t free_b = 0, total_b = 0;
    auto status = runtime_api::mem_get_info(&free_b, &total_b);
    return status == runtime_api::status::Success ? total_b - free_b : std::size_t(-1);
}

template <typename T>
T* malloc_device(std::size_t n) {
    void* p = nullptr;
    check_tmm_status(tmm_malloc_device(n * sizeof(T), &p));
    return static_cast<T*>(p);
}

// cudaHostAlloc(flags 0) + fill; caller owns the memory (release with tmm_free_pinned / cudaFreeHost)
template <typename T>
T* malloc_pinned(std::size_t N, T value = T()) {
    void* p = nullptr;
    check_tmm_status(tmm_malloc_pinned(N * sizeof(T), &p));
    T* ptr = static_cast<T*>(p);
    std::fill(ptr, ptr + N, value);
    return ptr;
}

template <typename T>
void copy_to_device(const T* from, T* to, std::size_t n) {
    runtime_api::memcpy(to, from, n * sizeof(T), runtime_api::flag::MemcpyHostToDevice);
}

template <typename T>
void copy_to_host(const T* from, T* to, std::size_t n) {
    runtime_api::memcpy(to, from, n * sizeof(T), runtime_api::flag::MemcpyDeviceToHost);
}

template <typename T>
void copy_to_device_async(const T* from, T* to, std::size_t n, runtime_api::StreamType stream = NULL) {
    check_runtime_status(runtime_api::memcpy_async(to, from, n * sizeof(T), runtime_api::flag::MemcpyHostToDevice, stream));
}

template <typename T>
void copy_to_host_async(const T* from, T* to, std::size_t n, runtime_api::StreamType stream = NULL) {
    check_runtime_status(runtime_api::memcpy_async(to, from, n * sizeof(T), runtime_api::flag::MemcpyDeviceToHost, stream));
}

}  // namespace gpu
