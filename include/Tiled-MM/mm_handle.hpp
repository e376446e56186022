// Drop-in counterpart of reference src/Tiled-MM/mm_handle.hpp (mm_handle<Scalar>, make_context).
// Same public surface; the object now owns a tmm_context: device panel/ring buffers, H2D / compute / D2H
// streams, an event pool and (optionally) a host-registration cache, instead of n_streams x {A,B,C} tile
// slabs plus one cuBLAS handle per stream (reference mm_handle.hpp:44-57, gpu_context.cpp:6-18).
#pragma once
#include "../tiled_mm_b200.h"
#include "device_buffer.hpp"
#include "device_vector.hpp"
#include "gpu_context.hpp"
#include "gpu_runtime_api.hpp"

#include <complex>
#include <memory>
#include <tuple>

namespace gpu {

template <typename Scalar> struct tmm_dtype;
template <> struct tmm_dtype<float> { static constexpr int value = TMM_F32; };
template <> struct tmm_dtype<double> { static constexpr int value = TMM_F64; };
template <> struct tmm_dtype<std::complex<float>> { static constexpr int value = TMM_C32; };
template <> struct tmm_dtype<std::complex<double>> { static constexpr int value = TMM_C64; };

template <typename Scalar>
class mm_handle {
public:
    // construction fixes the maxima that get_max_tile_sizes() reports; nothing is allocated yet
    mm_handle(int n_streams, int tile_m_max, int tile_n_max, int tile_k_max);
    ~mm_handle();
    mm_handle(mm_handle&&) = delete;
    mm_handle(const mm_handle&) = delete;
    mm_handle& operator=(const mm_handle&& other) = delete;

    // stream count and tile geometry: staging hints for the scheduler (never results) and the geometry of the slabs below
    int get_num_streams();
    void set_num_streams(int n_streams);
    void set_tile_sizes(int tile_m, int tile_n, int tile_k);
    void set_tile_sizes(int tile);
    void set_streams_and_tiles(int n_streams, int tile_m, int tile_n, int tile_k);
    std::tuple<int, int, int> get_max_tile_sizes();
    std::tuple<int, int, int> optimal_tile_sizes(int m, int n, int k);  // per dimension: what the reference would tile (m, n, k) with

    void set_full_sizes(int m, int n, int k);  // size the device-resident C to m x n now
    gpu_context& get_gpu_context();            // the context's streams, as the reference's accessor type

    // per-stream tile slabs (n_streams x tile): caller-visible device scratch, allocated on first use; the scheduler itself
    // stages through context-owned panels / rings instead (csrc/tmm_context.cu)
    device_buffer<Scalar>& get_device_buffer_a();
    device_buffer<Scalar>& get_device_buffer_b();
    device_buffer<Scalar>& get_device_buffer_c();
    // device C of the last copy_c_back=false gemm: column-major m x n, ld = m (reference README.md:102-103)
    device_vector<Scalar>& get_full_device_buffer_c();

    tmm_context* native() { return ctx_; }

private:
    tmm_context* ctx_ = nullptr;
    gpu_context view_{nullptr};
    int max_tile_m_ = 5000, max_tile_n_ = 5000, max_tile_k_ = 5000;  // fixed at construction (mm_handle.cpp:10-16)
    device_buffer<Scalar> a_buff_, b_buff_, c_buff_;
    device_vector<Scalar> full_c_;
};

template <typename Scalar>
std::unique_ptr<mm_handle<Scalar>> make_context(int streams, int max_tile_m, int max_tile_n, int max_tile_k) {
    return std::make_unique<mm_handle<Scalar>>(streams, max_tile_m, max_tile_n, max_tile_k);
}

template <typename Scalar>
std::unique_ptr<mm_handle<Scalar>> make_context() {
    return std::make_unique<mm_handle<Scalar>>(2, 5000, 5000, 5000);  // reference defaults, mm_handle.hpp:70-76
}

}  // namespace gpu
