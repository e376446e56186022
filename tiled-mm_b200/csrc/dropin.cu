// Out-of-line definitions + explicit instantiations behind include/Tiled-MM/*.hpp, so the shared library
// exports the same mangled C++ symbols as the reference library (gpu::gemm<T>, gpu::mm_handle<T>::*,
// gpu::get_blas_operation; reference tiled_mm.cpp:626-668, mm_handle.cpp:167-170).
#include "../../include/Tiled-MM/tiled_mm.hpp"

#include <cctype>
#include <cstdint>
#include <tuple>

#define TMM_EXPORT __attribute__((visibility("default")))

namespace gpu {

TMM_EXPORT blas_api::OperationType get_blas_operation(char trans) {
    // reference tiled_mm.cpp:168-179: 'T' -> transpose, 'C' -> conjugate transpose, anything else -> none
    return trans == 'T' ? blas_api::operation::Transpose : (trans == 'C' ? blas_api::operation::ConjugateTranspose : blas_api::operation::None);
}

template <typename Scalar>
mm_handle<Scalar>::mm_handle(int streams, int max_tile_m, int max_tile_n, int max_tile_k)
    : max_tile_m_(max_tile_m), max_tile_n_(max_tile_n), max_tile_k_(max_tile_k), a_buff_(streams), b_buff_(streams), c_buff_(streams) {
    check_tmm_status(tmm_context_create(tmm_dtype<Scalar>::value, streams, max_tile_m, max_tile_n, max_tile_k, &ctx_));
    view_ = gpu_context(ctx_);
    full_c_.bind_to_context(ctx_);
}

template <typename Scalar>
mm_handle<Scalar>::~mm_handle() { tmm_context_destroy(ctx_); }

template <typename Scalar>
void mm_handle<Scalar>::set_num_streams(int streams) {
    // reference mm_handle.cpp:36-45: the stream count of the context and of the three slab buffers
    const tile_dim a = a_buff_.get_tile_sizes(), c = c_buff_.get_tile_sizes();  // keep the current tile hints (0 x 0 before the first set)
    check_tmm_status(tmm_context_set_streams_and_tiles(ctx_, streams, c.rows() > 0 ? c.rows() : max_tile_m_, c.cols() > 0 ? c.cols() : max_tile_n_,
                                                       a.cols() > 0 ? a.cols() : max_tile_k_));
    a_buff_.set_num_streams(streams);
    b_buff_.set_num_streams(streams);
    c_buff_.set_num_streams(streams);
}

template <typename Scalar>
int mm_handle<Scalar>::get_num_streams() { return tmm_context_get_num_streams(ctx_); }

template <typename Scalar>
gpu_context& mm_handle<Scalar>::get_gpu_context() { return view_; }

// The reference (re)allocates n_streams x tile slabs here (mm_handle.cpp:57-66).  The slabs keep that geometry (A tile
// m x k, B tile k x n, C tile m x n) and are allocated when first asked for; for the scheduler the sizes are staging hints
// (they cap k-chunk and column-block sizes, never results), bounded by the maxima fixed at construction.
template <typename Scalar>
void mm_handle<Scalar>::set_tile_sizes(int tile_m, int tile_n, int tile_k) {
    set_streams_and_tiles(get_num_streams(), tile_m, tile_n, tile_k);
}
template <typename Scalar>
void mm_handle<Scalar>::set_tile_sizes(int tile_size) { set_tile_sizes(tile_size, tile_size, tile_size); }

template <typename Scalar>
void mm_handle<Scalar>::set_full_sizes(int m, int n, int) {
    check_tmm_status(tmm_context_reserve_device_c(ctx_, m, n));  // mm_handle.cpp:73-80: full C = m * n elements
}

template <typename Scalar>
std::tuple<int, int, int> mm_handle<Scalar>::optimal_tile_sizes(int m, int n, int k) {
    // per dimension, a function of (dim, construction-time maximum) only (mm_handle.cpp:89-133)
    const int tm = tmm_optimal_tile_size(m, max_tile_m_), tn = tmm_optimal_tile_size(n, max_tile_n_), tk = tmm_optimal_tile_size(k, max_tile_k_);
    if (tm < 0 || tn < 0 || tk < 0) { check_tmm_status(TMM_ERR_INVALID); }
    return std::make_tuple(tm, tn, tk);
}

template <typename Scalar>
std::tuple<int, int, int> mm_handle<Scalar>::get_max_tile_sizes() { return std::make_tuple(max_tile_m_, max_tile_n_, max_tile_k_); }

template <typename Scalar>
void mm_handle<Scalar>::set_streams_and_tiles(int streams, int tile_m, int tile_n, int tile_k) {
    // mm_handle.cpp:135-145 (asserts tile <= max there; clamped here)
    tile_m = tile_m < max_tile_m_ ? tile_m : max_tile_m_;
    tile_n = tile_n < max_tile_n_ ? tile_n : max_tile_n_;
    tile_k = tile_k < max_tile_k_ ? tile_k : max_tile_k_;
    check_tmm_status(tmm_context_set_streams_and_tiles(ctx_, streams, tile_m, tile_n, tile_k));
    a_buff_.set_streams_and_tiles(streams, tile_dim(tile_m, tile_k));
    b_buff_.set_streams_and_tiles(streams, tile_dim(tile_k, tile_n));
    c_buff_.set_streams_and_tiles(streams, tile_dim(tile_m, tile_n));
}

template <typename Scalar>
device_buffer<Scalar>& mm_handle<Scalar>::get_device_buffer_a() { return a_buff_; }
template <typename Scalar>
device_buffer<Scalar>& mm_handle<Scalar>::get_device_buffer_b() { return b_buff_; }
template <typename Scalar>
device_buffer<Scalar>& mm_handle<Scalar>::get_device_buffer_c() { return c_buff_; }

template <typename Scalar>
device_vector<Scalar>& mm_handle<Scalar>::get_full_device_buffer_c() { return full_c_; }

template <typename Scalar>
void gemm64(mm_handle<Scalar>& handle, char trans_a, char trans_b, long long m, long long n, long long k, Scalar alpha, Scalar* a, long long ld_a,
            Scalar* b, long long ld_b, Scalar beta, Scalar* c, long long ld_c, bool pin_host_buffers, bool copy_c_back) {
    if (m > 0 && n > 0 && k > 0 && m <= INT32_MAX && n <= INT32_MAX && k <= INT32_MAX) {
        // reference tiled_mm.cpp:556-563: every call re-derives the tile sizes from the problem and records them in the handle
        // (visible through get_device_buffer_*().get_tile_sizes()); here they are staging hints for the scheduler
        int tm, tn, tk;
        std::tie(tm, tn, tk) = handle.optimal_tile_sizes((int)m, (int)n, (int)k);
        handle.set_tile_sizes(tm, tn, tk);
    }
    check_tmm_status(tmm_gemm(handle.native(), trans_a, trans_b, m, n, k, &alpha, a, ld_a, b, ld_b, &beta, c, ld_c, pin_host_buffers ? 1 : 0,
                              copy_c_back ? 1 : 0));
}

template <typename Scalar>
void gemm(mm_handle<Scalar>& handle, char trans_a, char trans_b, int m, int n, int k, Scalar alpha, Scalar* a, int ld_a, Scalar* b, int ld_b,
          Scalar beta, Scalar* c, int ld_c, bool pin_host_buffers, bool copy_c_back) {
    gemm64<Scalar>(handle, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, pin_host_buffers, copy_c_back);
}

using zfloat = std::complex<float>;
using zdouble = std::complex<double>;

#define TMM_INSTANTIATE(T)                                                                                                              \
    template class TMM_EXPORT mm_handle<T>;                                                                                             \
    template TMM_EXPORT void gemm<T>(mm_handle<T>&, char, char, int, int, int, T, T*, int, T*, int, T, T*, int, bool, bool);            \
    template TMM_EXPORT void gemm64<T>(mm_handle<T>&, char, char, long long, long long, long long, T, T*, long long, T*, long long, T, T*, \
                                       long long, bool, bool);
TMM_INSTANTIATE(float)
TMM_INSTANTIATE(double)
TMM_INSTANTIATE(zfloat)
TMM_INSTANTIATE(zdouble)

}  // namespace gpu
