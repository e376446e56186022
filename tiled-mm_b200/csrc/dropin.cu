// Out-of-line definitions + explicit instantiations behind include/Tiled-MM/*.hpp, so the shared library
// exports the same mangled C++ symbols as the reference library (gpu::gemm<T>, gpu::mm_handle<T>::*,
// gpu::get_blas_operation; reference tiled_mm.cpp:626-668, mm_handle.cpp:167-170).
#include "../../include/Tiled-MM/tiled_mm.hpp"

#include <cctype>

#define TMM_EXPORT __attribute__((visibility("default")))

namespace gpu {

TMM_EXPORT blas_api::OperationType get_blas_operation(char trans) {
    // reference tiled_mm.cpp:168-179: 'T' -> transpose, 'C' -> conjugate transpose, anything else -> none
    return trans == 'T' ? blas_api::operation::Transpose : (trans == 'C' ? blas_api::operation::ConjugateTranspose : blas_api::operation::None);
}

template <typename Scalar>
mm_handle<Scalar>::mm_handle(int streams, int max_tile_m, int max_tile_n, int max_tile_k) {
    check_tmm_status(tmm_context_create(tmm_dtype<Scalar>::value, streams, max_tile_m, max_tile_n, max_tile_k, &ctx_));
    view_ = gpu_context(ctx_);
    full_c_.bind_to_context(ctx_);
}

template <typename Scalar>
mm_handle<Scalar>::~mm_handle() { tmm_context_destroy(ctx_); }

template <typename Scalar>
void mm_handle<Scalar>::set_num_streams(int streams) {
    int tm, tn, tk;
    check_tmm_status(tmm_context_get_max_tile_sizes(ctx_, &tm, &tn, &tk));
    check_tmm_status(tmm_context_set_streams_and_tiles(ctx_, streams, tm, tn, tk));
}

template <typename Scalar>
int mm_handle<Scalar>::get_num_streams() { return tmm_context_get_num_streams(ctx_); }

template <typename Scalar>
gpu_context& mm_handle<Scalar>::get_gpu_context() { return view_; }

// In the reference these (re)allocate the per-stream tile slabs (mm_handle.cpp:57-80).  Device storage here is sized
// per call from the problem and the free HBM, so they only record the hint / are no-ops.
template <typename Scalar>
void mm_handle<Scalar>::set_tile_sizes(int, int, int) {}
template <typename Scalar>
void mm_handle<Scalar>::set_tile_sizes(int) {}
template <typename Scalar>
void mm_handle<Scalar>::set_full_sizes(int, int, int) {}

template <typename Scalar>
std::tuple<int, int, int> mm_handle<Scalar>::optimal_tile_sizes(int m, int n, int k) {
    int tm, tn, tk;
    check_tmm_status(tmm_context_optimal_tile_sizes(ctx_, m, n, k, &tm, &tn, &tk));
    return std::make_tuple(tm, tn, tk);
}

template <typename Scalar>
std::tuple<int, int, int> mm_handle<Scalar>::get_max_tile_sizes() {
    int tm, tn, tk;
    check_tmm_status(tmm_context_get_max_tile_sizes(ctx_, &tm, &tn, &tk));
    return std::make_tuple(tm, tn, tk);
}

template <typename Scalar>
void mm_handle<Scalar>::set_streams_and_tiles(int streams, int tile_m, int tile_n, int tile_k) {
    check_tmm_status(tmm_context_set_streams_and_tiles(ctx_, streams, tile_m, tile_n, tile_k));
}

template <typename Scalar>
device_vector<Scalar>& mm_handle<Scalar>::get_full_device_buffer_c() { return full_c_; }

template <typename Scalar>
void gemm64(mm_handle<Scalar>& handle, char trans_a, char trans_b, long long m, long long n, long long k, Scalar alpha, Scalar* a, long long ld_a,
            Scalar* b, long long ld_b, Scalar beta, Scalar* c, long long ld_c, bool pin_host_buffers, bool copy_c_back) {
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
