// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" doorway into the UNMODIFIED reference library, compiled from the sources where they
// lie under /root/reference/src (see oracle/Makefile; nothing is copied into this repo).
// Everything in namespace gpu:: is built with hidden visibility so this .so can live in the same
// process as the product library, which exports the same C++ names.
//
// Two builds share this file:
//   oracle/_ref/libtiledmm_ref.so      reference + real libcudart/libcublas   (GPU box)
//   oracle/_ref/libtiledmm_ref_cpu.so  reference + oracle/cpu_cuda_emul.cpp   (runs anywhere)
#include <Tiled-MM/tiled_mm.hpp>
#include <Tiled-MM/device_vector.hpp>
#include <Tiled-MM/util.hpp>

#include <complex>
#include <cstddef>
#include <cstdio>
#include <exception>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
using zf = std::complex<float>;
using zd = std::complex<double>;

template <typename F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { std::fprintf(stderr, "[ref_shim] exception: %s\n", e.what()); return -1; }
    catch (...) { return -1; }
}

template <typename T>
int do_gemm(void* ctx, char ta, char tb, int m, int n, int k, const void* alpha, void* a, int lda,
            void* b, int ldb, const void* beta, void* c, int ldc, int pin, int copy_c_back) {
    return guarded([&] {
        gpu::gemm<T>(*static_cast<gpu::mm_handle<T>*>(ctx), ta, tb, m, n, k, *static_cast<const T*>(alpha),
                     static_cast<T*>(a), lda, static_cast<T*>(b), ldb, *static_cast<const T*>(beta),
                     static_cast<T*>(c), ldc, pin != 0, copy_c_back != 0);
    });
}
}  // namespace

REF_API void* ref_ctx_create(int dtype, int streams, int tile_m, int tile_n, int tile_k) {
    void* p = nullptr;
    guarded([&] {
        switch (dtype) {
        case 0: p = gpu::make_context<float>(streams, tile_m, tile_n, tile_k).release(); break;
        case 1: p = gpu::make_context<double>(streams, tile_m, tile_n, tile_k).release(); break;
        case 2: p = gpu::make_context<zf>(streams, tile_m, tile_n, tile_k).release(); break;
        case 3: p = gpu::make_context<zd>(streams, tile_m, tile_n, tile_k).release(); break;
        }
    });
    return p;
}

REF_API void ref_ctx_destroy(int dtype, void* ctx) {
    guarded([&] {
        switch (dtype) {
        case 0: delete static_cast<gpu::mm_handle<float>*>(ctx); break;
        case 1: delete static_cast<gpu::mm_handle<double>*>(ctx); break;
        case 2: delete static_cast<gpu::mm_handle<zf>*>(ctx); break;
        case 3: delete static_cast<gpu::mm_handle<zd>*>(ctx); break;
        }
    });
}

REF_API int ref_gemm(void* ctx, int dtype, char ta, char tb, int m, int n, int k, const void* alpha,
                     void* a, int lda, void* b, int ldb, const void* beta, void* c, int ldc,
                     int pin_host_buffers, int copy_c_back) {
    switch (dtype) {
    case 0: return do_gemm<float>(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers, copy_c_back);
    case 1: return do_gemm<double>(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers, copy_c_back);
    case 2: return do_gemm<zf>(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers, copy_c_back);
    case 3: return do_gemm<zd>(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers, copy_c_back);
    }
    return -2;
}

// Fetch `count` elements of the context's full device C (column-major m x n, ld = m) to the host,
// the way tests/test-multiply.cpp:339 does.
REF_API int ref_fetch_device_c(void* ctx, int dtype, void* dst, size_t count) {
    return guarded([&] {
        switch (dtype) {
        case 0: gpu::copy_to_host(static_cast<gpu::mm_handle<float>*>(ctx)->get_full_device_buffer_c().data(), static_cast<float*>(dst), count); break;
        case 1: gpu::copy_to_host(static_cast<gpu::mm_handle<double>*>(ctx)->get_full_device_buffer_c().data(), static_cast<double*>(dst), count); break;
        case 2: gpu::copy_to_host(static_cast<gpu::mm_handle<zf>*>(ctx)->get_full_device_buffer_c().data(), static_cast<zf*>(dst), count); break;
        case 3: gpu::copy_to_host(static_cast<gpu::mm_handle<zd>*>(ctx)->get_full_device_buffer_c().data(), static_cast<zd*>(dst), count); break;
        }
    });
}

REF_API int ref_optimal_tile_sizes(void* ctx, int dtype, int m, int n, int k, int* out3) {
    return guarded([&] {
        std::tuple<int, int, int> t;
        switch (dtype) {
        case 0: t = static_cast<gpu::mm_handle<float>*>(ctx)->optimal_tile_sizes(m, n, k); break;
        case 1: t = static_cast<gpu::mm_handle<double>*>(ctx)->optimal_tile_sizes(m, n, k); break;
        case 2: t = static_cast<gpu::mm_handle<zf>*>(ctx)->optimal_tile_sizes(m, n, k); break;
        default: t = static_cast<gpu::mm_handle<zd>*>(ctx)->optimal_tile_sizes(m, n, k); break;
        }
        out3[0] = std::get<0>(t); out3[1] = std::get<1>(t); out3[2] = std::get<2>(t);
    });
}

REF_API void* ref_malloc_pinned(size_t bytes) {
    void* p = nullptr;
    guarded([&] { p = gpu::malloc_pinned<char>(bytes, 0); });
    return p;
}

REF_API void ref_free_pinned(void* p) { gpu::runtime_api::StatusType s = GPU_PREFIX(FreeHost)(p); (void)s; }
