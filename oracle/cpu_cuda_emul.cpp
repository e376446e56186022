// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU emulation of the ~25 CUDA-runtime and 7 cuBLAS entry points the reference library calls
// (list: SURVEY.md §8c).  Linked INSTEAD of libcudart/libcublas, it lets the unmodified reference
// scheduler (tiled_mm.cpp round_robin*, tiled_matrix tiling, device_buffer slabs) execute in a
// GPU-less container.  The reference enqueues work in an order that is a valid serialisation of
// its stream/event DAG, so executing every call synchronously at enqueue time is a legal schedule.
// "Device" memory is malloc()ed and poisoned with NaN so stale-slab reads show up.
// The xGEMM below is the naive column-major triple loop honouring beta == 0 (C not read).
#include <cuda_runtime_api.h>
#include <cublas_v2.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

static std::atomic<uint64_t> g_h2d{0}, g_d2h{0}, g_gemm_calls{0};

extern "C" {

// --- counters the tests read -------------------------------------------------------------
__attribute__((visibility("default"))) void emul_reset_counters() { g_h2d = 0; g_d2h = 0; g_gemm_calls = 0; }
__attribute__((visibility("default"))) uint64_t emul_h2d_bytes() { return g_h2d; }
__attribute__((visibility("default"))) uint64_t emul_d2h_bytes() { return g_d2h; }
__attribute__((visibility("default"))) uint64_t emul_gemm_calls() { return g_gemm_calls; }

// --- runtime -----------------------------------------------------------------------------
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = std::malloc(bytes ? bytes : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    // poison: a double/float NaN pattern in every 8 bytes
    uint64_t nan64 = 0x7ff8dead7fc0beefULL;
    for (size_t i = 0; i + 8 <= bytes; i += 8) std::memcpy(static_cast<char*>(*p) + i, &nan64, 8);
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static void count(size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) g_h2d += bytes;
    if (kind == cudaMemcpyDeviceToHost) g_d2h += bytes;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) { std::memcpy(dst, src, bytes); count(bytes, kind); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) { return cudaMemcpy(dst, src, bytes, kind); }
cudaError_t cudaMemcpy2D(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind) {
    if (width > dpitch || width > spitch) return cudaErrorInvalidPitchValue;
    for (size_t r = 0; r < height; ++r)
        std::memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    count(width * height, kind);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t) {
    return cudaMemcpy2D(dst, dpitch, src, spitch, width, height, kind);
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { std::memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { std::memset(a, 0, sizeof(*a)); return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t* fr, size_t* to) { *fr = size_t(1) << 36; *to = size_t(1) << 37; return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// --- cuBLAS --------------------------------------------------------------------------------
cublasStatus_t cublasCreate_v2(cublasHandle_t* h) { *h = reinterpret_cast<cublasHandle_t>(std::malloc(8)); return CUBLAS_STATUS_SUCCESS; }
cublasStatus_t cublasDestroy_v2(cublasHandle_t h) { std::free(h); return CUBLAS_STATUS_SUCCESS; }
cublasStatus_t cublasSetStream_v2(cublasHandle_t, cudaStream_t) { return CUBLAS_STATUS_SUCCESS; }
}  // extern "C"

namespace {
template <typename T> T cj(T v) { return v; }
template <typename R> std::complex<R> cj(std::complex<R> v) { return std::conj(v); }

template <typename T>
cublasStatus_t gemm(cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const T* alpha, const T* a, int lda,
                    const T* b, int ldb, const T* beta, T* c, int ldc) {
    ++g_gemm_calls;
    if (m < 0 || n < 0 || k < 0) return CUBLAS_STATUS_INVALID_VALUE;
    auto A = [&](int i, int p) -> T {
        if (ta == CUBLAS_OP_N) return a[size_t(p) * lda + i];
        T v = a[size_t(i) * lda + p];
        return ta == CUBLAS_OP_C ? cj(v) : v;
    };
    auto B = [&](int p, int j) -> T {
        if (tb == CUBLAS_OP_N) return b[size_t(j) * ldb + p];
        T v = b[size_t(p) * ldb + j];
        return tb == CUBLAS_OP_C ? cj(v) : v;
    };
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            T acc = T(0);
            for (int p = 0; p < k; ++p) acc += A(i, p) * B(p, j);
            T& out = c[size_t(j) * ldc + i];
            out = (*beta == T(0)) ? (*alpha) * acc : (*alpha) * acc + (*beta) * out;
        }
    return CUBLAS_STATUS_SUCCESS;
}
}  // namespace

extern "C" {
cublasStatus_t cublasSgemm_v2(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const float* al,
                              const float* a, int lda, const float* b, int ldb, const float* be, float* c, int ldc) {
    return gemm<float>(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
cublasStatus_t cublasDgemm_v2(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const double* al,
                              const double* a, int lda, const double* b, int ldb, const double* be, double* c, int ldc) {
    return gemm<double>(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
cublasStatus_t cublasCgemm_v2(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const cuComplex* al,
                              const cuComplex* a, int lda, const cuComplex* b, int ldb, const cuComplex* be, cuComplex* c, int ldc) {
    using Z = std::complex<float>;
    return gemm<Z>(ta, tb, m, n, k, reinterpret_cast<const Z*>(al), reinterpret_cast<const Z*>(a), lda, reinterpret_cast<const Z*>(b), ldb,
                   reinterpret_cast<const Z*>(be), reinterpret_cast<Z*>(c), ldc);
}
cublasStatus_t cublasZgemm_v2(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const cuDoubleComplex* al,
                              const cuDoubleComplex* a, int lda, const cuDoubleComplex* b, int ldb, const cuDoubleComplex* be,
                              cuDoubleComplex* c, int ldc) {
    using Z = std::complex<double>;
    return gemm<Z>(ta, tb, m, n, k, reinterpret_cast<const Z*>(al), reinterpret_cast<const Z*>(a), lda, reinterpret_cast<const Z*>(b), ldb,
                   reinterpret_cast<const Z*>(be), reinterpret_cast<Z*>(c), ldc);
}
}
