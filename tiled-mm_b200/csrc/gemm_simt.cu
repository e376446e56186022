// Generic shared-memory-tiled SIMT GEMM (FFMA / complex FMA), all dtypes, N/T/C, any ld.
// complex<float> default path (reference blas_api::cgemm, gpu_blas_api.hpp:233-251; the tcgen05 embedding of gemm_c32_tc.cu
// is opt-in until it has run on hardware) and the float path for operands that do
// not meet the TMA alignment contract or when TMM_F32_MATH=simt is selected: true FP32 FFMA arithmetic like cuBLAS'
// default math mode (the reference never sets a TF32 math mode, gpu_blas_handle.hpp:11-17).
// Aligned float operands run on tcgen05/TMEM (gemm_f32_tc.cu, 3xTF32); see DESIGN.md.
#include "tmm_blas.h"

#include <cuComplex.h>

namespace tmm {
namespace simt {

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float conj(float a) { return a; }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double conj(double a) { return a; }
};
template <> struct Num<cuFloatComplex> {
    using T = cuFloatComplex;
    static __device__ __forceinline__ T zero() { return make_cuFloatComplex(0.f, 0.f); }
    static __device__ __forceinline__ T fma(T a, T b, T c) { return cuCfmaf(a, b, c); }
    static __device__ __forceinline__ T mul(T a, T b) { return cuCmulf(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return cuCaddf(a, b); }
    static __device__ __forceinline__ T conj(T a) { return cuConjf(a); }
};
template <> struct Num<cuDoubleComplex> {
    using T = cuDoubleComplex;
    static __device__ __forceinline__ T zero() { return make_cuDoubleComplex(0.0, 0.0); }
    static __device__ __forceinline__ T fma(T a, T b, T c) { return cuCfma(a, b, c); }
    static __device__ __forceinline__ T mul(T a, T b) { return cuCmul(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return cuCadd(a, b); }
    static __device__ __forceinline__ T conj(T a) { return cuConj(a); }
};

constexpr int TS = 64, TK = 16, TPB = 256, RT = 4;  // 64x64 CTA tile, 4x4 per thread

template <typename T>
__global__ void __launch_bounds__(TPB) gemm_kernel(int ta, int tb, int m, int n, int k, T alpha, const T* __restrict__ a, int64_t lda,
                                                    const T* __restrict__ b, int64_t ldb, T beta, int read_c, T* c, int64_t ldc) {
    __shared__ T sa[TK][TS + 1];
    __shared__ T sb[TK][TS + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.x * TS, n0 = blockIdx.y * TS;
    T acc[RT][RT];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < RT; ++j) acc[i][j] = Num<T>::zero();

    for (int k0 = 0; k0 < k; k0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += TPB) {
            int mm, kk;
            if (ta == 0) { mm = e % TS; kk = e / TS; } else { kk = e % TK; mm = e / TK; }
            T v = Num<T>::zero();
            if (m0 + mm < m && k0 + kk < k) {
                v = ta == 0 ? a[(int64_t)(k0 + kk) * lda + m0 + mm] : a[(int64_t)(m0 + mm) * lda + k0 + kk];
                if (ta == 2) v = Num<T>::conj(v);
            }
            sa[kk][mm] = v;
            int nn;
            if (tb == 0) { kk = e % TK; nn = e / TK; } else { nn = e % TS; kk = e / TS; }
            v = Num<T>::zero();
            if (n0 + nn < n && k0 + kk < k) {
                v = tb == 0 ? b[(int64_t)(n0 + nn) * ldb + k0 + kk] : b[(int64_t)(k0 + kk) * ldb + n0 + nn];
                if (tb == 2) v = Num<T>::conj(v);
            }
            sb[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            T av[RT], bv[RT];
#pragma unroll
            for (int i = 0; i < RT; ++i) av[i] = sa[kk][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < RT; ++j) bv[j] = sb[kk][ty + 16 * j];
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < RT; ++j) acc[i][j] = Num<T>::fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < RT; ++j) {
        const int col = n0 + ty + 16 * j;
        if (col >= n) continue;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int row = m0 + tx + 16 * i;
            if (row >= m) continue;
            T* p = c + (int64_t)col * ldc + row;
            T v = Num<T>::mul(alpha, acc[i][j]);
            if (read_c) v = Num<T>::add(v, Num<T>::mul(beta, *p));
            *p = v;
        }
    }
}

static int opcode(char t) { return t == 'N' ? 0 : (t == 'T' ? 1 : 2); }

template <typename T>
static cudaError_t launch(char ta, char tb, int m, int n, int k, T alpha, const void* a, int64_t lda, const void* b, int64_t ldb, T beta,
                          bool read_c, void* c, int64_t ldc, cudaStream_t st) {
    dim3 grid((m + TS - 1) / TS, (n + TS - 1) / TS);
    if (grid.y > 65535) return cudaErrorInvalidValue;
    gemm_kernel<T><<<grid, TPB, 0, st>>>(opcode(ta), opcode(tb), m, n, k, alpha, static_cast<const T*>(a), lda, static_cast<const T*>(b), ldb, beta,
                                         read_c ? 1 : 0, static_cast<T*>(c), ldc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace simt

cudaError_t sgemm_simt_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta,
                              float* c, int64_t ldc, cudaStream_t st) {
    return simt::launch<float>(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, beta != 0.f, c, ldc, st);
}
cudaError_t cgemm_simt_launch(char ta, char tb, int m, int n, int k, const float* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const float* be, void* c, int64_t ldc, cudaStream_t st) {
    return simt::launch<cuFloatComplex>(ta, tb, m, n, k, make_cuFloatComplex(al[0], al[1]), a, lda, b, ldb, make_cuFloatComplex(be[0], be[1]),
                                        be[0] != 0.f || be[1] != 0.f, c, ldc, st);
}
cudaError_t dgemm_simt_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb, double beta,
                              double* c, int64_t ldc, cudaStream_t st) {
    return simt::launch<double>(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, beta != 0.0, c, ldc, st);
}
cudaError_t zgemm_simt_launch(char ta, char tb, int m, int n, int k, const double* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                              const double* be, void* c, int64_t ldc, cudaStream_t st) {
    return simt::launch<cuDoubleComplex>(ta, tb, m, n, k, make_cuDoubleComplex(al[0], al[1]), a, lda, b, ldb, make_cuDoubleComplex(be[0], be[1]),
                                         be[0] != 0.0 || be[1] != 0.0, c, ldc, st);
}
#ifndef TMM_HAVE_ZGEMM_DMMA
cudaError_t zgemm_launch(char ta, char tb, int m, int n, int k, const double* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const double* be, void* c, int64_t ldc, cudaStream_t st) {
    return simt::launch<cuDoubleComplex>(ta, tb, m, n, k, make_cuDoubleComplex(al[0], al[1]), a, lda, b, ldb, make_cuDoubleComplex(be[0], be[1]),
                                         be[0] != 0.0 || be[1] != 0.0, c, ldc, st);
}
#endif
}  // namespace tmm
