// Drop-in counterpart of reference src/Tiled-MM/gpu_blas_api.hpp.
// The reference forwards blas_api::{s,d,c,z}gemm to cuBLAS (gpu_blas_api.hpp:194-252).  This library
// calls no vendor BLAS: the same four entry points run the hand-written sm_100a kernels through
// tmm_device_gemm (include/tiled_mm_b200.h).  OperationType keeps cuBLAS' numeric values (N=0,T=1,C=2).
#pragma once
#include "../tiled_mm_b200.h"
#include <complex>

namespace gpu {
namespace blas_api {

enum OperationType : int { OpNone = 0, OpTranspose = 1, OpConjugateTranspose = 2 };
using StatusType = int;          // TMM_OK / TMM_ERR_*
using HandleType = void*;        // a cudaStream_t: the only state a "handle" carried that matters here
using ComplexFloatType = std::complex<float>;
using ComplexDoubleType = std::complex<double>;

namespace operation {
constexpr OperationType None = OpNone;
constexpr OperationType Transpose = OpTranspose;
constexpr OperationType ConjugateTranspose = OpConjugateTranspose;
}  // namespace operation

namespace status {
constexpr StatusType Success = TMM_OK;
inline const char* get_string(StatusType) { return tmm_last_error(); }
}  // namespace status

// handle life cycle (reference gpu_blas_api.hpp:167-192 forwards these to cublasCreate / cublasDestroy / cublasSetStream): a handle is
// the stream its GEMMs run on, so creating one yields the default stream and binding a stream overwrites it
inline StatusType create(HandleType* handle) { *handle = nullptr; return status::Success; }
inline StatusType destroy(HandleType) { return status::Success; }
template <typename Stream>
inline StatusType set_stream(HandleType& handle, Stream stream) { handle = reinterpret_cast<HandleType>(stream); return status::Success; }

inline char op_char(OperationType op) { return op == OpNone ? 'N' : (op == OpTranspose ? 'T' : 'C'); }

// device pointers, column-major, host-pointer scalars: the cuBLAS v2 calling convention the reference uses
inline StatusType sgemm(HandleType stream, OperationType ta, OperationType tb, int m, int n, int k, const float* alpha, const float* a, int lda,
                        const float* b, int ldb, const float* beta, float* c, int ldc) {
    return tmm_device_gemm(TMM_F32, op_char(ta), op_char(tb), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream);
}
inline StatusType dgemm(HandleType stream, OperationType ta, OperationType tb, int m, int n, int k, const double* alpha, const double* a, int lda,
                        const double* b, int ldb, const double* beta, double* c, int ldc) {
    return tmm_device_gemm(TMM_F64, op_char(ta), op_char(tb), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream);
}
inline StatusType cgemm(HandleType stream, OperationType ta, OperationType tb, int m, int n, int k, const ComplexFloatType* alpha,
                        const ComplexFloatType* a, int lda, const ComplexFloatType* b, int ldb, const ComplexFloatType* beta, ComplexFloatType* c, int ldc) {
    return tmm_device_gemm(TMM_C32, op_char(ta), op_char(tb), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream);
}
inline StatusType zgemm(HandleType stream, OperationType ta, OperationType tb, int m, int n, int k, const ComplexDoubleType* alpha,
                        const ComplexDoubleType* a, int lda, const ComplexDoubleType* b, int ldb, const ComplexDoubleType* beta, ComplexDoubleType* c, int ldc) {
    return tmm_device_gemm(TMM_C64, op_char(ta), op_char(tb), m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, stream);
}

}  // namespace blas_api
}  // namespace gpu
