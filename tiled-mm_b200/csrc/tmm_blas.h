// Device-level GEMM layer: what replaces the reference's gpu_blas_api (cuBLAS forwarders,
// reference src/Tiled-MM/gpu_blas_api.hpp:194-252, driven from tiled_mm.cpp:181-268).
// All operands are DEVICE pointers, column-major; op in {N,T,C}; scalars are passed by value
// through host pointers (cuBLAS host pointer mode, as the reference uses it).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tmm {

enum DType : int { F32 = 0, F64 = 1, C32 = 2, C64 = 3 };

inline size_t dtype_size(int dt) { return dt == F32 ? 4 : (dt == F64 ? 8 : (dt == C32 ? 8 : 16)); }

// The TMA-fed kernels want A and B base pointers 16-byte aligned and lda/ldb * sizeof(elem) a multiple of 16.  The
// scheduler's device panels always satisfy this (it chooses the device pitch).  Other callers may pass anything cuBLAS
// accepts: an operand outside the contract is re-pitched once by the copy engine (FP64 types) or handled by the SIMT
// kernel (float types).  C has no alignment requirement beyond natural alignment.
// Returns cudaSuccess or the launch error; cudaErrorInvalidValue for bad trans / sizes.
cudaError_t device_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a,
                        int64_t lda, const void* b, int64_t ldb, const void* beta, void* c, int64_t ldc, cudaStream_t stream);

// C = beta * C over an m x n column-major block (beta == 0 writes zeros without reading C).
cudaError_t device_scale(int dtype, int64_t m, int64_t n, const void* beta, void* c, int64_t ldc, cudaStream_t stream);
// C += beta * S (m x n, column-major, separate leading dimensions)
cudaError_t device_add_scaled(int dtype, int64_t m, int64_t n, const void* beta, const void* s, int64_t lds, void* c, int64_t ldc, cudaStream_t stream);

// Number of kernels launched by this layer since process start (bench.py's gpu_launches).
uint64_t launch_count();

// per-dtype entry points (one translation unit each)
cudaError_t dgemm_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb,
                         double beta, double* c, int64_t ldc, cudaStream_t stream);
cudaError_t zgemm_launch(char ta, char tb, int m, int n, int k, const double* alpha2, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const double* beta2, void* c, int64_t ldc, cudaStream_t stream);
cudaError_t sgemm_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb,
                         float beta, float* c, int64_t ldc, cudaStream_t stream);
// float back ends: tcgen05/TMEM tensor-core kernel (terms = 3: FP32-accurate 3xTF32, 1: plain TF32) and the SIMT FFMA kernel
bool sgemm_tc_eligible(const void* a, int64_t lda, const void* b, int64_t ldb);
cudaError_t sgemm_tc_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb,
                            float beta, float* c, int64_t ldc, cudaStream_t stream, int terms);
cudaError_t sgemm_simt_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb,
                              float beta, float* c, int64_t ldc, cudaStream_t stream);
// process-wide math mode of the float GEMM: 3 = FP32-accurate on tensor cores (default), 1 = TF32, 0 = SIMT FFMA
int f32_math_mode();
void set_f32_math_mode(int mode);
cudaError_t cgemm_launch(char ta, char tb, int m, int n, int k, const float* alpha2, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const float* beta2, void* c, int64_t ldc, cudaStream_t stream);
// complex<float> back ends: the real embedding on the tcgen05 3xTF32 kernel (default; gemm_c32_tc.cu; returns cudaErrorMemoryAllocation when its
// scratch cannot be allocated, and the dispatcher then falls back) and the SIMT complex FMA kernel
cudaError_t cgemm_simt_launch(char ta, char tb, int m, int n, int k, const float* alpha2, const void* a, int64_t lda, const void* b, int64_t ldb,
                              const float* beta2, void* c, int64_t ldc, cudaStream_t stream);
cudaError_t cgemm_tc_launch(char ta, char tb, int m, int n, int k, const float* alpha2, const void* a, int64_t lda, const void* b, int64_t ldb,
                            const float* beta2, void* c, int64_t ldc, cudaStream_t stream);
// the three steps of cgemm_tc_launch for callers that prepare an operand once and multiply it many times (gemm_c32_tc.cu has the layouts)
cudaError_t cgemm_tc_embed_a(char ta, int m, int k, const float* alpha2, const void* a, int64_t lda, float* a2, int64_t pitch_a2, cudaStream_t stream);
cudaError_t cgemm_tc_split_b(char tb, int n, int k, const void* b, int64_t ldb, float* b2, int64_t pitch_b2, cudaStream_t stream);
cudaError_t cgemm_tc_prepared(char ta, char tb, int m, int n, int k, const float* a2, int64_t pitch_a2, const float* b2, int64_t pitch_b2,
                              const float* beta2, void* c, int64_t ldc, cudaStream_t stream);
// BF16 inputs, FP32 output / accumulation, on the tcgen05 TF32 path (bf16 is a subset of tf32: identical products) - gemm_bf16_tc.cu
cudaError_t bgemm_tc_launch(char ta, char tb, int m, int n, int k, float alpha, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb,
                            float beta, float* c, int64_t ldc, cudaStream_t stream);
// experimental native kind::f16 variant for k-contiguous operands (op(A) = T, op(B) = N); TMM_BF16_NATIVE=1
cudaError_t bgemm_tc_native_tn_launch(int m, int n, int k, float alpha, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, float beta, float* c,
                                      int64_t ldc, cudaStream_t stream);
// process-wide math mode of the complex<float> GEMM: 3 = FP32-accurate on tensor cores (default), 0 = SIMT
int c32_math_mode();
void set_c32_math_mode(int mode);

// experimental FP64 emulation on the int8 tensor cores (gemm_f64_i8.cu; TMM_F64_MATH=i8[:slices], default off): slice count of the process
// (0 = DMMA), and the launcher - any ld / alignment; cudaErrorMemoryAllocation = no scratch, the caller runs the DMMA kernel instead
int f64_i8_slices();
// the two halves of that product, for callers that multiply one sliced panel many times (the scheduler: a k-chunk of A meets several column
// stripes, the resident A meets every column block): S int8 slices of an operand's `rows` x k values, k contiguous -
// slice s, row i, k index l at q[(s * rows_pad + i) * pitch + l]; e[i] = power-of-two scale of row i
struct I8Slices {
    int8_t* q = nullptr;
    int* e = nullptr;
    int rows = 0, k = 0, slices = 0;
    int64_t rows_pad = 0, pitch = 0;
};
size_t i8_slices_layout(int rows, int k, int slices, int64_t* rows_pad, int64_t* pitch);  // bytes of q (e needs `rows` ints)
// element (row i, k index l) of the operand at x[i * stride_row + l * stride_k]
cudaError_t i8_slice_operand(const double* x, int64_t stride_row, int64_t stride_k, int rows, int k, const I8Slices& out, cudaStream_t stream);
// C[m x n] = alpha * A[a_row0 .. a_row0 + m) * B[b_row0 .. b_row0 + n)^T + beta * C on pre-sliced operands (same k, slice count and pitch)
cudaError_t i8_gemm_sliced(const I8Slices& a, int a_row0, int m, const I8Slices& b, int b_row0, int n, double alpha, double beta, double* c, int64_t ldc, cudaStream_t stream);
cudaError_t dgemm_i8_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb, double beta,
                            double* c, int64_t ldc, cudaStream_t stream, int slices);

// true-FP64 SIMT kernels: last resort for FP64 operands outside the TMA contract when no scratch can be allocated
cudaError_t dgemm_simt_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb,
                              double beta, double* c, int64_t ldc, cudaStream_t stream);
cudaError_t zgemm_simt_launch(char ta, char tb, int m, int n, int k, const double* alpha2, const void* a, int64_t lda, const void* b, int64_t ldb,
                              const double* beta2, void* c, int64_t ldc, cudaStream_t stream);

void count_launch();
int sm_count();
// stream-ordered scratch from the library's own per-device pool (keeps what it has grown to across calls; tmm_common.cu)
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t stream);
void scratch_free(void* p, cudaStream_t stream);  // nullptr is fine
void scratch_trim();                              // free blocks of the current device's pool go back to the driver
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
void* tensormap_encode_fn();

}  // namespace tmm
