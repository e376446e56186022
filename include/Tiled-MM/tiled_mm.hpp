// Drop-in counterpart of reference src/Tiled-MM/tiled_mm.hpp: gpu::gemm<Scalar> with the reference's exact
// signature and defaults (tiled_mm.hpp:69-79).  Column-major host pointers, op in {N,T,C} (case-insensitive),
// arbitrary leading dimensions, synchronous.  See include/tiled_mm_b200.h for the C ABI underneath.
#pragma once
#include "mm_handle.hpp"
#include "gpu_blas_api.hpp"

namespace gpu {

blas_api::OperationType get_blas_operation(char trans);

// C = alpha * op(A) * op(B) + beta * C; a, b, c are host pointers (same parameter order, types and defaults as the reference)
template <typename Scalar>
void gemm(mm_handle<Scalar>& handle, char trans_a, char trans_b, int m, int n, int k, Scalar alpha, Scalar* a, int ld_a, Scalar* b, int ld_b, Scalar beta,
          Scalar* c, int ld_c, bool pin_host_buffers = true, bool copy_c_back = true);

// 64-bit sizes (the reference's int offsets overflow at 2^31 elements, tiled_matrix.cpp:62-67)
template <typename Scalar>
void gemm64(mm_handle<Scalar>& handle, char trans_a, char trans_b, long long m, long long n, long long k, Scalar alpha, Scalar* a, long long ld_a,
            Scalar* b, long long ld_b, Scalar beta, Scalar* c, long long ld_c, bool pin_host_buffers = true, bool copy_c_back = true);

}  // namespace gpu
