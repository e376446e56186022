// CGEMM on the tcgen05 tensor cores:  C = alpha * op(A) * op(B) + beta * C for complex<float>, column-major, device operands.
// Replaces blas_api::cgemm (reference gpu_blas_api.hpp:233-251, called from tiled_mm.cpp:222-245).
//
// A complex product is a real product of twice the size.  With every complex number written as its (re, im) pair - which is how
// interleaved storage already looks when read as floats -
//     C~ = A' * B'          C~ : (2m x n)   the float view of C, rows (2i, 2i+1) = (re, im) of row i
//                           B' : (2k x n)   the float view of op(B), rows (2l, 2l+1) = (re, im) of row l
//                           A' : (2m x 2k)  column 2l   = (re, im) pairs of  w(:, l)      [ w = alpha * op(A) ]
//                                           column 2l+1 = (re, im) pairs of  i * w(:, l) = (-im, re)
// because  w * b = w * re(b) + (i w) * im(b).  The real product has 2 * (2m) * n * (2k) = 8mnk flops - exactly the complex flop
// count - and runs on the FP32-accurate 3xTF32 kernel of gemm_f32_tc.cu unchanged (same TMA / TMEM / windowed-promotion
// pipeline, same accuracy: small integers stay exact).  What this file adds is the operand preparation, one elementwise pass
// per operand - into buffers of the caller (the scheduler prepares every piece of a panel once, see below) or, for a self-contained call,
// into stream-ordered scratch of the library's pool:
//   A  always (it has to be duplicated as w and i*w; alpha and the conjugation of op 'C' are folded in for free):
//        2 x |A| bytes written; k-contiguous for op T/C, m-contiguous for op N - the SGEMM kernel takes either orientation
//   B  op N: nothing - the stored matrix read as floats IS B' (zero copy) when its base and pitch meet the TMA contract
//      op T/C: de-interleave into (n x 2k), n-contiguous (re column, +-im column)
//   C  beta real: nothing (the SGEMM epilogue applies it to re and im alike); beta complex: one scaling pass first
// The passes move O(|A| + |B|) bytes through HBM against O(mnk) tensor work: < 5 % at the scheduler's launch shapes.
// If the scratch cannot be allocated the SIMT kernel takes the call.
//
// STATUS: the default complex<float> path since round 2.  First hardware run (profiles/r2_experimental_first_run.txt): bit-exact on integer data
// for all nine op pairs, 8192^3 in 34.3 ms = 128 TF (8mnk) against 96.8 ms for the SIMT kernel and 62.3 ms for cuBLAS CGEMM.
// TMM_C32_MATH=simt / tmm_set_c32_math(TMM_CMATH_SIMT) selects the complex-FMA kernel (true FP32 arithmetic in every product).
#include "tmm_blas.h"
#include "tmm_prepass.cuh"  // embed_a_n, embed_a_t, split_b_t: device code only, also compiled for the CPU by tests/test_prepass_kernels.py

#include <cstdint>

namespace tmm {
namespace c32tc {

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }
static inline dim3 pass_grid(int contiguous, int columns) {
    return dim3((unsigned)((contiguous + 255) / 256), (unsigned)(columns < 1 ? 1 : (columns > 32768 ? 32768 : columns)));
}

}  // namespace c32tc

// ---- the three steps, for callers that prepare an operand once and multiply it many times (the scheduler: a k-chunk of A meets every column
// stripe, the resident A meets every phase-2 column block - csrc/tmm_context.cu run_resident) ----------------------------------------------

// A' of alpha * op(A) for an m x k block of op(A).  ta == 'N': a2 is (2m x 2k), m-contiguous, pitch_a2 floats per column (even);
// otherwise a2 holds A'^T = (2k x 2m), k-contiguous.  A k sub-range [p0, p0 + kc) of a larger A' starts at a2 + 2 * p0 * pitch_a2 ('N') or
// a2 + 2 * p0 (otherwise): chunks embedded one by one into the same buffer add up to the A' of the whole panel.
cudaError_t cgemm_tc_embed_a(char ta, int m, int k, const float* al, const void* a, int64_t lda, float* a2, int64_t pitch_a2, cudaStream_t st) {
    using namespace c32tc;
    if (m <= 0 || k <= 0) return cudaSuccess;
    const float2 alpha = make_float2(al[0], al[1]);
    if (ta == 'N') embed_a_n<<<pass_grid(m, k), 256, 0, st>>>(static_cast<const float2*>(a), lda, m, k, alpha, reinterpret_cast<float2*>(a2), pitch_a2 / 2);
    else embed_a_t<<<pass_grid(k, m), 256, 0, st>>>(static_cast<const float2*>(a), lda, k, m, alpha, ta == 'C' ? 1 : 0, reinterpret_cast<float2*>(a2), pitch_a2 / 2);
    count_launch();
    return cudaGetLastError();
}

// B'^T of op(B) for tb == 'T' / 'C': stored n x k complex (n contiguous) -> (n x 2k) floats, n-contiguous, pitch_b2 floats per column.
// (tb == 'N' needs no pass: the stored matrix read as floats is B', pitch 2 * ldb.)
cudaError_t cgemm_tc_split_b(char tb, int n, int k, const void* b, int64_t ldb, float* b2, int64_t pitch_b2, cudaStream_t st) {
    using namespace c32tc;
    if (n <= 0 || k <= 0) return cudaSuccess;
    split_b_t<<<pass_grid(n, k), 256, 0, st>>>(static_cast<const float2*>(b), ldb, n, k, tb == 'C' ? 1 : 0, b2, pitch_b2);
    count_launch();
    return cudaGetLastError();
}

// C = A' B' + beta * C on prepared operands: the real (2m x n) = (2m x 2k)(2k x n) product on the FP32-accurate 3xTF32 kernel.
// cudaErrorInvalidValue when a2 / b2 do not meet the TMA contract (16-byte aligned base, pitch a multiple of 4 floats).
cudaError_t cgemm_tc_prepared(char ta, char tb, int m, int n, int k, const float* a2, int64_t pitch_a2, const float* b2, int64_t pitch_b2,
                              const float* be, void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0 || k <= 0) return cudaSuccess;
    if (!sgemm_tc_eligible(a2, pitch_a2, b2, pitch_b2)) return cudaErrorInvalidValue;
    float beta_r = be[0];
    cudaError_t e = cudaSuccess;
    if (be[1] != 0.f) {  // complex beta: C <- beta * C first, then accumulate with 1
        e = device_scale(C32, m, n, be, c, ldc, st);
        beta_r = 1.f;
    }
    if (e == cudaSuccess) e = sgemm_tc_launch(ta == 'N' ? 'N' : 'T', tb == 'N' ? 'N' : 'T', 2 * m, n, 2 * k, 1.f, a2, pitch_a2, b2, pitch_b2, beta_r, static_cast<float*>(c), 2 * ldc, st, 3);
    return e;
}

// One self-contained call (blas_api::cgemm, device-pointer operands): the passes write into stream-ordered scratch.
// Returns cudaErrorMemoryAllocation when the scratch cannot be had (caller falls back to SIMT); any other error is final.
cudaError_t cgemm_tc_launch(char ta, char tb, int m, int n, int k, const float* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                            const float* be, void* c, int64_t ldc, cudaStream_t st) {
    using namespace c32tc;
    if (m <= 0 || n <= 0 || k <= 0) return cudaSuccess;
    if (m > INT32_MAX / 2 || k > INT32_MAX / 2) return cudaErrorMemoryAllocation;  // the doubled extents must fit the SGEMM's int sizes
    const bool a_n = ta == 'N', b_n = tb == 'N';

    // ---- A' ----
    const int64_t pitch_a = a_n ? round_up(2 * (int64_t)m, 32) : round_up(2 * (int64_t)k, 32);  // floats; 128-byte columns
    const int64_t cols_a = a_n ? 2 * (int64_t)k : 2 * (int64_t)m;
    float* a2 = nullptr;
    cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&a2), (size_t)pitch_a * (size_t)cols_a * sizeof(float), st);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaErrorMemoryAllocation; }
    e = cgemm_tc_embed_a(ta, m, k, al, a, lda, a2, pitch_a, st);

    // ---- B' ----
    const float* b2 = nullptr;
    int64_t ldb2 = 0;
    float* b_scratch = nullptr;
    if (b_n && (reinterpret_cast<uintptr_t>(b) & 15) == 0 && (ldb & 1) == 0) {
        b2 = static_cast<const float*>(b);  // zero copy: (2k x n) floats with pitch 2 * ldb
        ldb2 = 2 * ldb;
    } else if (e == cudaSuccess) {
        const int64_t pitch_b = b_n ? round_up(2 * (int64_t)k, 32) : round_up((int64_t)n, 32);
        const int64_t cols_b = b_n ? (int64_t)n : 2 * (int64_t)k;
        e = scratch_alloc(reinterpret_cast<void**>(&b_scratch), (size_t)pitch_b * (size_t)cols_b * sizeof(float), st);
        if (e != cudaSuccess) { cudaGetLastError(); scratch_free(a2, st); return cudaErrorMemoryAllocation; }
        if (b_n)  // same values, legal pitch: the copy engine re-pitches
            e = cudaMemcpy2DAsync(b_scratch, (size_t)pitch_b * sizeof(float), b, (size_t)ldb * sizeof(float2), (size_t)k * sizeof(float2), (size_t)n,
                                  cudaMemcpyDeviceToDevice, st);
        else e = cgemm_tc_split_b(tb, n, k, b, ldb, b_scratch, pitch_b, st);
        b2 = b_scratch;
        ldb2 = pitch_b;
    }
    if (e == cudaSuccess) e = cgemm_tc_prepared(ta, tb, m, n, k, a2, pitch_a, b2, ldb2, be, c, ldc, st);
    scratch_free(a2, st);
    scratch_free(b_scratch, st);
    return e;
}

}  // namespace tmm
