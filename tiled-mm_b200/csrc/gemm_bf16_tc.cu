// BF16-input GEMM on the tcgen05 tensor cores:  C(fp32) = alpha * op(A)(bf16) * op(B)(bf16) + beta * C, column-major, device operands.
// Additive entry point (north_star names BF16 as a kernel family; the reference API has no such type - SURVEY 8f-4).
//
// A bf16 number is a TF32 number with two trailing zero mantissa bits, so a kind::tf32 MMA over the widened operands forms exactly
// the products a kind::f16 (BF16) MMA would form, and accumulates them in FP32 the same way.  This first version therefore widens
// each operand once (bf16 -> fp32 is a 16-bit shift; one elementwise pass into stream-ordered scratch) and runs the hardware-
// validated single-term TF32 mode of gemm_f32_tc.cu (all four transposes, windowed FP32 promotion): bit-for-bit the arithmetic
// of a BF16 tensor-core GEMM at the TF32 issue rate.  A native kind::f16 pipeline (64-element K tiles, no widening pass, twice
// the rate) is the follow-up once it can be measured.
//
// STATUS: ran on hardware in round 2 (profiles/r2_experimental_first_run.txt): exact on integer-valued bf16 data for all four op pairs, 2e-6 on
// random data - tests/test_experimental_gpu.py::test_bf16_input_gemm, test_bf16_native_kind_f16_tn (the native path, TMM_BF16_NATIVE=1).
#include "tmm_blas.h"
#include "tmm_prepass.cuh"

#include <cstdint>
#include <cstdlib>

namespace tmm {
namespace bf16tc {

using c32tc::widen;  // tmm_prepass.cuh

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

}  // namespace bf16tc

cudaError_t bgemm_tc_launch(char ta, char tb, int m, int n, int k, float alpha, const void* a, int64_t lda, const void* b, int64_t ldb, float beta,
                            float* c, int64_t ldc, cudaStream_t st) {
    using namespace bf16tc;
    if (m <= 0 || n <= 0) return cudaSuccess;
    if (k <= 0) return device_scale(F32, m, n, &beta, c, ldc, st);
    {
        const char* nv = getenv("TMM_BF16_NATIVE");  // read per call (tests switch it)
        const bool native = nv && nv[0] == '1';
        if (native && ta != 'N' && tb == 'N' && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 && (lda & 7) == 0 && (ldb & 7) == 0)
            return bgemm_tc_native_tn_launch(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, st);
    }
    const int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    const int64_t pa = round_up(ar, 32), pb = round_up(br, 32);  // 128-byte columns: TMA-legal for any input ld
    float *a32 = nullptr, *b32 = nullptr;
    cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&a32), (size_t)pa * ac * sizeof(float), st);
    if (e != cudaSuccess) return e;
    e = scratch_alloc(reinterpret_cast<void**>(&b32), (size_t)pb * bc * sizeof(float), st);
    if (e != cudaSuccess) { scratch_free(a32, st); return e; }
    auto grid = [](int rows, int cols) { return dim3((unsigned)((rows + 255) / 256), (unsigned)(cols > 32768 ? 32768 : cols)); };
    widen<<<grid(ar, ac), 256, 0, st>>>(static_cast<const uint16_t*>(a), lda, ar, ac, a32, pa);
    widen<<<grid(br, bc), 256, 0, st>>>(static_cast<const uint16_t*>(b), ldb, br, bc, b32, pb);
    count_launch(); count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = sgemm_tc_launch(ta == 'N' ? 'N' : 'T', tb == 'N' ? 'N' : 'T', m, n, k, alpha, a32, pa, b32, pb, beta, c, ldc, st, /*terms=*/1);
    scratch_free(a32, st);
    scratch_free(b32, st);
    return e;
}

}  // namespace tmm
