// Elementwise operand-preparation kernels of the tcgen05 CGEMM embedding (gemm_c32_tc.cu) and of the BF16 entry point (gemm_bf16_tc.cu).
// Device code only - no runtime calls, no launch syntax - so that the very same source is also compiled for the CPU (with a ten-line
// shim for blockIdx / threadIdx / float2) by tests/test_prepass_kernels.py, which runs every "thread" in a loop and compares the
// result with the numpy restatement of the embedding: these kernels were written after the round's GPU budget was spent.
#pragma once
#include <cstdint>

namespace tmm {
namespace c32tc {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// op N: stored m x k complex (m contiguous).  out: (2m x 2k) floats, pitch floats per column (even), written as float2 pairs.
static __global__ void __launch_bounds__(256) embed_a_n(const float2* __restrict__ a, int64_t lda, int m, int k, float2 alpha, float2* __restrict__ out2, int64_t pitch2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int l = blockIdx.y; l < k; l += gridDim.y) {
        const float2 w = cmul(alpha, a[(int64_t)l * lda + i]);
        out2[(int64_t)(2 * l) * pitch2 + i] = w;                           // column 2l   : ( re,  im)
        out2[(int64_t)(2 * l + 1) * pitch2 + i] = make_float2(-w.y, w.x);  // column 2l+1 : (-im,  re)  = i * w
    }
}

// op T / C: stored k x m complex (k contiguous), element (l, i).  out = A'^T: (2k x 2m) floats, k-contiguous:
//   column 2i   rows (2l, 2l+1) = ( re, -im)        column 2i+1 rows (2l, 2l+1) = ( im,  re)         of w = alpha * op(a(l, i))
static __global__ void __launch_bounds__(256) embed_a_t(const float2* __restrict__ a, int64_t lda, int k, int m, float2 alpha, int conj, float2* __restrict__ out2,
                                                  int64_t pitch2) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= k) return;
    for (int i = blockIdx.y; i < m; i += gridDim.y) {
        float2 v = a[(int64_t)i * lda + l];
        if (conj) v.y = -v.y;
        const float2 w = cmul(alpha, v);
        out2[(int64_t)(2 * i) * pitch2 + l] = make_float2(w.x, -w.y);
        out2[(int64_t)(2 * i + 1) * pitch2 + l] = make_float2(w.y, w.x);
    }
}

// op T / C of B: stored n x k complex (n contiguous), element (j, l).  out = B'^T: (n x 2k) floats, n-contiguous:
//   column 2l = re(b(:, l)),  column 2l+1 = +-im(b(:, l))
static __global__ void __launch_bounds__(256) split_b_t(const float2* __restrict__ b, int64_t ldb, int n, int k, int conj, float* __restrict__ out, int64_t pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int l = blockIdx.y; l < k; l += gridDim.y) {
        const float2 v = b[(int64_t)l * ldb + j];
        out[(int64_t)(2 * l) * pitch + j] = v.x;
        out[(int64_t)(2 * l + 1) * pitch + j] = conj ? -v.y : v.y;
    }
}

// stored rows x cols bf16 (ld_in elements per column) -> fp32 with pitch floats per column
static __global__ void __launch_bounds__(256) widen(const uint16_t* __restrict__ in, int64_t ld_in, int rows, int cols, float* __restrict__ out, int64_t pitch) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int c = blockIdx.y; c < cols; c += gridDim.y) out[(int64_t)c * pitch + r] = __uint_as_float((uint32_t)in[(int64_t)c * ld_in + r] << 16);
}

}  // namespace c32tc

namespace f32tc {

// x -> (hi, lo): hi = x rounded to TF32 (nearest, ties away), lo = (x - hi) rounded to TF32; Inf/NaN keep lo = 0
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    const uint32_t u = __float_as_uint(x);
    uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
    float r = x - __uint_as_float(h);
    if ((u & 0x7F800000u) == 0x7F800000u) { h = (u & 0x007FFFFFu) ? 0x7FC00000u : u; r = 0.f; }
    else if ((h & 0x7F800000u) == 0x7F800000u) {  // a finite x within 2^-12 of FLT_MAX would round UP to Inf: take the largest finite TF32 instead, lo carries the rest
        h = (u & 0x80000000u) | 0x7F7FE000u;
        r = x - __uint_as_float(h);
    }
    hi = __uint_as_float(h);
    lo = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
}

// The same split without the special cases: exact for every x with |x| < 2^127 (rounding to TF32 then stays finite).
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}

// Four neighbouring elements, as the split warps of sgemm_tc_kernel take them (one 16-byte chunk): one comparison per element decides whether
// the chunk takes the five-instruction path; a chunk holding an Inf, a NaN or a value of 2^127 or more goes through split_tf32.  (The split
// warps sit between TMA and the tensor core on a three-stage ring: with the special cases applied to every element - ~19 instructions - the
// FP32-accurate mode ran at 96 TF instead of 155 TF at 8192^3, profiles/r2_bisect_tc.txt.)
__device__ __forceinline__ void split_tf32_x4(const float (&x)[4], float (&hi)[4], float (&lo)[4]) {
    constexpr float LIM = 1.7014118e38f;  // 2^127; a NaN compares false
    const bool plain = (fabsf(x[0]) < LIM) & (fabsf(x[1]) < LIM) & (fabsf(x[2]) < LIM) & (fabsf(x[3]) < LIM);
    if (plain) {
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32_fast(x[i], hi[i], lo[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(x[i], hi[i], lo[i]);
    }
}

// Variant for hi = raw bits: the hardware reads trunc(x) (top 19 bits), so lo = x - trunc(x) (exact), rounded to TF32.
// |lo| < 2^-10 |x| instead of 2^-11 |x| (one bit less accurate than the round-to-nearest split) but the tile is not rewritten.
__device__ __forceinline__ float lo_of_truncated(float x) {
    const uint32_t u = __float_as_uint(x);
    if ((u & 0x7F800000u) == 0x7F800000u) return 0.f;  // Inf / NaN travel in hi alone
    const float r = x - __uint_as_float(u & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
}

}  // namespace f32tc
}  // namespace tmm
