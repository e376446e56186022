// Operand preparation of the experimental FP64-emulating DGEMM (gemm_f64_i8.cu, TMM_F64_MATH=i8): every row of op(A) and every column
// of op(B) is scaled by a power of two and cut into signed 7-bit slices, stored as int8 matrices with k contiguous.
// Device code only - no runtime calls, no launch syntax - so that tests/test_slice_kernels.py compiles the very same source for the
// CPU (blockIdx / threadIdx shim) and compares it with the numpy restatement of tools/fp64_emulation_study.py.
//
//   x[i, l] = 2^e[i] * sum_s 2^-(P0 + BITS s) * q_s[i, l]  +  remainder,     |q_s| <= 2^(BITS-1) = 64,   |remainder| <= 2^(e[i] - P0 - BITS S) / 2
// e[i] = exponent of the largest magnitude of the row (|x| < 2^e), q_s = round-to-nearest of the running remainder: all steps are exact in
// FP64 (scaling by powers of two, rint, one subtraction of a representable number).  A row with an Inf or NaN is marked (e = NON_FINITE, zero
// slices) and comes out of the GEMM as NaN - a DGEMM would fill that row of C with Inf / NaN as well.
#pragma once
#include <cstdint>

namespace tmm {
namespace f64i8 {

constexpr int SLICE_BITS = 7;
constexpr int P0 = SLICE_BITS - 1;       // fractional bits of slice 0
constexpr int NO_DATA = -2000000000;     // e[] of a row that holds only zeros (cudaMemset pattern 0x88 = -2004318072 is below it)
constexpr int MAX_SLICES = 10;
constexpr int NON_FINITE = 5000;         // e[] of a row that holds an Inf or a NaN: its slices are zero, the GEMM epilogue writes NaN for it

// |x| < 2^e with e minimal for normal numbers (x = f * 2^e, 0.5 <= f < 1); zeros report NO_DATA; denormals count as < 2^-1022
__device__ __forceinline__ int exponent_above(double x) {
    const uint64_t u = (uint64_t)__double_as_longlong(x) & 0x7FFFFFFFFFFFFFFFull;
    if (u == 0) return NO_DATA;
    const int biased = (int)(u >> 52);
    if (biased == 2047) return NON_FINITE;
    return biased == 0 ? -1022 : biased - 1022;
}

// e[i] = max over l of exponent_above(x[i, l]); element (i, l) at x[i * stride_row + l * stride_k].  e[] preset below NO_DATA.
// grid.x covers the rows (one thread each), grid.y splits the k range.
static __global__ void __launch_bounds__(256) row_exponents(const double* __restrict__ x, int64_t stride_row, int64_t stride_k, int rows, int k, int k_per_block,
                                                            int* __restrict__ e) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int l0 = blockIdx.y * k_per_block, l1 = l0 + k_per_block < k ? l0 + k_per_block : k;
    int best = NO_DATA;
    const double* p = x + (int64_t)i * stride_row;
    for (int l = l0; l < l1; ++l) {
        const int ex = exponent_above(p[(int64_t)l * stride_k]);
        best = ex > best ? ex : best;
    }
    if (best > NO_DATA) atomicMax(&e[i], best);
}

// The same for operands whose k index is contiguous in memory (stride_k == 1: op(A) = T / C, op(B) = N): the kernel above would read with a
// stride of one row per lane.  Here blockIdx.y walks the rows (grid-stride) and the threads of a block read 4 consecutive k-values each,
// 1024 per step - coalesced; one atomicMax per warp (per thread where there are no warps: the CPU build of tests/test_slice_kernels.py).
static __global__ void __launch_bounds__(256) row_exponents_kmajor(const double* __restrict__ x, int64_t stride_row, int rows, int k, int* __restrict__ e) {
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {
        const double* p = x + (int64_t)i * stride_row;
        int best = NO_DATA;
        for (int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; l4 < k; l4 += gridDim.x * blockDim.x * 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (l4 + j < k) {
                    const int ex = exponent_above(p[l4 + j]);
                    best = ex > best ? ex : best;
                }
        }
#ifdef __CUDA_ARCH__
        best = __reduce_max_sync(0xffffffffu, best);
        if ((threadIdx.x & 31) == 0 && best > NO_DATA) atomicMax(&e[i], best);
#else
        if (best > NO_DATA) atomicMax(&e[i], best);
#endif
    }
}

// q_s[i, l] for s < slices, written as int8 at out[s * slice_stride + i * pitch + l] (bytes); four consecutive l per thread (one 32-bit store
// per slice); grid.x covers groups of four k-values, grid.y the rows (grid-stride).
static __global__ void __launch_bounds__(256) slice_rows(const double* __restrict__ x, int64_t stride_row, int64_t stride_k, int rows, int k, const int* __restrict__ e,
                                                         int8_t* __restrict__ out, int64_t pitch, int64_t slice_stride, int slices) {
    const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l4 >= k) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {
        const bool finite_row = e[i] < NON_FINITE;
        const int ei = (e[i] <= NO_DATA || !finite_row) ? 0 : e[i];
        // t = x 2^(P0 - e): slice s is q_s = rint(t_s), t_(s+1) = (t_s - q_s) 2^BITS - the same numbers as q_s = rint(r_s 2^(P0 + BITS s)),
        // r_(s+1) = r_s - q_s 2^-(P0 + BITS s), with one scalbn per element instead of three per slice (t - rint(t) and the scaling by a power
        // of two are exact); the slicing passes were 11 of the 41 ms of a device-resident 10000^3 call before (profiles/r2_ncu_kernels.md)
        double r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = (finite_row && l4 + j < k) ? scalbn(x[(int64_t)i * stride_row + (int64_t)(l4 + j) * stride_k], P0 - ei) : 0.0;
        for (int s = 0; s < slices; ++s) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double q = rint(r[j]);
                r[j] = (r[j] - q) * (double)(1 << SLICE_BITS);
                packed |= ((uint32_t)(uint8_t)(int8_t)(int)q) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(out + (int64_t)s * slice_stride + (int64_t)i * pitch + l4) = packed;
        }
    }
}

// The same for operands whose ROWS are contiguous in memory (stride_row == 1: op(A) = N, op(B) = T): consecutive threads take consecutive
// rows, so every read is a coalesced run along the rows; a thread handles K_PER_THREAD = 32 k-values and stores them as two 16-byte vectors
// per slice = one full 32-byte sector (the first version stored four separate words per 16 k-values: every warp store touched 32 sectors for
// 4 useful bytes each, and this pass was most of the 10 ms the passes took at 10000^3).  grid.x covers the rows, grid.y groups of 32 k-values
// (grid-stride).  pitch and slice_stride are multiples of 32.
constexpr int K_PER_THREAD = 32;
struct alignas(16) Q16 { uint32_t w[4]; };
static __global__ void __launch_bounds__(256) slice_rows_contiguous(const double* __restrict__ x, int64_t stride_k, int rows, int k, const int* __restrict__ e,
                                                                    int8_t* __restrict__ out, int64_t pitch, int64_t slice_stride, int slices) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const bool finite_row = e[i] < NON_FINITE;
    const int ei = (e[i] <= NO_DATA || !finite_row) ? 0 : e[i];
    for (int l0 = blockIdx.y * K_PER_THREAD; l0 < k; l0 += gridDim.y * K_PER_THREAD) {
        double r[K_PER_THREAD];
#pragma unroll
        for (int j = 0; j < K_PER_THREAD; ++j) r[j] = (finite_row && l0 + j < k) ? scalbn(x[(int64_t)(l0 + j) * stride_k + i], P0 - ei) : 0.0;
        for (int s = 0; s < slices; ++s) {
            Q16 q[K_PER_THREAD / 16];
#pragma unroll
            for (int v = 0; v < K_PER_THREAD / 16; ++v) q[v].w[0] = q[v].w[1] = q[v].w[2] = q[v].w[3] = 0;
#pragma unroll
            for (int j = 0; j < K_PER_THREAD; ++j) {
                const double t = rint(r[j]);
                r[j] = (r[j] - t) * (double)(1 << SLICE_BITS);
                q[j >> 4].w[(j >> 2) & 3] |= ((uint32_t)(uint8_t)(int8_t)(int)t) << (8 * (j & 3));
            }
            Q16* dst = reinterpret_cast<Q16*>(out + (int64_t)s * slice_stride + (int64_t)i * pitch + l0);
#pragma unroll
            for (int v = 0; v < K_PER_THREAD / 16; ++v) dst[v] = q[v];
        }
    }
}

}  // namespace f64i8
}  // namespace tmm
