// Opt-in FP64 math mode (TMM_F64_MATH=i8[:slices]); the default FP64 path is the DMMA kernel of gemm_f64.cu, as north_star specifies.
// First hardware run in round 2 (profiles/r2_f64_i8_first_run.txt): bit-exact against cuBLAS on integer data for all op pairs, within the FP64 parity
// bound on random data, 10000^3 device-resident at 47.6 TFLOP/s with 7 slices (1.34 x cuBLAS DGEMM) and 39.0 with 8.
//
// FP64-accurate DGEMM on the INTEGER tensor cores (Ozaki-type error-free slicing):  C = alpha * op(A) * op(B) + beta * C.
// Why: the DMMA kernel of gemm_f64.cu sits at 96 % of the chip's FP64 rate (36.9 TF), and at 10000^3 that rate - not PCIe (46 TF) - bounds
// the host-to-host call; tcgen05 has no FP64 kind but kind::i8 runs at ~120 x the FP64 rate.  Every row of op(A) / column of op(B) is scaled
// by a power of two and cut into S signed 7-bit slices (tmm_slice.cuh); slice products are EXACT int8 x int8 -> int32 GEMMs, and
//     C = 2^(ea[i] + eb[j]) * sum_{g < S} 2^-(2 P0 + 7 g) * ( sum_{s + t = g} Qa_s Qb_t^T )[i, j]      (terms with s + t >= S dropped)
// Accuracy (tools/fp64_emulation_study.py): S = 7 -> 2e-16, S = 8 -> the error of native FP64, relative to k max|A| max|B|; integer data exact.
// Non-finite operands: a row of op(A) / column of op(B) holding an Inf or NaN is detected by the exponent pass and its row / column of C is NaN.
//
// One kernel does the S (S + 1) / 2 slice GEMMs of a tile back to back - for a fixed g they share their scale, so they are ONE integer GEMM
// over the concatenated k range [Qa_0 | .. | Qa_g] [Qb_g | .. | Qb_0]^T accumulated in one int32 TMEM window - and keeps the FP64 sums in
// registers: structure and barrier protocol of the one-term path of gemm_f32_tc.cu (persistent CTA per SM; warp 0 TMA producer, warp 1
// MMA issuer, 6-stage ring of [128 x 128 B] K-major SWIZZLE_128B tiles, four rotating TMEM accumulators), with
//   * both operands always k-contiguous int8 (the slicing pass writes them that way whatever op(A) / op(B) are): byte geometry identical to
//     the validated K-major TF32 tile (128-byte rows, 32 bytes per UMMA_K step), kind::i8 with UMMA_K = 32
//   * TWO accumulate/epilogue warpgroups (warps 4-7: columns 0-63, warps 8-11: columns 64-127): int32 -> double, scaled by the group's power
//     of two (exact), added into 64 FP64 registers per thread; tile end: scalbn by ea[i] + eb[j], alpha / beta, plain stores (any ldc)
//   * int32 never overflows: |q| <= 64, so a window of W k-blocks adds at most W * 128 * 4096 < 2^31 for W <= 4095; longer groups are cut
// What bounds it (round 2): 627 cycles per k-block against 271 of integer MMA work - the six 32 KB stages in flight per SM over a ~1.9 us load
// latency are exactly the measured 15 TB/s of L2 -> SM traffic.  A variant on 2 x 2 clusters that TMA-multicasts operand halves (half the L2
// requests, same bytes landing per SM) ran at the same pace per CTA with fewer CTAs resident (33 clusters) and was removed
// (profiles/r2_i8_cluster_experiment.txt); more MACs per landed byte needs 256-wide tiles, i.e. another home for the FP64 sums.
#include "tmm_blas.h"
#include "tmm_tc.cuh"
#include "tmm_slice.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace tmm {
namespace f64i8 {

constexpr int BM = 128, BN = 128, BKB = 128, UMMA_KB = 32;  // k-block / MMA k-step in int8 elements (= bytes)
constexpr int STAGES = 6;
constexpr int OPERAND_BYTES = BM * BKB;          // 16 KB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;   // A | B
constexpr int ACC_BUFS = 4, TMEM_COLS = ACC_BUFS * BN;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_TMEM = 2, WARP_EPI0 = 4, EPI_WARPS = 8;
constexpr int THREADS = 512;
constexpr int REGS_CONTROL = 40, REGS_EPILOGUE = 208;  // 40 + 208 + 208 + 40 <= 512
constexpr int GROUP_COLS = 8;                    // raster width (measured 2 / 4 / 8 / 16: 41.3 / 40.6 / 40.4 / 41.3 ms at 10000^3, profiles/r2_i8_cluster_experiment.txt)
constexpr int MAX_WINDOW = 2048;                 // k-blocks per int32 accumulation window (overflow bound: 4095)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(2 * REGS_CONTROL + 2 * REGS_EPILOGUE <= 512, "setmaxnreg budget");

struct Params {
    double* c;
    int64_t ldc;
    int m, n, k;
    double alpha, beta;
    int read_c;
    int tiles_m, tiles_n;
    int slices;            // S
    int m_pad, n_pad;      // rows between consecutive slices in the A / B slice stacks (multiples of 128)
    int row0_a, row0_b;    // first row of this product inside the stacks (a stripe of a panel that was sliced once: see i8_gemm_sliced)
    const int* ea;         // [m] row exponents of op(A)
    const int* eb;         // [n] column exponents of op(B)
    uint64_t desc;         // K-major SWIZZLE_128B descriptor template
    uint32_t idesc;
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    const int per_group = GROUP_COLS * tiles_m;
    const int group = tile / per_group;
    const int r = tile - group * per_group;
    const int first = group * GROUP_COLS;
    const int width = min(GROUP_COLS, tiles_n - first);
    tm = r / width;
    tn = first + (r - tm * width);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
dgemm_i8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);  // TMA landed        -> MMA issuer
    uint64_t* empty_bar = full_bar + STAGES;                                         // MMAs retired      -> TMA producer
    uint64_t* acc_full_bar = empty_bar + STAGES;                                     // window complete   -> accumulate warps
    uint64_t* acc_empty_bar = acc_full_bar + ACC_BUFS;                               // window drained    -> MMA issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + ACC_BUFS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int b = 0; b < ACC_BUFS; ++b) {
            ptx::mbar_init(&acc_full_bar[b], 1);
            ptx::mbar_init(&acc_empty_bar[b], EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == WARP_TMEM) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.tiles_m * p.tiles_n;
    const int kblocks = (p.k + BKB - 1) / BKB;
    const int S = p.slices;

    if (warp == WARP_TMA) {
        // ===== TMA producer: for g = S-1 .. 0, for s = 0 .. g: A slice s against B slice g - s =====
        ptx::setmaxnreg_dec<REGS_CONTROL>();
        if (lane == 0) {
            ptx::prefetch_tensormap(&tmap_a);
            ptx::prefetch_tensormap(&tmap_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int tm, tn;
                tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
                for (int g = S - 1; g >= 0; --g)
                    for (int s = 0; s <= g; ++s) {
                        const int row_a = s * p.m_pad + p.row0_a + tm * BM, row_b = (g - s) * p.n_pad + p.row0_b + tn * BN;
                        for (int kb = 0; kb < kblocks; ++kb) {
                            tc::mbar_wait_guarded(&empty_bar[stage], phase ^ 1);
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                            unsigned char* sa = base + stage * STAGE_BYTES;
                            ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BKB, row_a);
                            ptx::tma_load_2d(sa + OPERAND_BYTES, &tmap_b, &full_bar[stage], kb * BKB, row_b);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ===== MMA issuer: one int32 window per group g (cut every MAX_WINDOW k-blocks) =====
        ptx::setmaxnreg_dec<REGS_CONTROL>();
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int g = S - 1; g >= 0; --g) {
                    const int group_kblocks = (g + 1) * kblocks;
                    uint32_t d_tmem = 0;
                    for (int kb = 0, wk = 0; kb < group_kblocks; ++kb) {
                        if (wk == 0) {
                            tc::mbar_wait_guarded(&acc_empty_bar[acc], acc_phase ^ 1);
                            tc::fence_after_thread_sync();
                            d_tmem = tmem_base + acc * BN;
                        }
                        tc::mbar_wait_guarded(&full_bar[stage], phase);
                        tc::fence_after_thread_sync();
                        const uint32_t sa = ptx::smem_u32(base + stage * STAGE_BYTES), sb = sa + OPERAND_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BKB / UMMA_KB; ++ks)
                            mma_i8(d_tmem, tc::smem_desc(p.desc, sa + ks * UMMA_KB), tc::smem_desc(p.desc, sb + ks * UMMA_KB), p.idesc, (wk | ks) ? 1u : 0u);
                        tc::mma_commit(&empty_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        if (++wk == MAX_WINDOW || kb == group_kblocks - 1) {
                            tc::mma_commit(&acc_full_bar[acc]);
                            if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
                            wk = 0;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= WARP_EPI0 && warp < WARP_EPI0 + EPI_WARPS) {
        // ===== accumulate + epilogue: warp = 4 + 4 h + q  ->  TMEM lanes 32q .. 32q+31 (rows), columns 64h .. 64h+63 =====
        ptx::setmaxnreg_inc<REGS_EPILOGUE>();
        const int q = (warp - WARP_EPI0) & 3, h = (warp - WARP_EPI0) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int tm, tn;
            tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
            double sum[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) sum[j] = 0.0;
            for (int g = S - 1; g >= 0; --g) {
                const int windows = ((g + 1) * kblocks + MAX_WINDOW - 1) / MAX_WINDOW;
                const double scale = scalbn(1.0, -(2 * P0 + SLICE_BITS * g));
                for (int w = 0; w < windows; ++w) {
                    tc::mbar_wait_guarded(&acc_full_bar[acc], acc_phase);
                    tc::fence_after_thread_sync();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * 64;
                    uint32_t v0[32], v1[32];
                    tc::tmem_ld_32x32b_x32(taddr, v0);
                    tc::tmem_ld_32x32b_x32(taddr + 32, v1);
                    tc::tmem_ld_wait();
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&acc_empty_bar[acc]);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        sum[j] += (double)(int)v0[j] * scale;        // |v| < 2^31: exact; the scale is a power of two
                        sum[32 + j] += (double)(int)v1[j] * scale;
                    }
                    if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
                }
            }
            const int row = tm * BM + q * 32 + lane;
            const bool row_ok = row < p.m;
            const int col0 = tn * BN + h * 64;
            int e_row = row_ok ? p.ea[row] : 0;
            const bool bad_row = e_row >= NON_FINITE;  // an Inf / NaN in this row of op(A): the whole row of the product is non-finite
            if (e_row <= NO_DATA || bad_row) e_row = 0;  // all-zero row: its sums are zero anyway
            double* cp = p.c + (int64_t)col0 * p.ldc + row;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                if (row_ok && col0 + j < p.n) {
                    int e_col = p.eb[col0 + j];
                    const bool bad = bad_row || e_col >= NON_FINITE;
                    if (e_col <= NO_DATA || bad) e_col = 0;
                    const double prod = bad ? __longlong_as_double(0x7FF8000000000000ll) : p.alpha * scalbn(sum[j], e_row + e_col);
                    cp[(int64_t)j * p.ldc] = p.read_c ? prod + p.beta * cp[(int64_t)j * p.ldc] : prod;
                }
            }
        }
    } else {
        ptx::setmaxnreg_dec<REGS_CONTROL>();  // TMEM warp, spare control warp, warps 12-15
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == WARP_TMEM) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

static CUresult make_map_i8(CUtensorMap* map, const void* base, uint64_t k, uint64_t rows, uint64_t pitch_bytes) {
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(
        tensormap_encode_fn());
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {k, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {BKB, BM};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}


// exponents + slices of one operand: `rows` rows of k values, element (i, l) at x[i * stride_row + l * stride_k]
static cudaError_t prepare(const double* x, int64_t stride_row, int64_t stride_k, int rows, int k, int slices, int* e, int8_t* out, int64_t pitch, int64_t slice_stride,
                           cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(e, 0x88, (size_t)rows * sizeof(int), st);  // every int below NO_DATA
    if (err != cudaSuccess) return err;
    if (stride_k == 1) {  // k contiguous: coalesced along k, one warp-level atomic per 1024 k-values
        row_exponents_kmajor<<<dim3((unsigned)std::min((k + 1023) / 1024, 8), (unsigned)std::min(rows, 32768)), 256, 0, st>>>(x, stride_row, rows, k, e);
    } else {
        const int k_per_block = 128;  // rows contiguous: a thread per row, the k range spread over grid.y for parallelism on narrow chunks
        row_exponents<<<dim3((unsigned)((rows + 255) / 256), (unsigned)((k + k_per_block - 1) / k_per_block)), 256, 0, st>>>(x, stride_row, stride_k, rows, k, k_per_block, e);
    }
    count_launch();
    if (stride_row == 1)  // rows contiguous in memory (op(A) = N, op(B) = T): coalesced along the rows
        slice_rows_contiguous<<<dim3((unsigned)((rows + 255) / 256), (unsigned)std::min((k + K_PER_THREAD - 1) / K_PER_THREAD, 4096)), 256, 0, st>>>(
            x, stride_k, rows, k, e, out, pitch, slice_stride, slices);
    else
        slice_rows<<<dim3((unsigned)((k + 1023) / 1024), (unsigned)std::min(rows, 32768)), 256, 0, st>>>(x, stride_row, stride_k, rows, k, e, out, pitch, slice_stride, slices);
    count_launch();
    return cudaGetLastError();
}


}  // namespace f64i8

// process-wide FP64 math mode: 0 = DMMA (default), otherwise the slice count of the int8 emulation (TMM_F64_MATH=i8 -> 8, i8:7 -> 7, ...)
int f64_i8_slices() {  // read per call (a getenv): tests and A/B scripts switch it between calls
    const char* e = getenv("TMM_F64_MATH");
    if (!e || strncmp(e, "i8", 2) != 0) return 0;
    const char* q = e + 2;
    int s = 8;
    if (q[0] == ':' && q[1]) s = atoi(q + 1);
    return s < 2 ? 2 : (s > f64i8::MAX_SLICES ? f64i8::MAX_SLICES : s);
}
size_t i8_slices_layout(int rows, int k, int slices, int64_t* rows_pad, int64_t* pitch) {
    const int64_t rp = f64i8::round_up(rows, f64i8::BM), pt = f64i8::round_up(k, 128);
    if (rows_pad) *rows_pad = rp;
    if (pitch) *pitch = pt;
    return (size_t)slices * (size_t)rp * (size_t)pt;
}

cudaError_t i8_slice_operand(const double* x, int64_t stride_row, int64_t stride_k, int rows, int k, const I8Slices& out, cudaStream_t st) {
    if (rows <= 0 || k <= 0) return cudaSuccess;
    return f64i8::prepare(x, stride_row, stride_k, rows, k, out.slices, out.e, out.q, out.pitch, out.rows_pad * out.pitch, st);
}

cudaError_t i8_gemm_sliced(const I8Slices& a, int a_row0, int m, const I8Slices& b, int b_row0, int n, double alpha, double beta, double* c, int64_t ldc, cudaStream_t st) {
    using namespace f64i8;
    if (m <= 0 || n <= 0) return cudaSuccess;
    if (a.k != b.k || a.slices != b.slices || a.pitch != b.pitch || a_row0 < 0 || b_row0 < 0 || a_row0 + m > a.rows || b_row0 + n > b.rows) return cudaErrorInvalidValue;
    const int S = a.slices, k = a.k;
    if ((int64_t)S * std::max(a.rows_pad, b.rows_pad) > INT32_MAX) return cudaErrorInvalidValue;  // TMA coordinates are 32-bit
    CUtensorMap map_a, map_b;
    if (make_map_i8(&map_a, a.q, (uint64_t)k, (uint64_t)S * a.rows_pad, (uint64_t)a.pitch) != CUDA_SUCCESS ||
        make_map_i8(&map_b, b.q, (uint64_t)k, (uint64_t)S * b.rows_pad, (uint64_t)b.pitch) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    Params p;
    p.c = c; p.ldc = ldc; p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
    p.read_c = beta != 0.0;
    p.tiles_m = (m + BM - 1) / BM; p.tiles_n = (n + BN - 1) / BN;
    p.slices = S; p.m_pad = (int)a.rows_pad; p.n_pad = (int)b.rows_pad;
    p.row0_a = a_row0; p.row0_b = b_row0;
    p.ea = a.e + a_row0; p.eb = b.e + b_row0;
    p.desc = tc::smem_desc_template(16, 8 * BKB, tc::LAYOUT_SW128);  // 128-byte rows, 8-row swizzle atoms 1024 B apart
    //        D = S32 (2)   A, B = signed 8 bit (1)        both K-major        N >> 3                     M >> 4
    p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dgemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
    const int grid = (int)std::min<int64_t>(tiles, sm_count());
    dgemm_i8_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map_a, map_b, p);
    count_launch();
    return cudaGetLastError();
}

// One self-contained product: slice both operands into stream-ordered scratch, multiply, release.
// Returns cudaErrorMemoryAllocation when the slice scratch cannot be had (the caller then runs the DMMA kernel); any other error is final.
cudaError_t dgemm_i8_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb, double beta, double* c,
                            int64_t ldc, cudaStream_t st, int slices) {
    using namespace f64i8;
    if (m <= 0 || n <= 0 || k <= 0) return cudaSuccess;
    const int S = slices < 2 ? 2 : (slices > MAX_SLICES ? MAX_SLICES : slices);
    I8Slices sa, sb;
    sa.rows = m; sb.rows = n; sa.k = sb.k = k; sa.slices = sb.slices = S;
    const size_t bytes_a = i8_slices_layout(m, k, S, &sa.rows_pad, &sa.pitch), bytes_b = i8_slices_layout(n, k, S, &sb.rows_pad, &sb.pitch);
    if ((int64_t)S * std::max(sa.rows_pad, sb.rows_pad) > INT32_MAX) return cudaErrorMemoryAllocation;  // TMA coordinates are 32-bit
    int* ex = nullptr;
    cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&sa.q), bytes_a, st);
    if (e == cudaSuccess) e = scratch_alloc(reinterpret_cast<void**>(&sb.q), bytes_b, st);
    if (e == cudaSuccess) e = scratch_alloc(reinterpret_cast<void**>(&ex), (size_t)(m + n) * sizeof(int), st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        scratch_free(sa.q, st);
        scratch_free(sb.q, st);
        return cudaErrorMemoryAllocation;
    }
    sa.e = ex; sb.e = ex + m;
    // op(A) row i, k index l:  N: a[l * lda + i]   T/C: a[i * lda + l]        op(B) column j, k index l:  N: b[j * ldb + l]   T/C: b[l * ldb + j]
    e = i8_slice_operand(a, ta == 'N' ? 1 : lda, ta == 'N' ? lda : 1, m, k, sa, st);
    if (e == cudaSuccess) e = i8_slice_operand(b, tb == 'N' ? ldb : 1, tb == 'N' ? 1 : ldb, n, k, sb, st);
    if (e == cudaSuccess) e = i8_gemm_sliced(sa, 0, m, sb, 0, n, alpha, beta, c, ldc, st);
    scratch_free(sa.q, st);
    scratch_free(sb.q, st);
    scratch_free(ex, st);
    return e;
}

}  // namespace tmm
