// DGEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C, column-major, device operands.
// Replaces blas_api::dgemm (reference gpu_blas_api.hpp:213-231, called from tiled_mm.cpp:200-217).
//
// tcgen05 has no FP64 kind, so FP64 stays on the warp-level tensor path: mma.sync m8n8k4 (SASS
// DMMA.8x8x4, 64 FMA/clk/SM = the chip's whole FP64 rate).  Structure:
//   * one CTA per 128 x 64 tile of C, TWO CTAs resident per SM: while one CTA runs its prologue
//     (first TMA round trip) or epilogue (C read-modify-write), the other owns the FP64 pipe;
//     the hardware block scheduler balances tiles dynamically and back-fills the tail of one
//     launch with the next launch from another stream.  Column-group raster for L2 reuse.
//   * two warpgroups per CTA.  Warpgroup 1 is the producer: it gives its registers back (setmaxnreg.dec 24) and one
//     lane drives TMA (cp.async.bulk.tensor.2d) into a 4-6 stage shared-memory ring, completion on mbarriers.
//     Warpgroup 0 = 4 math warps (2 x 2) that take those registers (setmaxnreg.inc 232): each owns a 64 x 32 block of
//     the CTA tile = 8 x 4 DMMA tiles = 128 accumulator registers per lane, with room left for fragments, the C
//     prefetch and the read-modify-write epilogue without a single spill (168 registers, the 2-CTA/SM cap without
//     the split, spilled inside the main loop: ncu long_scoreboard stalls, -3.7 % - profiles/r1_ncu_dgemm.md)
//   * operands are NEVER transposed or repacked: the TMA box is taken straight from the
//     column-major device panel in whichever orientation it is stored (N: m-/n-contiguous,
//     T/C: k-contiguous).  The box's contiguous extent is over-fetched by PAD = 4 doubles, which
//     makes the shared-memory row stride = 4 (mod 16) doubles -> every DMMA fragment load
//     (lane (g,t) reads [g][t] or [t][g]) is bank-conflict free in both orientations without
//     a swizzle; out-of-range box parts are zero-filled by TMA, which also handles all m/n/k edges
//   * epilogue: alpha * acc (+ beta * C) straight from the accumulator registers, plain global
//     accesses, so C may have any leading dimension (the device-resident C of copy_c_back=false has
//     ld = m, reference tiled_mm.cpp:446).
#include "tmm_blas.h"
#include "tmm_ptx.cuh"

#include <cstdio>
#include <type_traits>

namespace tmm {
namespace f64 {

constexpr int BM = 128, BN = 64, BK = 16, PAD = 4;
constexpr int MATH_WARPS = 4, THREADS = 2 * MATH_WARPS * 32;  // warpgroup 0 = math, warpgroup 1 = TMA producer (one active lane)
constexpr int MATH_REGS = 232, PRODUCER_REGS = 24;            // setmaxnreg split of the 2 x 128-register launch allocation
constexpr int SMEM_BUDGET = 113 * 1024;  // two CTAs per SM
constexpr int WM = 64, WN = 32;           // warp tile
constexpr int MI = WM / 8, NJ = WN / 8;   // DMMA tiles per warp
constexpr int RS_M = BM + PAD;            // row stride (doubles) of an m-contiguous A stage: [BK][BM+PAD]
constexpr int RS_N = BN + PAD;            // row stride of an n-contiguous B stage:          [BK][BN+PAD]
constexpr int RS_K = BK + PAD;            // row stride of a k-contiguous stage:             [BM|BN][BK+PAD]
constexpr int GROUP_COLS = 16;
static_assert(RS_M % 16 == 4 && RS_N % 16 == 4 && RS_K % 16 == 4, "row stride must be 4 mod 16 doubles for conflict-free fragment loads");

template <bool A_KMAJOR, bool B_NMAJOR>
struct Cfg {
    static constexpr int A_ELEMS = A_KMAJOR ? BM * RS_K : BK * RS_M;
    static constexpr int B_ELEMS = B_NMAJOR ? BK * RS_N : BN * RS_K;
    static constexpr int A_BYTES = A_ELEMS * 8, B_BYTES = B_ELEMS * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_BUDGET - 256) / STAGE_BYTES > 6 ? 6 : (SMEM_BUDGET - 256) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
    static_assert(A_BYTES % 128 == 0 && B_BYTES % 128 == 0, "TMA destination alignment");
    static_assert(STAGES >= 3, "need at least 3 stages");
};

struct Params {
    double* c;
    int64_t ldc;
    int m, n, k;
    double alpha, beta;
    int read_c;  // 0: C = alpha*acc (C never read, NaN-safe, reference tiled_mm.cpp:325); 1: += beta*C
    int tiles_m, tiles_n;
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

template <bool A_KMAJOR, bool B_NMAJOR>
__global__ void __launch_bounds__(THREADS, 2)
dgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    using C = Cfg<A_KMAJOR, B_NMAJOR>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve: [stage 0: A | B][stage 1: A | B]...[full barriers][empty barriers]
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    // read through a shuffle so that the compiler knows the warp index (and every branch taken on it: the roles, the edge path of the main
    // loop) to be warp-uniform and keeps the pipeline bookkeeping inside those branches on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], MATH_WARPS);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    int tm, tn;
    tile_coords(blockIdx.x, p.tiles_m, p.tiles_n, tm, tn);
    const int kblocks = (p.k + BK - 1) / BK;

    if (warp >= MATH_WARPS) {
        // ===== producer warpgroup: hands its registers to the math warpgroup, then one lane drives TMA =====
        ptx::setmaxnreg_dec<PRODUCER_REGS>();
        if (warp == MATH_WARPS && lane == 0) {
            ptx::prefetch_tensormap(&tmap_a);
            ptx::prefetch_tensormap(&tmap_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                unsigned char* sa = base + stage * C::STAGE_BYTES;
                unsigned char* sb = sa + C::A_BYTES;
                if (A_KMAJOR) ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, tm * BM);
                else          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], tm * BM, kb * BK);
                if (B_NMAJOR) ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], tn * BN, kb * BK);
                else          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, tn * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===== math warps (2 x 2) =====
    ptx::setmaxnreg_inc<MATH_REGS>();
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * WM;  // 0 or 64
    const int wn = (warp & 1) * WN;   // 0 or 32
    // per-lane fragment base offsets (doubles) inside a stage
    const int a_off = A_KMAJOR ? (wm + g) * RS_K + t : t * RS_M + wm + g;
    const int b_off = (B_NMAJOR ? t * RS_N + wn + g : (wn + g) * RS_K + t) + C::A_ELEMS;
    constexpr int A_I_STRIDE = A_KMAJOR ? 8 * RS_K : 8;      // next DMMA tile along m
    constexpr int A_K_STRIDE = A_KMAJOR ? 4 : 4 * RS_M;      // next k4 step
    constexpr int B_J_STRIDE = B_NMAJOR ? 8 : 8 * RS_K;
    constexpr int B_K_STRIDE = B_NMAJOR ? 4 * RS_N : 4;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // Edge tiles: TMA zero-fills what lies past m / n, and a warp would multiply those zeros at full price (dgemm 10000^3: 79 x 157 tiles cover
    // 10112 x 10048, 1.6 % of all DMMAs).  The count of 8-row / 8-column DMMA tiles that hold at least one element of C is warp-uniform, so an
    // edge warp predicates its DMMAs on it (a warp entirely outside C issues none and only keeps the pipeline's barriers moving); interior warps
    // run the unpredicated loop.
    const int mi_valid = min(MI, max(0, (p.m - (tm * BM + wm) + 7) >> 3));
    const int nj_valid = min(NJ, max(0, (p.n - (tn * BN + wn) + 7) >> 3));

    auto main_loop = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        int stage = 0;
        uint32_t phase = 0;
        // read-modify-write launches pull this warp's 64 x 32 block of C towards L2 a few k-blocks before the epilogue
        // needs it (early enough to cover the HBM latency, late enough not to be evicted again by the A/B stream)
        const int prefetch_kb = p.read_c ? max(0, kblocks - 12) : -1;
        for (int kb = 0; kb < kblocks; ++kb) {
            if (kb == prefetch_kb) {
                // lane l covers column wn + l: 64 rows = 512 B = 4 (5 when unaligned) 128-byte lines
                const int col = tn * BN + wn + lane;
                if (col < p.n) {
                    const char* cp = reinterpret_cast<const char*>(p.c + (int64_t)col * p.ldc + tm * BM + wm);
                    const int rows = min(WM, p.m - (tm * BM + wm));
#pragma unroll
                    for (int o = 0; o < 5; ++o)
                        if (o * 128 < rows * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + o * 128));
                }
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            const double* st = reinterpret_cast<const double*>(base + stage * C::STAGE_BYTES);
            const double* as = st + a_off;
            const double* bs = st + b_off;
            if (!EDGE || (mi_valid > 0 && nj_valid > 0)) {
#pragma unroll
                for (int ks = 0; ks < BK / 4; ++ks) {
                    double af[MI], bf[NJ];
#pragma unroll
                    for (int i = 0; i < MI; ++i) af[i] = as[ks * A_K_STRIDE + i * A_I_STRIDE];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) bf[j] = bs[ks * B_K_STRIDE + j * B_J_STRIDE];
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        if (EDGE && i >= mi_valid) continue;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (EDGE && j >= nj_valid) continue;
                            ptx::dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    };
    if (mi_valid == MI && nj_valid == NJ) main_loop(std::false_type{});
    else main_loop(std::true_type{});

    // epilogue: lane (g,t) owns rows 8i+g, columns 8j+2t, 8j+2t+1 of its warp tile
    const int row0 = tm * BM + wm + g;
    const int col0 = tn * BN + wn + 2 * t;
    if (p.read_c) {
        // read-modify-write: all 16 loads of a column pair are issued before any is consumed
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            double old[2][MI];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = col0 + 8 * j + h;
                const double* cp = p.c + (int64_t)col * p.ldc;
#pragma unroll
                for (int i = 0; i < MI; ++i) old[h][i] = (col < p.n && row0 + 8 * i < p.m) ? __ldcs(cp + row0 + 8 * i) : 0.0;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = col0 + 8 * j + h;
                if (col < p.n) {
                    double* cp = p.c + (int64_t)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        if (row0 + 8 * i < p.m) cp[row0 + 8 * i] = p.alpha * acc[i][j][h] + p.beta * old[h][i];
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = col0 + 8 * j + h;
                if (col < p.n) {
                    double* cp = p.c + (int64_t)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < MI; ++i)
                        if (row0 + 8 * i < p.m) cp[row0 + 8 * i] = p.alpha * acc[i][j][h];
                }
            }
        }
    }
}

static CUresult make_map(CUtensorMap* map, const double* base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems, uint32_t box0, uint32_t box1) {
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(tensormap_encode_fn());
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld_elems * 8};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <bool AK, bool BN_>
static cudaError_t launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t st) {
    using C = Cfg<AK, BN_>;
    static bool configured[64] = {false};
    auto kern = dgemm_kernel<AK, BN_>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured[dev] = true;
    }
    const int64_t grid = (int64_t)p.tiles_m * p.tiles_n;
    if (grid > INT32_MAX) return cudaErrorInvalidValue;
    kern<<<(unsigned)grid, THREADS, C::SMEM_BYTES, st>>>(ta, tb, p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace f64

cudaError_t dgemm_launch(char ta, char tb, int m, int n, int k, double alpha, const double* a, int64_t lda, const double* b, int64_t ldb,
                         double beta, double* c, int64_t ldc, cudaStream_t stream) {
    using namespace f64;
    if (m <= 0 || n <= 0) return cudaSuccess;
    const bool a_k = (ta != 'N'), b_n = (tb != 'N');
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (lda & 1) || (ldb & 1)) return cudaErrorInvalidValue;
    CUtensorMap map_a, map_b;
    CUresult r;
    // A: op(A) is m x k.  N: stored m x k (m contiguous);  T/C: stored k x m (k contiguous)
    r = a_k ? make_map(&map_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, BK + PAD, BM) : make_map(&map_a, a, (uint64_t)m, (uint64_t)k, (uint64_t)lda, BM + PAD, BK);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(A) failed: %d\n", (int)r); return cudaErrorInvalidValue; }
    // B: op(B) is k x n.  N: stored k x n (k contiguous);  T/C: stored n x k (n contiguous)
    r = b_n ? make_map(&map_b, b, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, BN + PAD, BK) : make_map(&map_b, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, BK + PAD, BN);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(B) failed: %d\n", (int)r); return cudaErrorInvalidValue; }
    Params p;
    p.c = c; p.ldc = ldc; p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
    p.read_c = (beta != 0.0);
    p.tiles_m = (m + BM - 1) / BM; p.tiles_n = (n + BN - 1) / BN;
    if (a_k) return b_n ? launch<true, true>(map_a, map_b, p, stream) : launch<true, false>(map_a, map_b, p, stream);
    return b_n ? launch<false, true>(map_a, map_b, p, stream) : launch<false, false>(map_a, map_b, p, stream);
}

}  // namespace tmm
