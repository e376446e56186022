// ZGEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C, complex<double>, column-major, device operands.
// Replaces blas_api::zgemm (reference gpu_blas_api.hpp:233-252, called from tiled_mm.cpp:243-266).
//
// Same skeleton as the DGEMM kernel (gemm_f64.cu): one producer warp drives TMA into a shared-memory ring,
// four math warps (with the producer warpgroup's registers, setmaxnreg) issue FP64 DMMA.8x8x4, two CTAs per SM.  What is specific to complex:
//   * operands stay INTERLEAVED (re, im) exactly as std::complex<double> stores them - no planar repack on the
//     host, the copy engines or the device.  The TMA descriptor views the matrix as doubles with a doubled
//     contiguous extent; one LDS.128 delivers (re, im) of a fragment element.
//   * one complex DMMA tile = 4 real DMMAs:  re += ar*br;  re += x*bi;  im += z*bi;  im += y*br   with
//     x = -/+ai, y = +/-ai, z = +/-ar chosen by sign-bit masks, so 'C' (conjugate) costs nothing: conjugation is a
//     sign flip (one LOP3 on the high word) applied to the fragment registers, never a pass over memory.
//   * CTA tile 64 x 64 complex, warp tile 32 x 32 complex = 4 x 4 DMMA tiles x (re, im) = 128 accumulator
//     registers per lane (the same budget as the real kernel), 64 DMMAs per 8 LDS.128.
//   * shared-memory row strides in 16-byte units: m-/n-contiguous stage rows 64 + 2 = 66 (= 2 mod 8), k-contiguous
//     rows 8 + 4 = 12 (= 4 mod 8): each quarter-warp of an LDS.128 fragment load touches 8 distinct 16-byte bank groups
//     in both orientations, without a swizzle.  TMA zero-fills out-of-range parts of the box (all m/n/k edges).
#include "tmm_blas.h"
#include "tmm_ptx.cuh"

#include <cstdio>
#include <type_traits>

namespace tmm {
namespace c64 {

constexpr int BM = 64, BN = 64, BK = 8;  // complex elements
constexpr int PAD_MN = 2, PAD_K = 4;
constexpr int MATH_WARPS = 4, THREADS = 2 * MATH_WARPS * 32;  // warpgroup 0 = math, warpgroup 1 = TMA producer (one active lane)
constexpr int MATH_REGS = 232, PRODUCER_REGS = 24;            // setmaxnreg split of the 2 x 128-register launch allocation
constexpr int SMEM_BUDGET = 113 * 1024;  // two CTAs per SM
constexpr int WM = 32, WN = 32;
constexpr int MI = WM / 8, NJ = WN / 8;
constexpr int RS_M = BM + PAD_MN, RS_N = BN + PAD_MN, RS_K = BK + PAD_K;  // row strides in complex elements
constexpr int GROUP_COLS = 16;
static_assert(RS_M % 8 == 2 && RS_N % 8 == 2 && RS_K % 8 == 4, "row strides chosen for conflict-free LDS.128 fragment loads");

template <bool A_KMAJOR, bool B_NMAJOR>
struct Cfg {
    static constexpr int A_ELEMS = A_KMAJOR ? BM * RS_K : BK * RS_M;
    static constexpr int B_ELEMS = B_NMAJOR ? BK * RS_N : BN * RS_K;
    static constexpr int A_BYTES = A_ELEMS * 16, B_BYTES = B_ELEMS * 16;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_BUDGET - 256) / STAGE_BYTES > 6 ? 6 : (SMEM_BUDGET - 256) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
    static_assert(A_BYTES % 128 == 0 && B_BYTES % 128 == 0, "TMA destination alignment");
    static_assert(STAGES >= 3, "need at least 3 stages");
};

struct Params {
    double2* c;
    int64_t ldc;  // complex elements
    int m, n, k;
    double alpha_re, alpha_im, beta_re, beta_im;
    int read_c;
    int tiles_m, tiles_n;
    uint32_t mask_x, mask_y, mask_z;  // sign-bit masks (0 or 0x80000000) for x = -+ai, y = +-ai, z = +-ar
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

__device__ __forceinline__ double flip(double v, uint32_t mask) {
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

template <bool A_KMAJOR, bool B_NMAJOR>
__global__ void __launch_bounds__(THREADS, 2)
zgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    using C = Cfg<A_KMAJOR, B_NMAJOR>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
                // coordinates along the contiguous dimension are in doubles (2 per complex element)
                if (A_KMAJOR) ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], 2 * kb * BK, tm * BM);
                else          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], 2 * tm * BM, kb * BK);
                if (B_NMAJOR) ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], 2 * tn * BN, kb * BK);
                else          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], 2 * kb * BK, tn * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    ptx::setmaxnreg_inc<MATH_REGS>();
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * WM;
    const int wn = (warp & 1) * WN;
    const int a_off = A_KMAJOR ? (wm + g) * RS_K + t : t * RS_M + wm + g;
    const int b_off = (B_NMAJOR ? t * RS_N + wn + g : (wn + g) * RS_K + t) + C::A_ELEMS;
    constexpr int A_I_STRIDE = A_KMAJOR ? 8 * RS_K : 8;
    constexpr int A_K_STRIDE = A_KMAJOR ? 4 : 4 * RS_M;
    constexpr int B_J_STRIDE = B_NMAJOR ? 8 : 8 * RS_K;
    constexpr int B_K_STRIDE = B_NMAJOR ? 4 * RS_N : 4;

    double acr[MI][NJ][2], aci[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acr[i][j][0] = acr[i][j][1] = 0.0; aci[i][j][0] = aci[i][j][1] = 0.0; }

    // edge tiles: DMMA tiles that lie entirely past m / n are skipped (warp-uniform counts; see gemm_f64.cu)
    const int mi_valid = min(MI, max(0, (p.m - (tm * BM + wm) + 7) >> 3));
    const int nj_valid = min(NJ, max(0, (p.n - (tn * BN + wn) + 7) >> 3));

    auto main_loop = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        int stage = 0;
        uint32_t phase = 0;
        const int prefetch_kb = p.read_c ? max(0, kblocks - 12) : -1;
        for (int kb = 0; kb < kblocks; ++kb) {
            if (kb == prefetch_kb) {
                // lane l covers column wn + l of this warp's 32 x 32 block of C: 32 rows = 512 B = 4 (5 if unaligned) lines
                const int col = tn * BN + wn + lane;
                if (col < p.n) {
                    const char* cp = reinterpret_cast<const char*>(p.c + (int64_t)col * p.ldc + tm * BM + wm);
                    const int rows = min(WM, p.m - (tm * BM + wm));
#pragma unroll
                    for (int o = 0; o < 5; ++o)
                        if (o * 128 < rows * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + o * 128));
                }
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            const double2* st = reinterpret_cast<const double2*>(base + stage * C::STAGE_BYTES);
            const double2* as = st + a_off;
            const double2* bs = st + b_off;
            if (!EDGE || (mi_valid > 0 && nj_valid > 0)) {
#pragma unroll
                for (int ks = 0; ks < BK / 4; ++ks) {
                    double br[NJ], bi[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const double2 v = bs[ks * B_K_STRIDE + j * B_J_STRIDE];
                        br[j] = v.x; bi[j] = v.y;
                    }
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        if (EDGE && i >= mi_valid) continue;
                        const double2 v = as[ks * A_K_STRIDE + i * A_I_STRIDE];
                        // all sign handling sits on the A fragment (4 values) so the B fragments stay untouched
                        const double ar = v.x, az = flip(v.x, p.mask_z), ax = flip(v.y, p.mask_x), ay = flip(v.y, p.mask_y);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (EDGE && j >= nj_valid) continue;
                            ptx::dmma_884(acr[i][j][0], acr[i][j][1], ar, br[j]);
                            ptx::dmma_884(aci[i][j][0], aci[i][j][1], az, bi[j]);
                            ptx::dmma_884(acr[i][j][0], acr[i][j][1], ax, bi[j]);
                            ptx::dmma_884(aci[i][j][0], aci[i][j][1], ay, br[j]);
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

    // epilogue: lane (g,t) owns rows 8i+g, columns 8j+2t, 8j+2t+1 of its warp tile; one 16-byte access per element
    const int row0 = tm * BM + wm + g;
    const int col0 = tn * BN + wn + 2 * t;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = col0 + 8 * j + h;
            if (col >= p.n) continue;
            double2* cp = p.c + (int64_t)col * p.ldc;
            double2 old[MI];
            if (p.read_c) {
#pragma unroll
                for (int i = 0; i < MI; ++i) old[i] = (row0 + 8 * i < p.m) ? __ldcs(cp + row0 + 8 * i) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                if (row0 + 8 * i < p.m) {
                    const double re = acr[i][j][h], im = aci[i][j][h];
                    double2 v;
                    v.x = p.alpha_re * re - p.alpha_im * im;
                    v.y = p.alpha_re * im + p.alpha_im * re;
                    if (p.read_c) {
                        v.x += p.beta_re * old[i].x - p.beta_im * old[i].y;
                        v.y += p.beta_re * old[i].y + p.beta_im * old[i].x;
                    }
                    cp[row0 + 8 * i] = v;
                }
            }
        }
    }
}

// the matrix seen as doubles: contiguous extent 2 * rows, row pitch ld * 16 bytes
static CUresult make_map(CUtensorMap* map, const void* base, uint64_t rows_c, uint64_t cols, uint64_t ld_c, uint32_t box_rows_c, uint32_t box_cols) {
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(tensormap_encode_fn());
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {2 * rows_c, cols};
    cuuint64_t strides[1] = {ld_c * 16};
    cuuint32_t box[2] = {2 * box_rows_c, box_cols};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <bool AK, bool BN_>
static cudaError_t launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t st) {
    using C = Cfg<AK, BN_>;
    static bool configured[64] = {false};
    auto kern = zgemm_kernel<AK, BN_>;
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

}  // namespace c64

cudaError_t zgemm_launch(char ta, char tb, int m, int n, int k, const double* al, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const double* be, void* c, int64_t ldc, cudaStream_t stream) {
    using namespace c64;
    if (m <= 0 || n <= 0) return cudaSuccess;
    const bool a_k = (ta != 'N'), b_n = (tb != 'N');
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return cudaErrorInvalidValue;
    CUtensorMap map_a, map_b;
    CUresult r;
    r = a_k ? make_map(&map_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, BK + PAD_K, BM) : make_map(&map_a, a, (uint64_t)m, (uint64_t)k, (uint64_t)lda, BM + PAD_MN, BK);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(A, complex) failed: %d\n", (int)r); return cudaErrorInvalidValue; }
    r = b_n ? make_map(&map_b, b, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, BN + PAD_MN, BK) : make_map(&map_b, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, BK + PAD_K, BN);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(B, complex) failed: %d\n", (int)r); return cudaErrorInvalidValue; }
    Params p;
    p.c = static_cast<double2*>(c); p.ldc = ldc; p.m = m; p.n = n; p.k = k;
    p.alpha_re = al[0]; p.alpha_im = al[1]; p.beta_re = be[0]; p.beta_im = be[1];
    p.read_c = (be[0] != 0.0 || be[1] != 0.0);
    p.tiles_m = (m + BM - 1) / BM; p.tiles_n = (n + BN - 1) / BN;
    // effective a = (ar, sa*ai), b = (br, sb*bi), sa/sb = -1 for 'C':  re += ar*br + (-sa*sb*ai)*bi ;  im += (sb*ar)*bi + (sa*ai)*br
    const bool ca = (ta == 'C'), cb = (tb == 'C');
    constexpr uint32_t SIGN = 0x80000000u;
    p.mask_x = (ca != cb) ? 0u : SIGN;
    p.mask_y = ca ? SIGN : 0u;
    p.mask_z = cb ? SIGN : 0u;  // applied to ar
    if (a_k) return b_n ? launch<true, true>(map_a, map_b, p, stream) : launch<true, false>(map_a, map_b, p, stream);
    return b_n ? launch<false, true>(map_a, map_b, p, stream) : launch<false, false>(map_a, map_b, p, stream);
}

}  // namespace tmm
