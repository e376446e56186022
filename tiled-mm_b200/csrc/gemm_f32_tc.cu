// SGEMM for sm_100a on the 5th-generation tensor cores:  C = alpha * op(A) * op(B) + beta * C, column-major, device operands.
// Replaces blas_api::sgemm (reference gpu_blas_api.hpp:194-211, called from tiled_mm.cpp:181-198).
//
// The reference runs cuBLAS in its default math mode (gpu_blas_handle.hpp:11-17 never enables TF32), i.e. FP32-accurate
// results.  tcgen05 has no FP32 kind, so every FP32 operand element x is split into two TF32 numbers
//     hi = tf32(x)   (round to nearest, 11 significant bits)          lo = tf32(x - hi)   (x - hi is exact)
// and each product is evaluated as  lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  by three kind::tf32 MMAs that accumulate in FP32
// in tensor memory ("3xTF32": the dropped lo*lo term and the rounding of lo are both ~2^-24 relative, the size of an
// FP32 rounding error).  Small integers are exact in TF32 (lo = 0), so integer-valued parity tests stay bit-exact.
// mode = 1 issues only hi_a*hi_b on the raw FP32 bits (plain TF32, ~3x faster, 2^-11 relative error) - opt-in.
//
// Structure - one persistent CTA per SM, 16 warps with fixed roles, everything hand-written (TMA, mbarrier, tcgen05):
//   warp 0      TMA producer: raw FP32 operand tiles straight from the column-major panels in their stored orientation
//               (never transposed or repacked in global memory), 128B-swizzled, into a 3-stage ring
//   warps 8-15  split stage: each 16-byte chunk of the landed tile is rewritten in place as `hi` and its `lo` twin is
//               written at the same offset of a second buffer.  The pass is purely element-wise, so it is independent of
//               the operand's orientation and swizzle; fence.proxy.async publishes it to the tensor core
//   warp 1      MMA issuer: one lane issues 12 tcgen05.mma (4 k-steps x 3 terms, 128x128x8 each) per stage, operands
//               described in place: k-contiguous tiles (A^T, B) as K-major SWIZZLE_128B, m/n-contiguous tiles (A, B^T)
//               as MN-major "128B swizzle, 32B atoms" (the only MN-major layout the TF32 kind accepts) - so all four
//               transpose combinations run the same kernel with different descriptors.  tcgen05.commit releases the
//               stage to the producer and, after the last k-block, hands the accumulator to the epilogue
//   warps 4-7   accumulate + epilogue.  The tensor core adds into its FP32 accumulator with truncation (measured: the error
//               of a plain TMEM accumulation grows like k^1.5, 30x the error of an FFMA chain at k = 4096), so TMEM only
//               ever holds the partial sum of a short k-window (default 4 k-blocks = 128): after each window these warps
//               drain it with tcgen05.ld (TMEM lane = row, column = column) and add it - round to nearest - into 128 FP32
//               registers per thread (setmaxnreg gives this warpgroup 232 registers, taken from the other roles).  Four
//               window accumulators (4 x 128 columns = all of TMEM) rotate, so draining and the tile's global-memory
//               epilogue overlap the MMAs of the next windows / the next tile.
//               Tile end: alpha/beta and plain coalesced global stores (any ldc; the device-resident C of
//               copy_c_back=false has ld = m, reference tiled_mm.cpp:446)
//   warp 2      TMEM allocation / release
// Tiles are handed out DYNAMICALLY: the producer draws tile numbers from a per-launch counter in global memory (atomicAdd) and passes them to the
// other roles through a four-entry ring in shared memory (tile_full / tile_empty mbarriers; -1 ends the kernel).  The scheduler runs several of
// these launches at once on different streams (column stripes of phase 1, column blocks of phase 2); a persistent CTA keeps its SM until its
// launch has no tiles left, so with the static stride (tile += gridDim.x) a CTA that got its SM late finished late and the launch ended with
// most SMs idle (cgemm 8000^3 host-to-host: the last phase-2 block, 1 ms of work, took 6.2 ms - profiles/r2_cgemm_diag.txt).  With the
// counter a late CTA simply finds fewer tiles.  The last CTA to leave zeroes the counter pair for the launch that takes the slot next.
// TMM_TC_SCHED=static selects the fixed stride.
// All m / n / k edges are handled by TMA zero fill plus masked stores.
#include "tmm_blas.h"
#include "tmm_tc.cuh"
#include "tmm_prepass.cuh"  // split_tf32, lo_of_truncated (device code only; also compiled for the CPU by tests/test_prepass_kernels.py)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace tmm {
namespace f32tc {

constexpr int BM = 128, BN = 128, BK = 32, UMMA_K = 8;
constexpr int STAGES = 3;
constexpr int OPERAND_BYTES = BM * BK * 4;      // one 128 x 32 FP32 tile (BM == BN)
constexpr int STAGE_BYTES = 4 * OPERAND_BYTES;  // A hi | A lo | B hi | B lo
constexpr int ATOM_MN = 32;                     // floats per 128-byte swizzle row of an MN-major tile
constexpr int MN_BOX_BYTES = ATOM_MN * BK * 4;  // one [32 (m|n) x BK] TMA box of an MN-major tile
constexpr int ACC_BUFS = 4;                     // window accumulators in TMEM (4 x 128 columns = all of it)
constexpr int TMEM_COLS = ACC_BUFS * BN;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_TMEM = 2, WARP_EPI0 = 4, WARP_SPLIT0 = 8, SPLIT_WARPS = 8;
constexpr int THREADS = (WARP_SPLIT0 + SPLIT_WARPS) * 32;
constexpr int GROUP_COLS = 16;
constexpr int REGS_CONTROL = 56, REGS_SPLIT = 88, REGS_EPILOGUE = 232;  // setmaxnreg split of the 512 x 128 launch allocation
constexpr int WINDOW_KBLOCKS = 4;  // k-blocks summed in TMEM before promotion to the FP32 register accumulators
constexpr int TILE_RING = 4;       // tile numbers in flight between the producer and the other roles
constexpr int SCHED_SLOTS = 4096;  // counter pairs in global memory, one per launch in flight (taken round robin)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
static_assert(BM == BN, "operand tiles share one size");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
    float* c;
    int64_t ldc;
    int m, n, k;
    float alpha, beta;
    int read_c;  // 0: C = alpha*acc (C never read, NaN-safe, reference tiled_mm.cpp:325); 1: += beta*C
    int tiles_m, tiles_n;
    int a_mn_major, b_mn_major;  // operand orientation in shared memory (A: op N, B: op T/C)
    uint64_t desc_a, desc_b;     // shared-memory descriptor templates (everything but the address)
    uint32_t kstep_a, kstep_b;   // bytes between consecutive UMMA_K slices of a tile
    uint32_t idesc;
    int terms;   // 3: FP32-accurate 3xTF32;  1: plain TF32 on the raw bits
    int window;  // k-blocks per TMEM accumulation window
    int bk;          // operand elements per k-block: 32 FP32 (= 128 B); 64 for the 16-bit variant below - the byte geometry is the same
    int f16_kind;    // experimental: operands are bf16, K-major, multiplied by kind::f16 MMAs (one term); see bgemm_tc_native_launch
    int* sched;       // [0] next tile, [1] CTAs that have drawn their last number; nullptr: static stride
    int split_trunc;  // experimental (TMM_TC_SPLIT=trunc): hi = the raw FP32 bits (the tensor core ignores the low 13 mantissa bits), only lo is written
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

// The kernel body, parameterised by the ring geometry: NSTAGES stages of [A | (A lo) | B | (B lo)] with B at B_OFF bytes from A.
//   <3, 2 * OPERAND_BYTES>  the FP32-accurate layout (hi | lo twins, 64 KB per stage)                     -> sgemm_tc_kernel
//   (a six-stage ring of raw tiles for the one-term modes and variants taking A from tensor memory / on CTA pairs were measured in round 2 and
//    removed: 241 vs 243 TF, and 77 - 88 vs 155 TF - profiles/r2_tc_variants.txt)
template <int NSTAGES, int B_OFF>
__device__ __forceinline__ void sgemm_tc_body(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const Params& p) {
    constexpr int STAGES = NSTAGES, STAGE_BYTES = 2 * B_OFF;
    extern __shared__ unsigned char smem_raw[];
    // swizzled tiles need 1024-byte alignment: [stage: A hi | A lo | B hi | B lo] x STAGES, then the barriers
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);  // TMA landed          -> split warps
    uint64_t* ready_bar = full_bar + STAGES;                                         // hi/lo written       -> MMA issuer
    uint64_t* empty_bar = ready_bar + STAGES;                                        // MMAs retired        -> TMA producer
    uint64_t* acc_full_bar = empty_bar + STAGES;                                     // accumulator final   -> epilogue
    uint64_t* acc_empty_bar = acc_full_bar + ACC_BUFS;                               // accumulator drained -> MMA issuer
    uint64_t* tile_full_bar = acc_empty_bar + ACC_BUFS;                              // tile number written -> MMA issuer, accumulate and split warps
    uint64_t* tile_empty_bar = tile_full_bar + TILE_RING;                            // ... read by all of them -> producer
    int* tile_ring = reinterpret_cast<int*>(tile_empty_bar + TILE_RING);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ring + TILE_RING);
    static_assert((3 * NSTAGES + 2 * ACC_BUFS + 2 * TILE_RING) * 8 + TILE_RING * 4 + 4 <= 256, "barrier block");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&ready_bar[s], SPLIT_WARPS);
            ptx::mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int b = 0; b < ACC_BUFS; ++b) {
            ptx::mbar_init(&acc_full_bar[b], 1);
            ptx::mbar_init(&acc_empty_bar[b], 4);
        }
#pragma unroll
        for (int r = 0; r < TILE_RING; ++r) {
            ptx::mbar_init(&tile_full_bar[r], 1);
            ptx::mbar_init(&tile_empty_bar[r], 1 + 4 + (p.terms == 3 ? SPLIT_WARPS : 0));  // MMA lane + accumulate warps + split warps
        }
        ptx::fence_mbar_init();
    }
    if (warp == WARP_TMEM) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.tiles_m * p.tiles_n;
    const int kblocks = (p.k + p.bk - 1) >> (p.bk == 64 ? 6 : 5);  // bk is 32 (FP32 operands) or 64 (the 16-bit variant); a shift is cheap enough to recompute per role
    // consumer side of the tile ring: every role but the producer reads the same sequence of tile numbers, -1 last.  One counter per thread
    // carries both the slot (low bits) and the phase (next bit): the accumulate warps have no register to spare.
    static_assert((TILE_RING & (TILE_RING - 1)) == 0 && (ACC_BUFS & (ACC_BUFS - 1)) == 0, "ring sizes are powers of two");
    uint32_t tr_count = 0;
    auto next_tile_warp = [&]() -> int {  // whole warp, converged
        const uint32_t slot = tr_count & (TILE_RING - 1);
        tc::mbar_wait_guarded(&tile_full_bar[slot], (tr_count / TILE_RING) & 1);
        // (the shuffles here and below tell the compiler that the tile number - and with it every loop that runs on it - is warp-uniform:
        //  without them the pipeline bookkeeping of all roles leaves the uniform datapath; measured on the plain-TF32 mode: 221 instead of 240 TF at 8192^3)
        const int t = __shfl_sync(0xffffffffu, tile_ring[slot], 0);
        if (lane == 0) ptx::mbar_arrive(&tile_empty_bar[slot]);
        ++tr_count;
        return t;
    };
    auto next_tile_lane = [&]() -> int {  // a single lane
        const uint32_t slot = tr_count & (TILE_RING - 1);
        tc::mbar_wait_guarded(&tile_full_bar[slot], (tr_count / TILE_RING) & 1);
        const int t = __shfl_sync(0x1u, tile_ring[slot], 0);
        ptx::mbar_arrive(&tile_empty_bar[slot]);
        ++tr_count;
        return t;
    };

    // Every role starts with its share of the warpgroup-wide register reallocation (setmaxnreg): the control and split
    // warpgroups hand registers to the accumulate/epilogue warpgroup.
    if (warp == WARP_TMA) {
        // ===== TMA producer =====
        ptx::setmaxnreg_dec<REGS_CONTROL>();
        if (lane == 0) {
            ptx::prefetch_tensormap(&tmap_a);
            ptx::prefetch_tensormap(&tmap_b);
            int stage = 0;
            uint32_t phase = 0;
            auto draw = [&](int after) -> int {  // the next tile of this CTA, -1 when the launch has none left
                const int t = p.sched ? atomicAdd(&p.sched[0], 1) : (after < 0 ? (int)blockIdx.x : after + (int)gridDim.x);
                return __shfl_sync(0x1u, t < total_tiles ? t : -1, 0);
            };
            int tile = draw(-1);
            for (;;) {
                const uint32_t slot = tr_count & (TILE_RING - 1);
                tc::mbar_wait_guarded(&tile_empty_bar[slot], ((tr_count / TILE_RING) & 1) ^ 1);
                tile_ring[slot] = tile;
                ptx::mbar_arrive(&tile_full_bar[slot]);
                ++tr_count;
                if (tile < 0) break;
                const int next = draw(tile);  // issued ahead of this tile's loads: the atomic's round trip hides behind them
                int tm, tn;
                tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
                tile = next;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait_guarded(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * OPERAND_BYTES);
                    unsigned char* sa = base + stage * STAGE_BYTES;
                    unsigned char* sb = sa + B_OFF;
                    if (p.a_mn_major) {
#pragma unroll
                        for (int j = 0; j < BM / ATOM_MN; ++j) ptx::tma_load_2d(sa + j * MN_BOX_BYTES, &tmap_a, &full_bar[stage], tm * BM + j * ATOM_MN, kb * p.bk);
                    } else {
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * p.bk, tm * BM);
                    }
                    if (p.b_mn_major) {
#pragma unroll
                        for (int j = 0; j < BN / ATOM_MN; ++j) ptx::tma_load_2d(sb + j * MN_BOX_BYTES, &tmap_b, &full_bar[stage], tn * BN + j * ATOM_MN, kb * p.bk);
                    } else {
                        ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * p.bk, tn * BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (p.sched) {
                // every CTA draws exactly one number past the end; the CTA that reports it last knows that nobody will touch the pair again
                if (atomicAdd(&p.sched[1], 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; }
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ===== MMA issuer =====
        ptx::setmaxnreg_dec<REGS_CONTROL>();
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            const bool split = p.terms == 3;
            while (next_tile_lane() >= 0) {
                uint32_t d_tmem = 0;
                for (int kb = 0, wk = 0; kb < kblocks; ++kb) {
                    if (wk == 0) {  // open a window: its TMEM accumulator must have been drained
                        tc::mbar_wait_guarded(&acc_empty_bar[acc], acc_phase ^ 1);
                        tc::fence_after_thread_sync();
                        d_tmem = tmem_base + acc * BN;
                    }
                    tc::mbar_wait_guarded(split ? &ready_bar[stage] : &full_bar[stage], phase);
                    tc::fence_after_thread_sync();
                    const uint32_t a_hi = ptx::smem_u32(base + stage * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + OPERAND_BYTES;
                    const uint32_t b_hi = a_hi + B_OFF;
                    const uint32_t b_lo = b_hi + OPERAND_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                        const uint32_t oa = ks * p.kstep_a, ob = ks * p.kstep_b;
                        const uint32_t first = (wk | ks) ? 1u : 0u;  // the first MMA of a window overwrites
                        const uint64_t da_hi = tc::smem_desc(p.desc_a, a_hi + oa), db_hi = tc::smem_desc(p.desc_b, b_hi + ob);
                        if (split) {
                            const uint64_t da_lo = tc::smem_desc(p.desc_a, a_lo + oa), db_lo = tc::smem_desc(p.desc_b, b_lo + ob);
                            tc::mma_tf32(d_tmem, da_lo, db_hi, p.idesc, first);  // small terms first
                            tc::mma_tf32(d_tmem, da_hi, db_lo, p.idesc, 1u);
                            tc::mma_tf32(d_tmem, da_hi, db_hi, p.idesc, 1u);
                        } else {
                            if (p.f16_kind) tc::mma_f16(d_tmem, da_hi, db_hi, p.idesc, first);
                            else tc::mma_tf32(d_tmem, da_hi, db_hi, p.idesc, first);
                        }
                    }
                    tc::mma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++wk == p.window || kb == kblocks - 1) {  // close the window: hand its partial sum to the accumulate warps
                        tc::mma_commit(&acc_full_bar[acc]);
                        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
                        wk = 0;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= WARP_EPI0 && warp < WARP_EPI0 + 4) {
        // ===== accumulate + epilogue: warp q owns TMEM lanes 32q .. 32q+31 = rows 32q + lane of the tile =====
        ptx::setmaxnreg_inc<REGS_EPILOGUE>();
        const int q = warp - WARP_EPI0;
        uint32_t acc_count = 0;  // window accumulator = low bits, phase = the next bit
        const int windows = (kblocks + p.window - 1) / p.window;
        for (int tile = next_tile_warp(); tile >= 0; tile = next_tile_warp()) {
            int tm, tn;
            tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.f;
            for (int w = 0; w < windows; ++w) {
                const uint32_t acc = acc_count & (ACC_BUFS - 1);
                tc::mbar_wait_guarded(&acc_full_bar[acc], (acc_count / ACC_BUFS) & 1);
                tc::fence_after_thread_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll
                for (int h = 0; h < BN / 64; ++h) {
                    uint32_t v0[32], v1[32];
                    tc::tmem_ld_32x32b_x32(taddr + h * 64, v0);
                    tc::tmem_ld_32x32b_x32(taddr + h * 64 + 32, v1);
                    tc::tmem_ld_wait();
                    if (h == BN / 64 - 1) {  // window accumulator fully read: hand it back to the MMA issuer
                        tc::fence_before_thread_sync();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&acc_empty_bar[acc]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        sum[h * 64 + j] += __uint_as_float(v0[j]);
                        sum[h * 64 + 32 + j] += __uint_as_float(v1[j]);
                    }
                }
                ++acc_count;
            }
            const int row = tm * BM + q * 32 + lane;
            const bool row_ok = row < p.m;
            const int col0 = tn * BN;
            float* cp = p.c + (int64_t)col0 * p.ldc + row;
            if (p.read_c) {
#pragma unroll
                for (int cb = 0; cb < BN / 32; ++cb) {
                    float old[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) old[j] = (row_ok && col0 + cb * 32 + j < p.n) ? __ldcs(cp + (int64_t)(cb * 32 + j) * p.ldc) : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (row_ok && col0 + cb * 32 + j < p.n) cp[(int64_t)(cb * 32 + j) * p.ldc] = p.alpha * sum[cb * 32 + j] + p.beta * old[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < BN; ++j)
                    if (row_ok && col0 + j < p.n) cp[(int64_t)j * p.ldc] = p.alpha * sum[j];
            }
        }
    } else if (warp < WARP_EPI0) {
        ptx::setmaxnreg_dec<REGS_CONTROL>();  // TMEM warp and the spare warp of the control warpgroup
    } else if (p.terms != 3) {
        ptx::setmaxnreg_dec<REGS_SPLIT>();    // plain TF32: the tensor core reads the raw tiles, nothing to split
    } else {
        // ===== split stage: raw FP32 -> (hi in place, lo in the twin buffer), 16 bytes per access =====
        ptx::setmaxnreg_dec<REGS_SPLIT>();
        const int t = threadIdx.x - WARP_SPLIT0 * 32;
        constexpr int CHUNKS = 2 * OPERAND_BYTES / 16;  // A and B raw tiles
        constexpr int PER_THREAD = CHUNKS / (SPLIT_WARPS * 32);
        static_assert(CHUNKS % (SPLIT_WARPS * 32) == 0, "chunk split");
        int stage = 0;
        uint32_t phase = 0;
        while (next_tile_warp() >= 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                tc::mbar_wait_guarded(&full_bar[stage], phase);
                unsigned char* st = base + stage * STAGE_BYTES;
                float4 x[PER_THREAD];
#pragma unroll
                for (int i = 0; i < PER_THREAD; ++i) {
                    const int idx = t + i * (SPLIT_WARPS * 32);
                    const int off = idx * 16 + (idx >= OPERAND_BYTES / 16 ? OPERAND_BYTES : 0);  // skip over A lo
                    x[i] = *reinterpret_cast<const float4*>(st + off);
                }
                if (p.split_trunc) {
#pragma unroll
                    for (int i = 0; i < PER_THREAD; ++i) {
                        const int idx = t + i * (SPLIT_WARPS * 32);
                        const int off = idx * 16 + (idx >= OPERAND_BYTES / 16 ? OPERAND_BYTES : 0);
                        *reinterpret_cast<float4*>(st + off + OPERAND_BYTES) =
                            make_float4(lo_of_truncated(x[i].x), lo_of_truncated(x[i].y), lo_of_truncated(x[i].z), lo_of_truncated(x[i].w));
                    }
                } else
#pragma unroll
                for (int i = 0; i < PER_THREAD; ++i) {
                    const int idx = t + i * (SPLIT_WARPS * 32);
                    const int off = idx * 16 + (idx >= OPERAND_BYTES / 16 ? OPERAND_BYTES : 0);
                    const float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
                    float hi[4], lo[4];
                    split_tf32_x4(v, hi, lo);
                    *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(st + off + OPERAND_BYTES) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&ready_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == WARP_TMEM) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


__global__ void __launch_bounds__(THREADS, 1)
sgemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    sgemm_tc_body<STAGES, 2 * OPERAND_BYTES>(tmap_a, tmap_b, p);
}

static CUresult make_map(CUtensorMap* map, const float* base, uint64_t dim0, uint64_t dim1, uint64_t ld_elems, uint32_t box0, uint32_t box1,
                         CUtensorMapSwizzle swizzle) {
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(tensormap_encode_fn());
    if (!encode) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld_elems * 4};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// One counter pair per launch in flight, from a zeroed per-device array taken round robin (the kernel's last CTA zeroes its pair again).
// nullptr (static stride) when TMM_TC_SCHED=static or the array cannot be had.
static int* sched_slot() {
    static std::mutex mu;
    static int* base[64] = {};
    static bool tried[64] = {};
    static unsigned next[64] = {};
    static const bool is_static = [] { const char* v = getenv("TMM_TC_SCHED"); return v && v[0] == 's'; }();
    if (is_static) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried[dev]) {
        tried[dev] = true;
        void* q = nullptr;
        if (cudaMalloc(&q, SCHED_SLOTS * 2 * sizeof(int)) == cudaSuccess && cudaMemset(q, 0, SCHED_SLOTS * 2 * sizeof(int)) == cudaSuccess) base[dev] = static_cast<int*>(q);
        else cudaGetLastError();
    }
    return base[dev] ? base[dev] + 2 * (next[dev]++ % SCHED_SLOTS) : nullptr;
}

// developer override of the MN-major descriptor fields (tools/tc_test.cu sweeps them on a new driver / chip)
static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    return v && *v ? (uint32_t)strtoul(v, nullptr, 0) : dflt;
}

}  // namespace f32tc

bool sgemm_tc_eligible(const void* a, int64_t lda, const void* b, int64_t ldb) {
    return !((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (lda & 3) || (ldb & 3));
}

cudaError_t sgemm_tc_launch(char ta, char tb, int m, int n, int k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta,
                            float* c, int64_t ldc, cudaStream_t stream, int terms) {
    using namespace f32tc;
    if (m <= 0 || n <= 0) return cudaSuccess;
    if (!sgemm_tc_eligible(a, lda, b, ldb)) return cudaErrorInvalidValue;
    const bool a_mn = (ta == 'N'), b_mn = (tb != 'N');
    CUtensorMap map_a, map_b;
    CUresult r;
    const CUtensorMapSwizzle swz_k = CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapSwizzle swz_mn = (CUtensorMapSwizzle)env_u32("TMM_TC_MN_SWIZZLE", (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    // A: op(A) is m x k.  N: stored m x k (m contiguous) -> boxes [32 m x BK];  T/C: stored k x m (k contiguous) -> one box [BK x BM]
    r = a_mn ? make_map(&map_a, a, (uint64_t)m, (uint64_t)k, (uint64_t)lda, ATOM_MN, BK, swz_mn)
             : make_map(&map_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, BK, BM, swz_k);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(A, f32) failed: %d\n", (int)r); return cudaErrorInvalidValue; }
    // B: op(B) is k x n.  N: stored k x n (k contiguous) -> one box [BK x BN];  T/C: stored n x k (n contiguous) -> boxes [32 n x BK]
    r = b_mn ? make_map(&map_b, b, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, ATOM_MN, BK, swz_mn)
             : make_map(&map_b, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, BK, BN, swz_k);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[tiled-mm_b200] cuTensorMapEncodeTiled(B, f32) failed: %d\n", (int)r); return cudaErrorInvalidValue; }

    Params p;
    p.c = c; p.ldc = ldc; p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
    p.read_c = (beta != 0.f);
    p.tiles_m = (m + BM - 1) / BM; p.tiles_n = (n + BN - 1) / BN;
    p.a_mn_major = a_mn; p.b_mn_major = b_mn;
    // K-major tile: rows of 128 B (BK floats), 8-row swizzle atoms 1024 B apart; the next UMMA_K slice is 32 B further along the row.
    const uint64_t desc_k = tc::smem_desc_template(16, 8 * BK * 4, tc::LAYOUT_SW128);
    // MN-major tile: per k one 128-B row of 32 floats; 4-row atoms 512 B apart (stride offset), the next 32 rows/columns of
    // the tile in the next TMA box (leading offset); a UMMA_K slice is 8 k-rows = 1024 B.
    const uint64_t desc_mn = tc::smem_desc_template(env_u32("TMM_TC_MN_LBO", MN_BOX_BYTES), env_u32("TMM_TC_MN_SBO", 4 * ATOM_MN * 4),
                                                    env_u32("TMM_TC_MN_LAYOUT", tc::LAYOUT_SW128_ATOM32B));
    const uint32_t kstep_mn = env_u32("TMM_TC_MN_KSTEP", UMMA_K * ATOM_MN * 4);
    p.desc_a = a_mn ? desc_mn : desc_k; p.kstep_a = a_mn ? kstep_mn : UMMA_K * 4;
    p.desc_b = b_mn ? desc_mn : desc_k; p.kstep_b = b_mn ? kstep_mn : UMMA_K * 4;
    p.idesc = tc::instr_desc(tc::FMT_TF32, BM, BN, a_mn, b_mn);
    p.terms = terms == 1 ? 1 : 3;
    p.bk = BK; p.f16_kind = 0;
    p.window = (int)env_u32("TMM_TC_WINDOW", WINDOW_KBLOCKS);
    if (p.window < 1) p.window = 1;
    {
        const char* sv = getenv("TMM_TC_SPLIT");  // "trunc": hi = the raw bits (the tensor core ignores the low 13 mantissa bits), only lo is written: 171 vs 155 TF at
                                                  // 8192^3, representation error 2x that of the round-to-nearest split (profiles/r2_tc_variants.txt); opt-in
        p.split_trunc = (sv && sv[0] == 't') ? 1 : 0;
    }

    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
    if (tiles > INT32_MAX) return cudaErrorInvalidValue;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    p.sched = sched_slot();
    sgemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map_a, map_b, p);
    count_launch();
    return cudaGetLastError();
}

// Opt-in (TMM_BF16_NATIVE=1; green on hardware since round 2, tests/test_experimental_gpu.py::test_bf16_native_kind_f16_tn): bf16 operands straight through kind::f16 MMAs for the "TN" case, where both
// operands are k-contiguous.  A 128 x 64 bf16 tile has exactly the byte geometry of the 128 x 32 FP32 K-major tile above (128-byte rows,
// SWIZZLE_128B, 32 bytes per UMMA_K slice - 16 bf16 instead of 8 tf32), so the same kernel runs it with a BF16 tensor map, bk = 64 and
// the f16 instruction kind; no widening pass, twice the MMA rate.  Other op pairs keep the widening path of gemm_bf16_tc.cu.
cudaError_t bgemm_tc_native_tn_launch(int m, int n, int k, float alpha, const void* a, int64_t lda, const void* b, int64_t ldb, float beta, float* c,
                                      int64_t ldc, cudaStream_t stream) {
    using namespace f32tc;
    if (m <= 0 || n <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (lda & 7) || (ldb & 7)) return cudaErrorInvalidValue;
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(
        tensormap_encode_fn());
    if (!encode) return cudaErrorInvalidValue;
    constexpr int BK16 = 64;
    auto make = [&](CUtensorMap* map, const void* base, uint64_t rows_k, uint64_t cols, uint64_t ld) {
        cuuint64_t dims[2] = {rows_k, cols};
        cuuint64_t strides[1] = {ld * 2};
        cuuint32_t box[2] = {BK16, BM};
        cuuint32_t estr[2] = {1, 1};
        return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap map_a, map_b;
    if (make(&map_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda) != CUDA_SUCCESS || make(&map_b, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    Params p;
    p.c = c; p.ldc = ldc; p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
    p.read_c = (beta != 0.f);
    p.tiles_m = (m + BM - 1) / BM; p.tiles_n = (n + BN - 1) / BN;
    p.a_mn_major = 0; p.b_mn_major = 0;
    p.desc_a = p.desc_b = tc::smem_desc_template(16, 8 * 128, tc::LAYOUT_SW128);
    p.kstep_a = p.kstep_b = 32;  // 16 bf16
    p.idesc = tc::instr_desc(tc::FMT_BF16, BM, BN, false, false);
    p.terms = 1; p.window = WINDOW_KBLOCKS; p.split_trunc = 0;
    p.bk = BK16; p.f16_kind = 1;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
    if (tiles > INT32_MAX) return cudaErrorInvalidValue;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    p.sched = sched_slot();
    sgemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map_a, map_b, p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace tmm
