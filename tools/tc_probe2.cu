// Unit probes of the tcgen05 features the experimental kernels rely on (A operand from tensor memory, tcgen05.st, kind::i8), each in
// isolation: one CTA, one MMA, operands written by hand - so that if sgemm_tc_ts_kernel / dgemm_i8_kernel misbehave on hardware, one short
// GPU call tells WHICH assumption is wrong.  Development tool (GPU box):  make -C tools && ./build/tc_probe2
//   1  tcgen05.st -> tcgen05.ld round trip (lane = thread of the warp's lane quarter, register j = column j)
//   2  kind::tf32 MMA, A from shared memory (the hardware-validated form: checks this probe's own hand-written SWIZZLE_128B tiles)
//   3  kind::tf32 MMA, A from TENSOR MEMORY ([a_tmem] form; lane = row, one 32-bit column per k)
//   4  kind::i8 MMA, both operands from shared memory, int32 accumulator
//   5, 6, 7  tf32 (A from shared / tensor memory) and i8 MMAs on a CTA pair (cta_group::2, M = 256): see pair_probe_kernel
// Every wait is guarded (trap after ~10 s).
#include "../tiled-mm_b200/csrc/tmm_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace tmm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

constexpr int ROWS = 128, ROW_BYTES = 128, TILE_BYTES = ROWS * ROW_BYTES;  // one K-major SWIZZLE_128B tile: 128 rows x 128 bytes
constexpr int TMEM_COLS = 256, A_COL = 128;                                // D in columns 0..127, A (TMEM) from column 128

// byte offset of byte b of row r in a K-major SWIZZLE_128B tile (1024-byte aligned): the 16-byte chunk index is xor-ed with r & 7
__host__ __device__ inline int sw128(int r, int b) { return r * ROW_BYTES + ((((b >> 4) ^ (r & 7)) << 4) | (b & 15)); }

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(desc_a), "l"(desc_b),
                 "r"(idesc), "r"(accumulate)
                 : "memory");
}

// a_tile / b_tile: 16 KB images of the swizzled tiles (modes 2-4); a_rows: [128][16] 32-bit values for tcgen05.st (modes 1, 3); out: [128][128] 32-bit
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, const unsigned char* a_tile, const unsigned char* b_tile, const uint32_t* a_rows, uint32_t* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sa = base;
    unsigned char* sb = base + TILE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 2 * TILE_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(slot, TMEM_COLS);
    for (int i = t; i < TILE_BYTES / 16; i += 128) {
        reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_tile)[i];
        reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_tile)[i];
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);

    if (mode == 1 || mode == 3) {  // thread t = TMEM lane t writes its 16 values to columns A_COL .. A_COL + 15
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = a_rows[t * 16 + j];
        tc::tmem_st_32x32b_x16(lane_base + A_COL, v);
        tc::tmem_st_wait();
        tc::fence_before_thread_sync();
    }
    __syncthreads();
    tc::fence_after_thread_sync();

    if (mode >= 2) {
        if (t == 0) {
            const uint64_t dtmpl = tc::smem_desc_template(16, 8 * ROW_BYTES, tc::LAYOUT_SW128);
            const uint64_t da = tc::smem_desc(dtmpl, ptx::smem_u32(sa)), db = tc::smem_desc(dtmpl, ptx::smem_u32(sb));
            if (mode == 2) tc::mma_tf32(tmem, da, db, tc::instr_desc(tc::FMT_TF32, 128, 128, false, false), 0u);
            if (mode == 3) tc::mma_tf32_ts(tmem, tmem + A_COL, db, tc::instr_desc(tc::FMT_TF32, 128, 128, false, false), 0u);
            if (mode == 4) mma_i8(tmem, da, db, (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24), 0u);
            tc::mma_commit(bar);
        }
        tc::mbar_wait_guarded(bar, 0);
        tc::fence_after_thread_sync();
    }
    // read back: mode 1 -> the 32 columns from A_COL; otherwise D columns 0..127
#pragma unroll 1
    for (int cb = 0; cb < (mode == 1 ? 1 : 4); ++cb) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(lane_base + (mode == 1 ? A_COL : cb * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[t * 128 + cb * 32 + j] = v[j];
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) { tc::fence_after_thread_sync(); tc::tmem_dealloc(tmem, TMEM_COLS); }
}

// ---- CTA-pair probes (cluster of two): M = 256, N = 128, one cta_group::2 MMA issued by rank 0 -------------------------------------------
//   5  kind::tf32, A and B from shared memory   6  kind::tf32, A from tensor memory   (each CTA: its 128 rows of A, its 64-row half of B)
//   7  kind::i8, M = 256, N = 256 (the shape of igemm_group_kernel): each CTA its 128 rows of A and its 128-row half of B, 256 int32 columns
// Exercises exactly the pieces sgemm_tc_ts_kernel<true> / igemm_group_kernel add: cta_group::2 TMEM allocation, remote arrive on rank 0's
// barrier + cluster-scope acquire wait, the pair MMA reading both CTAs' operands, the multicast commit, cluster barriers.
__device__ __forceinline__ void mma_i8_pair_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(desc_a), "l"(desc_b),
                 "r"(idesc), "r"(accumulate)
                 : "memory");
}
constexpr int PAIR_TMEM_COLS = 512, PAIR_A_COL = 256;  // D in columns 0..255 at most, A (TMEM) from column 256
__device__ __forceinline__ void mma_tf32_pair_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(desc_a), "l"(desc_b),
                 "r"(idesc), "r"(accumulate)
                 : "memory");
}

// a_tiles: two 16 KB images (rank 0, rank 1); b_tiles: two 16 KB images holding 64 (modes 5, 6) or 128 (mode 7) rows each; a_rows: [256][16]; out: [256][N]
__global__ void __launch_bounds__(128, 1) pair_probe_kernel(int mode, const unsigned char* a_tiles, const unsigned char* b_tiles, const uint32_t* a_rows, uint32_t* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sa = base;
    unsigned char* sb = base + TILE_BYTES;
    uint64_t* ready_bar = reinterpret_cast<uint64_t*>(base + 2 * TILE_BYTES);  // rank 0's: both CTAs' operands are in place
    uint64_t* done_bar = ready_bar + 1;                                         // every CTA's own: the MMA has completed (multicast commit)
    uint32_t* slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t rank = tc::cluster_ctarank();
    if (t == 0) { ptx::mbar_init(ready_bar, 2); ptx::mbar_init(done_bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc_pair(slot, PAIR_TMEM_COLS);
    for (int i = t; i < TILE_BYTES / 16; i += 128) {
        reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_tiles + rank * TILE_BYTES)[i];
        reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_tiles + rank * TILE_BYTES)[i];
    }
    tc::fence_proxy_async_all();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::cluster_sync();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    if (mode == 6) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = a_rows[(rank * 128 + t) * 16 + j];
        tc::tmem_st_32x32b_x16(lane_base + PAIR_A_COL, v);
        tc::tmem_st_wait();
        tc::fence_before_thread_sync();
    }
    __syncthreads();
    if (t == 0) tc::mbar_arrive_cluster(ready_bar, 0);
    if (rank == 0 && t == 0) {
        tc::mbar_wait_cluster_guarded(ready_bar, 0);
        tc::fence_after_thread_sync();
        const uint64_t dtmpl = tc::smem_desc_template(16, 8 * ROW_BYTES, tc::LAYOUT_SW128);
        const uint64_t da = tc::smem_desc(dtmpl, ptx::smem_u32(sa)), db = tc::smem_desc(dtmpl, ptx::smem_u32(sb));
        const uint32_t idesc = tc::instr_desc(tc::FMT_TF32, 256, 128, false, false);
        if (mode == 5) mma_tf32_pair_ss(tmem, da, db, idesc, 0u);
        else if (mode == 6) tc::mma_tf32_ts_pair(tmem, tmem + PAIR_A_COL, db, idesc, 0u);
        else mma_i8_pair_ss(tmem, da, db, (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24), 0u);
        tc::mma_commit_pair(done_bar, 3);
    }
    tc::mbar_wait_guarded(done_bar, 0);
    tc::fence_after_thread_sync();
    const int n_cols = mode == 7 ? 256 : 128;
#pragma unroll 1
    for (int cb = 0; cb < n_cols / 32; ++cb) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(lane_base + cb * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[(rank * 128 + t) * n_cols + cb * 32 + j] = v[j];
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::cluster_sync();
    if (warp == 0) { tc::fence_after_thread_sync(); tc::tmem_dealloc_pair(tmem, PAIR_TMEM_COLS); }
}

static int run_pair(int mode) {
    const bool i8 = mode == 7;
    const int K = i8 ? 32 : 8, N = i8 ? 256 : 128;
    std::vector<unsigned char> a_tiles(2 * TILE_BYTES, 0), b_tiles(2 * TILE_BYTES, 0);
    std::vector<uint32_t> a_rows(256 * 16, 0), out(256 * N, 0xDEADBEEFu);
    auto A = [&](int i, int k) { return (i * 3 + k * 5) % 11 - 5; };
    auto B = [&](int j, int k) { return (j * 7 + k * 2) % 9 - 4; };
    for (int i = 0; i < 256; ++i)
        for (int k = 0; k < K; ++k) {
            if (i8) { a_tiles[(i / 128) * TILE_BYTES + sw128(i % 128, k)] = (unsigned char)(signed char)A(i, k); continue; }
            float f = (float)A(i, k);
            memcpy(&a_tiles[(i / 128) * TILE_BYTES + sw128(i % 128, 4 * k)], &f, 4);
            memcpy(&a_rows[i * 16 + k], &f, 4);
        }
    for (int j = 0; j < N; ++j)
        for (int k = 0; k < K; ++k) {   // rank r holds columns (N / 2) r .. (N / 2) r + N / 2 - 1 of the tile
            if (i8) { b_tiles[(j / (N / 2)) * TILE_BYTES + sw128(j % (N / 2), k)] = (unsigned char)(signed char)B(j, k); continue; }
            float f = (float)B(j, k);
            memcpy(&b_tiles[(j / (N / 2)) * TILE_BYTES + sw128(j % (N / 2), 4 * k)], &f, 4);
        }
    unsigned char *da, *db; uint32_t *dr, *dout;
    CK(cudaMalloc(&da, a_tiles.size())); CK(cudaMalloc(&db, b_tiles.size())); CK(cudaMalloc(&dr, a_rows.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
    CK(cudaMemcpy(da, a_tiles.data(), a_tiles.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b_tiles.data(), b_tiles.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, a_rows.data(), a_rows.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dout, out.data(), out.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 2 * TILE_BYTES + 1024 + 64;
    CK(cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_probe_kernel, mode, (const unsigned char*)da, (const unsigned char*)db, (const uint32_t*)dr, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int wrong = 0, fi = -1, fj = -1;
    auto got = [&](int i, int j) { if (i8) return (double)(int)out[i * N + j]; float f; memcpy(&f, &out[i * N + j], 4); return (double)f; };
    auto want = [&](int i, int j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A(i, k) * B(j, k); return s; };
    for (int i = 0; i < 256; ++i)
        for (int j = 0; j < N; ++j)
            if (got(i, j) != want(i, j)) { if (!wrong) { fi = i; fj = j; } ++wrong; }
    printf("probe %d (CTA pair, M = 256, N = %d, %s, A from %s): %s", mode, N, i8 ? "i8" : "tf32", mode == 6 ? "tensor memory" : "shared memory", wrong ? "WRONG" : "OK");
    if (wrong) {
        printf("  %d wrong of %d, first at (%d,%d): got %g expected %g; wrong per CTA:", wrong, 256 * N, fi, fj, got(fi, fj), want(fi, fj));
        for (int r = 0; r < 2; ++r) { int c = 0; for (int i = 128 * r; i < 128 * r + 128; ++i) for (int j = 0; j < N; ++j) c += got(i, j) != want(i, j); printf(" rank %d: %d", r, c); }
    }
    printf("\n");
    cudaFree(da); cudaFree(db); cudaFree(dr); cudaFree(dout);
    return wrong ? 1 : 0;
}

static int run(int mode) {
    std::vector<unsigned char> a_tile(TILE_BYTES, 0), b_tile(TILE_BYTES, 0);
    std::vector<uint32_t> a_rows(128 * 16, 0), out(128 * 128, 0xDEADBEEFu);
    std::vector<double> expect(128 * 128, 0.0);
    // small integers: exact in TF32 and int8.  A[i][k], B[j][k]; K = 8 (tf32) or 32 (int8)
    const int K = mode == 4 ? 32 : 8;
    auto A = [&](int i, int k) { return (i * 3 + k * 5) % 11 - 5; };
    auto B = [&](int j, int k) { return (j * 7 + k * 2) % 9 - 4; };
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < K; ++k) {
            if (mode == 4) {
                a_tile[sw128(r, k)] = (unsigned char)(signed char)A(r, k);
                b_tile[sw128(r, k)] = (unsigned char)(signed char)B(r, k);
            } else {
                float fa = (float)A(r, k), fb = (float)B(r, k);
                memcpy(&a_tile[sw128(r, 4 * k)], &fa, 4);
                memcpy(&b_tile[sw128(r, 4 * k)], &fb, 4);
            }
        }
    for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 16; ++j) {
            if (mode == 1) a_rows[r * 16 + j] = (uint32_t)(r * 100 + j);
            else { float f = j < 8 ? (float)A(r, j) : 0.f; memcpy(&a_rows[r * 16 + j], &f, 4); }
        }
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A(i, k) * B(j, k);
            expect[i * 128 + j] = s;
        }
    unsigned char *da, *db; uint32_t *dr, *dout;
    CK(cudaMalloc(&da, TILE_BYTES)); CK(cudaMalloc(&db, TILE_BYTES)); CK(cudaMalloc(&dr, a_rows.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
    CK(cudaMemcpy(da, a_tile.data(), TILE_BYTES, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b_tile.data(), TILE_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, a_rows.data(), a_rows.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dout, out.data(), out.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 2 * TILE_BYTES + 1024 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(mode, da, db, dr, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int wrong = 0, fi = -1, fj = -1;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < (mode == 1 ? 16 : 128); ++j) {
            bool ok;
            if (mode == 1) ok = out[i * 128 + j] == (uint32_t)(i * 100 + j);
            else if (mode == 4) ok = (double)(int)out[i * 128 + j] == expect[i * 128 + j];
            else { float f; memcpy(&f, &out[i * 128 + j], 4); ok = (double)f == expect[i * 128 + j]; }
            if (!ok) { if (!wrong) { fi = i; fj = j; } ++wrong; }
        }
    const char* names[] = {"", "tcgen05.st -> tcgen05.ld", "tf32 MMA, A from shared memory", "tf32 MMA, A from tensor memory", "i8 MMA, int32 accumulator"};
    printf("probe %d (%s): %s", mode, names[mode], wrong ? "WRONG" : "OK");
    if (wrong) {
        printf("  %d wrong, first at (%d,%d): got 0x%08x", wrong, fi, fj, out[fi * 128 + fj]);
        if (mode != 1) printf(" expected %g; row %d got:", expect[fi * 128 + fj], fi);
        if (mode != 1) for (int j = 0; j < 8; ++j) { if (mode == 4) printf(" %d", (int)out[fi * 128 + j]); else { float f; memcpy(&f, &out[fi * 128 + j], 4); printf(" %g", f); } }
    }
    printf("\n");
    cudaFree(da); cudaFree(db); cudaFree(dr); cudaFree(dout);
    return wrong ? 1 : 0;
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 1) return atoi(argv[1]) >= 5 ? run_pair(atoi(argv[1])) : run(atoi(argv[1]));
    for (int mode = 1; mode <= 7; ++mode) {
        bad += mode >= 5 ? run_pair(mode) : run(mode);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("context lost after probe %d; run the rest one by one: tc_probe2 <n>\n", mode); return 3; }
    }
    return bad ? 3 : 0;
}
