// Context + tile scheduler + C ABI of tiled_mm_b200.
//
// Replaces, for the hot path gpu::make_context -> gpu::gemm:
//   mm_handle / gpu_context / device_buffer       (reference mm_handle.cpp, gpu_context.cpp, device_buffer.hpp)
//   gemm front end, round_robin, round_robin_without_copy_c, copy_tile_*   (reference tiled_mm.cpp:45-123,270-624)
//
// It is NOT the reference's pipeline.  The reference walks C tiles, re-sends the A and B tiles of every
// (m,n) tile and chains copy -> gemm on one stream per slot (tiled_mm.cpp:292-358).  Here:
//   * every A/B element crosses PCIe exactly once whenever A, B and C fit in HBM ("resident" regime):
//       phase 1: A streams in as k-chunks together with the first column block of B; each chunk is one
//                large GEMM launch accumulating into C[:, 0:n1]   (compute starts after the first small chunk)
//       phase 2: A is now resident; the remaining column blocks of B arrive one by one, each is multiplied
//                with full k in one launch and its C block streams back immediately (short D2H tail)
//   * otherwise ("streaming" regime, out-of-core): C super-blocks stay resident while k-chunks of A and B
//     flow through a ring of device slots; the kernel epilogue accumulates across chunks (beta' = 1 after the
//     first chunk, reference tiled_mm.cpp:309)
//   * H2D, compute (several streams, so the tail of one launch is back-filled by the next) and D2H run on
//     separate streams tied together by pre-allocated events; all offsets are 64-bit.
// User tile sizes / stream counts are hints (they only cap chunk sizes / stream counts), never results.
#include "tmm_internal.h"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <mutex>

namespace tmm {

static thread_local std::string g_last_error;
const char* last_error_cstr() { return g_last_error.c_str(); }

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    // reference util.hpp:13-19: print the CUDA error string to stderr, then "GPU ERROR"
    fprintf(stderr, "error: GPU API call : %s (%s)\n", cudaGetErrorString(e), what);
    int code = (e == cudaErrorMemoryAllocation) ? TMM_ERR_NOMEM
               : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice) ? TMM_ERR_NOGPU
                                                                                                                       : TMM_ERR_CUDA;
    return fail(code, "GPU ERROR: %s: %s", what, cudaGetErrorString(e));
}

int nccl_fail(int rc, const char* what) {
    const nccl::Api& n = nccl::api();
    const char* msg = (n.ok && n.GetErrorString) ? n.GetErrorString(rc) : "NCCL unavailable";
    fprintf(stderr, "error: NCCL call : %s (%s)\n", msg, what);
    return fail(TMM_ERR_CUDA, "GPU ERROR: %s: %s", what, msg);
}

}  // namespace tmm

using tmm::fail;
using tmm::cuda_fail;
using tmm::DevBuf;
using tmm::DeviceGuard;
#define CU(x) TMM_CU(x)

namespace {
int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }
}  // namespace

namespace {

struct Call {
    tmm_context* ctx;
    int dtype;
    size_t es;  // element size
    char ta, tb;
    int64_t m, n, k;
    const void *alpha, *beta;
    const char *a, *b;
    char* c;
    int64_t lda, ldb, ldc;
    bool copy_c_back, beta_nonzero;
    bool device_operands = false;  // a, b or c is a DEVICE pointer (additive: the reference takes host pointers only): copies infer their direction
    // stored shapes (reference tiled_mm.cpp:507-514)
    int64_t a_rows, a_cols, b_rows, b_cols;
    unsigned char one[16];  // scalar 1 of the dtype (beta' for k-chunks > 0, tiled_mm.cpp:309)
    unsigned char zero[16] = {0};  // scalar 0 (first launch of a block whose beta * C is added at the end)
    int64_t m_plan, n_plan;  // dims the schedule was planned for: == m, n on one GPU; the grid-wide maximum block dims on a GPU
                             // grid, where every rank must walk the same schedule so that the panel exchanges line up
};

void make_one(int dtype, unsigned char* out) {
    memset(out, 0, 16);
    if (dtype == TMM_F32 || dtype == TMM_C32) { float v = 1.f; memcpy(out, &v, 4); }
    else { double v = 1.0; memcpy(out, &v, 8); }
}

bool scalar_is_zero(int dtype, const void* p) {
    switch (dtype) {
    case TMM_F32: return *(const float*)p == 0.f;
    case TMM_F64: return *(const double*)p == 0.0;
    case TMM_C32: return ((const float*)p)[0] == 0.f && ((const float*)p)[1] == 0.f;
    default: return ((const double*)p)[0] == 0.0 && ((const double*)p)[1] == 0.0;
    }
}

// ---- copy helpers (64-bit offsets throughout; reference copy_tile_* tiled_mm.cpp:45-123) -----------
int h2d_2d(Call& cl, void* dst, int64_t dpitch_elems, const char* src, int64_t spitch_elems, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return TMM_OK;
    CU(cudaMemcpy2DAsync(dst, (size_t)dpitch_elems * cl.es, src, (size_t)spitch_elems * cl.es, (size_t)rows * cl.es, (size_t)cols,
                         cl.device_operands ? cudaMemcpyDefault : cudaMemcpyHostToDevice, st));
    cl.ctx->stats.h2d_bytes += (uint64_t)rows * cols * cl.es;
    cl.ctx->stats.h2d_copies++;
    return TMM_OK;
}
int d2h_2d(Call& cl, char* dst, int64_t dpitch_elems, const void* src, int64_t spitch_elems, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return TMM_OK;
    CU(cudaMemcpy2DAsync(dst, (size_t)dpitch_elems * cl.es, src, (size_t)spitch_elems * cl.es, (size_t)rows * cl.es, (size_t)cols,
                         cl.device_operands ? cudaMemcpyDefault : cudaMemcpyDeviceToHost, st));
    cl.ctx->stats.d2h_bytes += (uint64_t)rows * cols * cl.es;
    cl.ctx->stats.d2h_copies++;
    return TMM_OK;
}

// A sub-block of op(A) restricted to rows [i0,i0+mi) of op(A) and k range [p0,p0+kc): where it lives in the
// stored (untransposed) host matrix, and its stored extent.
struct Sub { int64_t row, col, rows, cols; };
Sub a_sub(const Call& cl, int64_t i0, int64_t mi, int64_t p0, int64_t kc) {
    return cl.ta == 'N' ? Sub{i0, p0, mi, kc} : Sub{p0, i0, kc, mi};
}
Sub b_sub(const Call& cl, int64_t p0, int64_t kc, int64_t j0, int64_t nj) {
    return cl.tb == 'N' ? Sub{p0, j0, kc, nj} : Sub{j0, p0, nj, kc};
}

// Bring the stored sub-block `s` of A (or B) to its device position.  One GPU: a single 2-D H2D copy on s_h2d.  On a GPU grid
// the sub-block is shared by the grid row (A) / grid column (B): this rank uploads only its 1/p share and the shares are
// all-gathered over NVLink on s_comm (tmm_dist.cu).  Callers order consumers after BOTH streams (panels_ready).
int fetch_a(Call& cl, char* dst, int64_t dpitch, const Sub& s, int ring_slot = -1) {
    const tmm::Grid& g = cl.ctx->grid;
    const char* src = cl.a + ((size_t)s.col * cl.lda + s.row) * cl.es;
    if (g.pc > 1) return tmm::dist_exchange(cl.ctx, cl.ctx->grid.rowl, cl.es, src, cl.lda, s.rows, s.cols, dst, dpitch, ring_slot);
    return h2d_2d(cl, dst, dpitch, src, cl.lda, s.rows, s.cols, cl.ctx->s_h2d);
}
int fetch_b(Call& cl, char* dst, int64_t dpitch, const Sub& s, int ring_slot = -1) {
    const tmm::Grid& g = cl.ctx->grid;
    const char* src = cl.b + ((size_t)s.col * cl.ldb + s.row) * cl.es;
    if (g.pr > 1) return tmm::dist_exchange(cl.ctx, cl.ctx->grid.coll, cl.es, src, cl.ldb, s.rows, s.cols, dst, dpitch, ring_slot);
    return h2d_2d(cl, dst, dpitch, src, cl.ldb, s.rows, s.cols, cl.ctx->s_h2d);
}
// make `consumer` wait for everything fetched so far
int panels_ready(Call& cl, const cudaStream_t* consumers, int n_consumers) {
    tmm_context* ctx = cl.ctx;
    cudaEvent_t ev;
    CU(ctx->get_event(&ev));
    CU(cudaEventRecord(ev, ctx->s_h2d));
    for (int i = 0; i < n_consumers; ++i) CU(cudaStreamWaitEvent(consumers[i], ev, 0));
    if (ctx->grid.active()) {
        // NCCL-staged links complete on s_comm; DMA-pushed links complete when the peers' arrival counters say so
        tmm::Grid& g = ctx->grid;
        if ((g.rowl.active() && !g.rowl.direct) || (g.coll.active() && !g.coll.direct)) {
            CU(ctx->get_event(&ev));
            CU(cudaEventRecord(ev, ctx->s_comm));
            for (int i = 0; i < n_consumers; ++i) CU(cudaStreamWaitEvent(consumers[i], ev, 0));
        }
        for (int i = 0; i < n_consumers; ++i) {
            int rc = tmm::link_wait(ctx, g.rowl, consumers[i]);
            if (!rc) rc = tmm::link_wait(ctx, g.coll, consumers[i]);
            if (rc) return rc;
        }
    }
    return TMM_OK;
}
int panels_ready(Call& cl, cudaStream_t consumer) { return panels_ready(cl, &consumer, 1); }

int launch_gemm(Call& cl, int64_t mi, int64_t nj, int64_t kc, const void* da, int64_t pa, const void* db, int64_t pb, const void* beta, void* dc,
                int64_t ldc, cudaStream_t st) {
    tmm_context* ctx = cl.ctx;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->profiling) {
        CU(ctx->get_timing_event(&t0));
        CU(ctx->get_timing_event(&t1));
        CU(cudaEventRecord(t0, st));
    }
    cudaError_t e = tmm::device_gemm(cl.dtype, cl.ta, cl.tb, mi, nj, kc, cl.alpha, da, pa, db, pb, beta, dc, ldc, st);
    if (e != cudaSuccess) return cuda_fail(e, "device_gemm");
    if (ctx->profiling) { CU(cudaEventRecord(t1, st)); ctx->gemm_events.push_back({t0, t1}); }
    return TMM_OK;
}

size_t device_budget(tmm_context* ctx) {
    if (ctx->budget_override) return ctx->budget_override;
    if (ctx->budget_cached) return ctx->budget_cached;  // refreshed whenever an allocation fails or the context grows full C
    size_t fr = 0, to = 0;
#ifndef TMM_EMULATED
    tmm::scratch_trim();  // cached operand-preparation scratch goes back to the driver before the free memory is read
#endif
    if (cudaMemGetInfo(&fr, &to) != cudaSuccess) return (size_t)8 << 30;
    size_t held = ctx->buf_a.cap + ctx->buf_b.cap + ctx->buf_c.cap + ctx->buf_cs.cap;
    double avail = (double)fr + (double)held;
    ctx->budget_cached = (size_t)(avail * 0.92);
    return ctx->budget_cached;
}

using tmm::TraceScope;

#ifndef TMM_EMULATED
// ---- opt-in FP64 emulation on the int8 tensor cores (TMM_F64_MATH=i8[:S], csrc/gemm_f64_i8.cu) inside the resident schedule ------------------
// Launch by launch the emulation would slice its operands again and again: a k-chunk of A meets up to four column stripes, and the resident
// A meets every phase-2 column block (9 slicing passes over A at 10000^3; host-to-host 69 ms against 58.6 ms with DMMA, although the kernel
// itself is 1.34 x faster - profiles/r2_f64_i8_first_run.txt).  Here every panel piece is sliced ONCE, where it lands, and the slice GEMMs
// run on the cached slices: per k-chunk the A chunk (all stripes share it) and each stripe's B chunk, then - for phase 2 - all of A over the
// full k (one more pass, its own row scales) and each column block of B.  Slices live in context-owned storage carved per call.
struct I8Cache {
    bool on = false;
    int slices = 0;
    char* q_base = nullptr; size_t q_off = 0, q_cap = 0;
    int* e_base = nullptr; size_t e_off = 0, e_cap = 0;
    std::vector<tmm::I8Slices> a_chunk;               // [chunk]
    std::vector<std::vector<tmm::I8Slices>> b_chunk;  // [chunk][stripe]
    tmm::I8Slices a_full;
    std::vector<tmm::I8Slices> b_block;               // [block]
    bool carve(tmm::I8Slices* out, int rows, int k) {
        out->rows = rows; out->k = k; out->slices = slices;
        const size_t bytes = tmm::i8_slices_layout(rows, k, slices, &out->rows_pad, &out->pitch);
        const size_t q_at = (q_off + 1023) / 1024 * 1024;
        if (q_at + bytes > q_cap || e_off + (size_t)rows > e_cap) return false;
        out->q = reinterpret_cast<int8_t*>(q_base + q_at); q_off = q_at + bytes;
        out->e = e_base + e_off; e_off += (size_t)rows;
        return true;
    }
};
#endif

// ------------------------------------------------------------------------------------------------
// Resident regime: device holds all of A, B and C.
// ------------------------------------------------------------------------------------------------
int run_resident(Call& cl, const tmm::Plan& pl, void* dC, int64_t ldc_dev, void* dCs) {
    tmm_context* ctx = cl.ctx;
    const size_t es = cl.es;
    const int64_t pa = pl.pitch_a, pb = pl.pitch_b;
    cudaError_t e;
    int alloc_rc = TMM_OK;
    if ((e = ctx->buf_a.reserve(pl.bytes_a)) != cudaSuccess) alloc_rc = cuda_fail(e, "cudaMalloc(A panels)");
    else if ((e = ctx->buf_b.reserve(pl.bytes_b)) != cudaSuccess) alloc_rc = cuda_fail(e, "cudaMalloc(B panels)");
    if (ctx->grid.active()) {
        int rc = tmm::grid_bind(ctx, alloc_rc);  // collective even when the allocation failed here: the peers must not start without us
        if (rc) return rc;
        for (void* old : ctx->retired) cudaFree(old);  // every peer has re-mapped: nobody holds the outgrown buffers any more
        ctx->retired.clear();
    } else if (alloc_rc) return alloc_rc;
    char* dA = (char*)ctx->buf_a.p;
    char* dB = (char*)ctx->buf_b.p;
    const int ncs = ctx->n_compute();
    const int64_t n1 = std::min<int64_t>(pl.n1, cl.n);  // the plan is for (m_plan, n_plan) >= (m, n): clamp to this rank's block
    ctx->stats.k_chunks = (int)pl.chunks.size();
    ctx->stats.c_blocks = 1 + (int)pl.blocks.size();
    // beta != 0 with a staging copy: every block starts its accumulation from zero (like beta == 0: no C upload in front of the first GEMM) and
    // beta * C is added from the staging copy once the block's last launch is done - the C uploads move behind the A / B panels that gate compute.
    // (Round 1 uploaded C[:, 0:n1] stripe by stripe in front of the first k-chunks: at dgemm 10000^3, beta = 1, the SMs idled ~8 ms early in the
    //  call and it took 67 ms against 58 ms for beta = 0 although its uploads, 46 ms, fit under the 56 ms of GEMM - profiles/r2_beta1_trace_before.txt.)
    const bool defer_c = cl.beta_nonzero && dCs != nullptr;
    const bool c_first = cl.beta_nonzero && !defer_c;

    cudaEvent_t ev;
    // ---- phase 1: A streams in as k-chunks with the first column block of B.  The block is cut into P1 column stripes, each a
    // chain of accumulating launches on its own high-priority stream: a chain is serial (chunk c+1 reads what chunk c wrote),
    // so a single chain drains the SMs at every launch boundary; with two or more staggered chains the block scheduler always
    // has CTAs of another stripe to dispatch while one stripe's launch drains.
    int P1 = (int)std::min<int64_t>(tmm_context::MAX_P1, std::max<int64_t>(1, n1 / 1024));
    {
        const char* v = getenv("TMM_PLAN_P1SPLIT");
        if (v && *v) P1 = std::max(1, std::min<int>(std::min<int64_t>(tmm_context::MAX_P1, (n1 + 63) / 64), atoi(v)));
    }
    const int n_chunks = (int)pl.chunks.size();
    std::vector<int64_t> chunk_p0(n_chunks + 1, 0);
    for (int ci = 0; ci < n_chunks; ++ci) chunk_p0[ci + 1] = chunk_p0[ci] + pl.chunks[ci];
    // NS column stripes over P1 streams (stripe j runs on stream j % P1).  Default: NS = P1 equal stripes.
    constexpr int MAX_STRIPES = 16;
    int NS = P1;
    int64_t s_off[MAX_STRIPES + 1];
    for (int s = 0; s <= NS; ++s) s_off[s] = s == NS ? n1 : std::min<int64_t>(n1, round_up(n1 * s / NS, 64));
    if (c_first) {
        // Experiment (TMM_PLAN_CSTRIPES=<count> | chunks; measured in round 2: 71.8 ms against 67.1 ms for the default stripes - not the default): with beta != 0 the work that is unlocked per
        // uploaded byte is largest when the columns of C arrive at the same pace as the k-columns of A, i.e. one stripe per k-chunk
        // with widths in proportion to the chunk widths, instead of a few fat stripes whose C delays the first chunks.
        const char* v = getenv("TMM_PLAN_CSTRIPES");
        if (v && *v) {
            const int want = (v[0] == 'c') ? n_chunks : atoi(v);
            const int ns = std::max(1, std::min(std::min(want, n_chunks), std::min<int>(MAX_STRIPES, (int)((n1 + 63) / 64))));
            if (ns > 1 && cl.k > 0) {
                NS = ns;
                P1 = std::min<int>(tmm_context::MAX_P1, NS);
                // stripe boundaries follow the k-chunk boundaries, scaled from [0, k) to [0, n1); the chunks beyond NS share the last stripe
                for (int s = 0; s <= NS; ++s) s_off[s] = s == NS ? n1 : std::min<int64_t>(n1, round_up((int64_t)((double)chunk_p0[s] / (double)chunk_p0[NS] * (double)n1), 64));
            }
        }
    }
#ifndef TMM_EMULATED
    I8Cache i8;
    if (cl.dtype == TMM_F64 && tmm::f64_i8_slices() > 0 && cl.k > 0) {
        // size the slice storage for this plan; if it cannot be had, every launch slices for itself (the plain path of device_gemm)
        i8.slices = tmm::f64_i8_slices();
        const int64_t m_pad = round_up(cl.m, 128);
        size_t q_need = 0, e_need = 0;
        int64_t k_pitch_sum = 0, n1_pad = 0;
        for (int ci = 0; ci < n_chunks; ++ci) k_pitch_sum += round_up(pl.chunks[ci], 128);
        for (int sidx = 0; sidx < NS; ++sidx) n1_pad += round_up(std::max<int64_t>(0, s_off[sidx + 1] - s_off[sidx]), 128);
        q_need += (size_t)i8.slices * (size_t)(m_pad + n1_pad) * (size_t)k_pitch_sum + (size_t)(n_chunks * (NS + 1)) * 1024;
        e_need += (size_t)n_chunks * (size_t)(cl.m + n1);
        if (!pl.blocks.empty() && n1 < cl.n) {
            int64_t nb_pad = 0;
            for (int64_t nb : pl.blocks) nb_pad += round_up(nb, 128);
            q_need += (size_t)i8.slices * (size_t)(m_pad + nb_pad) * (size_t)round_up(cl.k, 128) + (pl.blocks.size() + 1) * 1024;
            e_need += (size_t)cl.m + (size_t)(cl.n - n1);
        }
        size_t fr = 0, to = 0;
        const size_t held = ctx->i8_q.cap + ctx->i8_e.cap;
        if (cudaMemGetInfo(&fr, &to) == cudaSuccess && q_need + e_need * sizeof(int) <= fr + held - std::min<size_t>(fr + held, (size_t)1 << 30) &&
            ctx->i8_q.reserve(q_need) == cudaSuccess && ctx->i8_e.reserve(e_need * sizeof(int)) == cudaSuccess) {
            i8.on = true;
            i8.q_base = static_cast<char*>(ctx->i8_q.p); i8.q_cap = ctx->i8_q.cap;
            i8.e_base = static_cast<int*>(ctx->i8_e.p); i8.e_cap = ctx->i8_e.cap / sizeof(int);
            i8.a_chunk.resize(n_chunks);
            i8.b_chunk.assign(n_chunks, std::vector<tmm::I8Slices>(NS));
            i8.b_block.resize(pl.blocks.size());
            for (int ci = 0; ci < n_chunks && i8.on; ++ci) {
                i8.on = i8.carve(&i8.a_chunk[ci], (int)cl.m, (int)pl.chunks[ci]);
                for (int sidx = 0; sidx < NS && i8.on; ++sidx)
                    if (s_off[sidx + 1] > s_off[sidx]) i8.on = i8.carve(&i8.b_chunk[ci][sidx], (int)(s_off[sidx + 1] - s_off[sidx]), (int)pl.chunks[ci]);
            }
            if (i8.on && !pl.blocks.empty() && n1 < cl.n) {
                i8.on = i8.carve(&i8.a_full, (int)cl.m, (int)cl.k);
                int64_t jb = pl.n1;
                for (size_t blk = 0; blk < pl.blocks.size() && i8.on; ++blk) {
                    const int64_t nb = std::min<int64_t>(pl.blocks[blk], cl.n - jb);
                    if (nb <= 0) break;
                    i8.on = i8.carve(&i8.b_block[blk], (int)nb, (int)cl.k);
                    jb += nb;
                }
            }
        } else cudaGetLastError();
    }
    const double i8_alpha = *static_cast<const double*>(cl.alpha);
    // element (row i of op(A), k index l) of a device sub-block of A at base[i * stride_row + l * stride_k]; likewise columns of op(B)
    const int64_t a_sr = cl.ta == 'N' ? 1 : pa, a_sk = cl.ta == 'N' ? pa : 1, b_sr = cl.tb == 'N' ? pb : 1, b_sk = cl.tb == 'N' ? 1 : pb;
#endif
    // ---- complex<float> on the tcgen05 kernel (csrc/gemm_c32_tc.cu): operands prepared ONCE, where they land ---------------------------------
    // The kernel multiplies the real embedding A' (2m x 2k floats, alpha folded in) with B' (the stored B read as floats for op N, a split
    // copy for op T / C).  Launch by launch that meant a stream-ordered allocation and a pass over A for every stripe of every k-chunk and over
    // all of A for every phase-2 block; the allocations could block the enqueueing host thread for tens of milliseconds when the pool had to
    // wait for a block to come back (profiles/r2_split_fix.txt: cgemm 8000^3 30 ms in one process, 89 ms after other sizes had shaped the pool).
    // Here A' of the whole panel lives in a context buffer: each k-chunk is embedded once when it arrives (on the first stripe's stream, the
    // other stripe streams wait for it) and the chunks add up to the A' of the resident A that phase 2 multiplies; B pieces are split by the
    // launch that consumes them, into the matching position of a second context buffer.  No allocation on the launch path.
    struct C32Prep {
        bool on = false;
        float* a2 = nullptr; int64_t pitch_a2 = 0;
        float* b2 = nullptr; int64_t pitch_b2 = 0;
        cudaEvent_t last_embed = nullptr;
    } c32;
    if (cl.dtype == TMM_C32 && tmm::c32_math_mode() == TMM_CMATH_TC && tmm::f32_math_mode() != 0 && cl.k > 0 && cl.m > 0 && cl.n > 0 &&
        cl.m <= INT32_MAX / 2 && cl.k <= INT32_MAX / 2) {
        c32.pitch_a2 = cl.ta == 'N' ? round_up(2 * cl.m, 32) : round_up(2 * cl.k, 32);
        const size_t a2_bytes = (size_t)c32.pitch_a2 * (size_t)(2 * (cl.ta == 'N' ? cl.k : cl.m)) * sizeof(float);
        c32.pitch_b2 = cl.tb == 'N' ? 2 * pb : round_up(cl.n, 32);
        const size_t b2_bytes = cl.tb == 'N' ? 0 : (size_t)c32.pitch_b2 * (size_t)(2 * cl.k) * sizeof(float);
        if (ctx->c32_a2.reserve(a2_bytes) == cudaSuccess && (b2_bytes == 0 || ctx->c32_b2.reserve(b2_bytes) == cudaSuccess)) {
            c32.on = true;
            c32.a2 = static_cast<float*>(ctx->c32_a2.p);
            c32.b2 = static_cast<float*>(ctx->c32_b2.p);
        } else cudaGetLastError();  // no room: every launch prepares its own operands (device_gemm)
    }
    // the launch of one block on prepared operands: rows [0, m) of op(A), columns [j0, j0 + nj) of op(B), k range [p0, p0 + kc)
    auto c32_gemm = [&](int64_t j0, int64_t nj, int64_t p0, int64_t kc, const void* beta, void* dc, cudaStream_t st) -> int {
        const float* a2 = cl.ta == 'N' ? c32.a2 + (size_t)(2 * p0) * c32.pitch_a2 : c32.a2 + 2 * p0;
        const Sub sbs = b_sub(cl, p0, kc, j0, nj);
        const char* db = dB + ((size_t)sbs.col * pb + sbs.row) * es;
        const float* b2;
        if (cl.tb == 'N') b2 = reinterpret_cast<const float*>(db);  // zero copy: the stored k x n block read as (2k x n) floats
        else {
            float* out = c32.b2 + (size_t)(2 * p0) * c32.pitch_b2 + j0;
            cudaError_t e1 = tmm::cgemm_tc_split_b(cl.tb, (int)nj, (int)kc, db, pb, out, c32.pitch_b2, st);
            if (e1 != cudaSuccess) return cuda_fail(e1, "complex<float> operand split (B)");
            b2 = out;
        }
        cudaError_t e1 = tmm::cgemm_tc_prepared(cl.ta, cl.tb, (int)cl.m, (int)nj, (int)kc, a2, c32.pitch_a2, b2, c32.pitch_b2, static_cast<const float*>(beta), dc, ldc_dev, st);
        if (e1 == cudaErrorInvalidValue) {  // a sub-block off the 16-byte grid (odd chunk start): this launch prepares its own operands
            cudaGetLastError();
            const Sub sa = a_sub(cl, 0, cl.m, p0, kc);
            return launch_gemm(cl, cl.m, nj, kc, dA + ((size_t)sa.col * pa + sa.row) * es, pa, db, pb, beta, dc, ldc_dev, st);
        }
        return e1 == cudaSuccess ? TMM_OK : cuda_fail(e1, "complex<float> GEMM on prepared operands");
    };
    // beta != 0: host C is read (only then, reference tiled_mm.cpp:325).  Uploading the whole C[:, 0:n1] before the first chunk would
    // keep the SMs idle for |C block| / BW_pcie (8 ms at 10000^3); instead stripe s's C travels right before k-chunk s, so the first
    // chain starts after ONE stripe of C and the later stripes join one chunk apart, catching up on the chunks that arrived first.
    auto upload_c_stripe = [&](int sidx) -> int {
        const int64_t js = s_off[sidx], ws = s_off[sidx + 1] - js;
        if (ws <= 0) return TMM_OK;
        TraceScope ts(ctx, ctx->s_h2d, "h2dC", js, ws);
        return h2d_2d(cl, (char*)dC + (size_t)js * ldc_dev * es, ldc_dev, cl.c + (size_t)js * cl.ldc * es, cl.ldc, cl.m, ws, ctx->s_h2d);
    };
    auto launch_stripe_chunk = [&](int sidx, int ci) -> int {
        const int64_t js = s_off[sidx], ws = s_off[sidx + 1] - js;
        if (ws <= 0) return TMM_OK;
        const int64_t p0 = chunk_p0[ci], kc = pl.chunks[ci];
        const Sub sa = a_sub(cl, 0, cl.m, p0, kc), sbs = b_sub(cl, p0, kc, js, ws);
        TraceScope ts(ctx, ctx->s_p1[sidx % P1], "gemm1", js, ws, kc);
#ifndef TMM_EMULATED
        if (i8.on) {  // this stripe's B chunk is sliced here, once; the A chunk was sliced when it arrived (below)
            cudaStream_t st = ctx->s_p1[sidx % P1];
            cudaError_t e8 = tmm::i8_slice_operand(reinterpret_cast<const double*>(dB + ((size_t)sbs.col * pb + sbs.row) * es), b_sr, b_sk, (int)ws, (int)kc, i8.b_chunk[ci][sidx], st);
            if (e8 == cudaSuccess)
                e8 = tmm::i8_gemm_sliced(i8.a_chunk[ci], 0, (int)cl.m, i8.b_chunk[ci][sidx], 0, (int)ws, i8_alpha, ci == 0 ? (defer_c ? 0.0 : *static_cast<const double*>(cl.beta)) : 1.0,
                                         reinterpret_cast<double*>((char*)dC + (size_t)js * ldc_dev * es), ldc_dev, st);
            return e8 == cudaSuccess ? TMM_OK : cuda_fail(e8, "int8 slice GEMM (phase 1)");
        }
#endif
        const void* beta1 = ci == 0 ? (defer_c ? (const void*)cl.zero : cl.beta) : (const void*)cl.one;
        if (c32.on) return c32_gemm(js, ws, p0, kc, beta1, (char*)dC + (size_t)js * ldc_dev * es, ctx->s_p1[sidx % P1]);
        return launch_gemm(cl, cl.m, ws, kc, dA + ((size_t)sa.col * pa + sa.row) * es, pa, dB + ((size_t)sbs.col * pb + sbs.row) * es, pb,
                           beta1, (char*)dC + (size_t)js * ldc_dev * es, ldc_dev, ctx->s_p1[sidx % P1]);
    };
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int64_t p0 = chunk_p0[ci], kc = pl.chunks[ci];
        if (c_first && ci < NS) { int rc = upload_c_stripe(ci); if (rc) return rc; }
        const Sub sa = a_sub(cl, 0, cl.m, p0, kc), sb = b_sub(cl, p0, kc, 0, n1);
        {
            TraceScope ts(ctx, ctx->s_h2d, "h2dAB", p0, kc);
            int rc = fetch_a(cl, dA + ((size_t)sa.col * pa + sa.row) * es, pa, sa);
            if (rc) return rc;
            rc = fetch_b(cl, dB + ((size_t)sb.col * pb + sb.row) * es, pb, sb);
            if (rc) return rc;
        }
        {
            int rc = panels_ready(cl, ctx->s_p1, P1);
            if (rc) return rc;
        }
        if (c32.on) {  // the chunk of A that just arrived: embedded once on the first stripe's stream, the other stripe streams wait for it
            TraceScope ts(ctx, ctx->s_p1[0], "embedA", p0, kc);
            float* out = cl.ta == 'N' ? c32.a2 + (size_t)(2 * p0) * c32.pitch_a2 : c32.a2 + 2 * p0;
            cudaError_t e1 = tmm::cgemm_tc_embed_a(cl.ta, (int)cl.m, (int)kc, static_cast<const float*>(cl.alpha), dA + ((size_t)sa.col * pa + sa.row) * es, pa, out, c32.pitch_a2, ctx->s_p1[0]);
            if (e1 != cudaSuccess) return cuda_fail(e1, "complex<float> operand embedding (A chunk)");
            CU(ctx->get_event(&c32.last_embed));
            CU(cudaEventRecord(c32.last_embed, ctx->s_p1[0]));
            for (int j = 1; j < P1; ++j) CU(cudaStreamWaitEvent(ctx->s_p1[j], c32.last_embed, 0));
        }
#ifndef TMM_EMULATED
        if (i8.on) {  // the chunk of A that just arrived: sliced once on the first stripe's stream, the other stripe streams wait for it
            TraceScope ts(ctx, ctx->s_p1[0], "sliceA", p0, kc);
            cudaError_t e8 = tmm::i8_slice_operand(reinterpret_cast<const double*>(dA + ((size_t)sa.col * pa + sa.row) * es), a_sr, a_sk, (int)cl.m, (int)kc, i8.a_chunk[ci], ctx->s_p1[0]);
            if (e8 != cudaSuccess) return cuda_fail(e8, "int8 slicing of an A chunk");
            if (P1 > 1) {
                CU(ctx->get_event(&ev));
                CU(cudaEventRecord(ev, ctx->s_p1[0]));
                for (int j = 1; j < P1; ++j) CU(cudaStreamWaitEvent(ctx->s_p1[j], ev, 0));
            }
        }
#endif
        for (int s = 0; s < NS; ++s) {
            if (c_first && s > ci) continue;                              // this stripe's C has not been sent yet
            if (c_first && s == ci)                                       // it has now: catch up on the chunks that are already here
                for (int cj = 0; cj < ci; ++cj) { int rc = launch_stripe_chunk(s, cj); if (rc) return rc; }
            int rc = launch_stripe_chunk(s, ci);
            if (rc) return rc;
        }
    }
    if (c_first)
        for (int s = n_chunks; s < NS; ++s) {                             // fewer k-chunks than stripes: the remaining chains start here
            int rc = upload_c_stripe(s);
            if (!rc) rc = panels_ready(cl, &ctx->s_p1[s % P1], 1);
            for (int cj = 0; cj < n_chunks && !rc; ++cj) rc = launch_stripe_chunk(s, cj);
            if (rc) return rc;
        }
    // phase-2 geometry is needed here already: with deferred C the first column block's B is fetched BEFORE the C stripes of phase 1, so that the
    // block's GEMM can back-fill the SMs while those stripes travel
    bool first_block_prefetched = false;
    if (defer_c) {
        if (!pl.blocks.empty() && pl.n1 < cl.n) {
            const int64_t nb0 = std::min<int64_t>(pl.blocks[0], cl.n - pl.n1);
            if (nb0 > 0) {
                const Sub sb0 = b_sub(cl, 0, cl.k, pl.n1, nb0);
                TraceScope ts(ctx, ctx->s_h2d, "h2dB", pl.n1, nb0);
                int rc = fetch_b(cl, dB + ((size_t)sb0.col * pb + sb0.row) * es, pb, sb0);
                if (!rc) rc = panels_ready(cl, ctx->s_compute[1]);  // block 0's stream waits for the copies up to HERE, not for the C stripes that follow
                if (rc) return rc;
                first_block_prefetched = true;
            }
        }
        for (int sidx = 0; sidx < NS; ++sidx) {
            const int64_t js = s_off[sidx], ws = s_off[sidx + 1] - js;
            if (ws <= 0) continue;
            cudaStream_t st = ctx->s_p1[sidx % P1];
            {
                TraceScope ts(ctx, ctx->s_h2d, "h2dC", js, ws);
                int rc = h2d_2d(cl, (char*)dCs + (size_t)js * ldc_dev * es, ldc_dev, cl.c + (size_t)js * cl.ldc * es, cl.ldc, cl.m, ws, ctx->s_h2d);
                if (rc) return rc;
            }
            CU(ctx->get_event(&ev));
            CU(cudaEventRecord(ev, ctx->s_h2d));
            CU(cudaStreamWaitEvent(st, ev, 0));
            TraceScope ts(ctx, st, "addC", js, ws);
            cudaError_t e2 = tmm::device_add_scaled(cl.dtype, cl.m, ws, cl.beta, (char*)dCs + (size_t)js * ldc_dev * es, ldc_dev, (char*)dC + (size_t)js * ldc_dev * es, ldc_dev, st);
            if (e2 != cudaSuccess) return cuda_fail(e2, "C += beta * C_host");
        }
    }
    if (cl.copy_c_back) {
        // stripes finish in order; each leaves for the host as soon as its chain is done
        for (int s = 0; s < NS; ++s) {
            const int64_t js = s_off[s], ws = s_off[s + 1] - js;
            if (ws <= 0) continue;
            CU(ctx->get_event(&ev));
            CU(cudaEventRecord(ev, ctx->s_p1[s % P1]));
            CU(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
            const int64_t piece = std::max<int64_t>(64, round_up(ws / 2, 64));
            for (int64_t j = js; j < js + ws; j += piece) {
                const int64_t w = std::min(piece, js + ws - j);
                TraceScope ts(ctx, ctx->s_d2h, "d2hC", j, w);
                int rc = d2h_2d(cl, cl.c + (size_t)j * cl.ldc * es, cl.ldc, (char*)dC + (size_t)j * ldc_dev * es, ldc_dev, cl.m, w, ctx->s_d2h);
                if (rc) return rc;
            }
        }
    }
    // ---- phase 2: A resident; remaining column blocks of B, full k each, C block streams back at once
    int64_t j0 = pl.n1;
#ifndef TMM_EMULATED
    bool a_full_sliced = false;
    cudaEvent_t a_full_ev = nullptr;
#endif
    for (size_t blk = 0; blk < pl.blocks.size(); ++blk) {
        const int64_t nb = std::min<int64_t>(pl.blocks[blk], cl.n - j0);
        if (nb <= 0) break;  // this rank's block is narrower than the planned one (same for its whole grid column)
        char* dcb = (char*)dC + (size_t)j0 * ldc_dev * es;
        if (c_first) {
            TraceScope ts(ctx, ctx->s_h2d, "h2dC", j0, nb);
            int rc = h2d_2d(cl, dcb, ldc_dev, cl.c + (size_t)j0 * cl.ldc * es, cl.ldc, cl.m, nb, ctx->s_h2d);
            if (rc) return rc;
        }
        Sub sb = b_sub(cl, 0, cl.k, j0, nb);
        char* db = dB + ((size_t)sb.col * pb + sb.row) * es;
        if (!(blk == 0 && first_block_prefetched)) {
            TraceScope ts(ctx, ctx->s_h2d, "h2dB", j0, nb);
            int rc = fetch_b(cl, db, pb, sb);
            if (rc) return rc;
        }
        cudaStream_t cs = ctx->s_compute[1 + blk % (ncs - 1)];
        if (!(blk == 0 && first_block_prefetched)) {
            int rc = panels_ready(cl, cs);
            if (rc) return rc;
        }
#ifndef TMM_EMULATED
        if (i8.on) {
            if (!a_full_sliced) {  // all of A is on the device now (the copies this stream just waited for): one slicing pass over the full k
                TraceScope ts(ctx, cs, "sliceA", 0, cl.k);
                const Sub sfull = a_sub(cl, 0, cl.m, 0, cl.k);
                cudaError_t e8 = tmm::i8_slice_operand(reinterpret_cast<const double*>(dA + ((size_t)sfull.col * pa + sfull.row) * es), a_sr, a_sk, (int)cl.m, (int)cl.k, i8.a_full, cs);
                if (e8 != cudaSuccess) return cuda_fail(e8, "int8 slicing of A");
                CU(ctx->get_event(&a_full_ev));
                CU(cudaEventRecord(a_full_ev, cs));
                a_full_sliced = true;
            } else CU(cudaStreamWaitEvent(cs, a_full_ev, 0));
            TraceScope ts(ctx, cs, "gemm2", j0, nb);
            cudaError_t e8 = tmm::i8_slice_operand(reinterpret_cast<const double*>(db), b_sr, b_sk, (int)nb, (int)cl.k, i8.b_block[blk], cs);
            if (e8 == cudaSuccess)
                e8 = tmm::i8_gemm_sliced(i8.a_full, 0, (int)cl.m, i8.b_block[blk], 0, (int)nb, i8_alpha, defer_c ? 0.0 : *static_cast<const double*>(cl.beta), reinterpret_cast<double*>(dcb), ldc_dev, cs);
            if (e8 != cudaSuccess) return cuda_fail(e8, "int8 slice GEMM (phase 2)");
        } else
#endif
        if (c32.on) {  // the chunks embedded in phase 1 are the A' of the whole panel: wait for the last of them, then multiply on full k
            if (c32.last_embed) CU(cudaStreamWaitEvent(cs, c32.last_embed, 0));
            TraceScope ts(ctx, cs, "gemm2", j0, nb);
            int rc = c32_gemm(j0, nb, 0, cl.k, defer_c ? (const void*)cl.zero : cl.beta, dcb, cs);
            if (rc) return rc;
        } else {
            TraceScope ts(ctx, cs, "gemm2", j0, nb);
            int rc = launch_gemm(cl, cl.m, nb, cl.k, dA, pa, db, pb, defer_c ? (const void*)cl.zero : cl.beta, dcb, ldc_dev, cs);
            if (rc) return rc;
        }
        if (defer_c) {  // the block's share of the caller's C travels behind its B; added when both the GEMM and the copy are done
            char* dsb = (char*)dCs + (size_t)j0 * ldc_dev * es;
            {
                TraceScope ts(ctx, ctx->s_h2d, "h2dC", j0, nb);
                int rc = h2d_2d(cl, dsb, ldc_dev, cl.c + (size_t)j0 * cl.ldc * es, cl.ldc, cl.m, nb, ctx->s_h2d);
                if (rc) return rc;
            }
            CU(ctx->get_event(&ev));
            CU(cudaEventRecord(ev, ctx->s_h2d));
            CU(cudaStreamWaitEvent(cs, ev, 0));
            TraceScope ts(ctx, cs, "addC", j0, nb);
            cudaError_t e2 = tmm::device_add_scaled(cl.dtype, cl.m, nb, cl.beta, dsb, ldc_dev, dcb, ldc_dev, cs);
            if (e2 != cudaSuccess) return cuda_fail(e2, "C += beta * C_host");
        }
        if (cl.copy_c_back) {
            CU(ctx->get_event(&ev));
            CU(cudaEventRecord(ev, cs));
            CU(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
            TraceScope ts(ctx, ctx->s_d2h, "d2hC", j0, nb);
            int rc = d2h_2d(cl, cl.c + (size_t)j0 * cl.ldc * es, cl.ldc, dcb, ldc_dev, cl.m, nb, ctx->s_d2h);
            if (rc) return rc;
        }
        j0 += nb;
    }
    return TMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Streaming regime (out-of-core): C super-blocks resident, k-chunks of A and B through a slot ring.
// ------------------------------------------------------------------------------------------------
int run_streaming(Call& cl, const tmm::Plan& pl, void* dC_full, int64_t ldc_full) {
    tmm_context* ctx = cl.ctx;
    const size_t es = cl.es;
    const int SLOTS = pl.slots;
    const int64_t MB = pl.MB, NB = pl.NB, kc = pl.kc;
    const bool c_is_full = pl.c_is_full;
    cudaError_t e;
    int alloc_rc = TMM_OK;
    if ((e = ctx->buf_a.reserve(pl.bytes_a)) != cudaSuccess) alloc_rc = cuda_fail(e, "cudaMalloc(A ring)");
    else if ((e = ctx->buf_b.reserve(pl.bytes_b)) != cudaSuccess) alloc_rc = cuda_fail(e, "cudaMalloc(B ring)");
    else if (!c_is_full && (e = ctx->buf_c.reserve(pl.bytes_c)) != cudaSuccess) alloc_rc = cuda_fail(e, "cudaMalloc(C blocks)");

    if (ctx->grid.active()) {
        int rc = tmm::grid_bind(ctx, alloc_rc);  // collective even when the allocation failed here: the peers must not start without us
        if (rc) return rc;
        for (void* old : ctx->retired) cudaFree(old);  // every peer has re-mapped: nobody holds the outgrown buffers any more
        ctx->retired.clear();
    } else if (alloc_rc) return alloc_rc;
    cudaStream_t cs = ctx->s_compute[0];
    // complex<float> on the tcgen05 kernel: every ring slot gets a twin for its prepared operands (A' = 2 x the A slot, B'^T = the B slot for
    // op(B) = T / C), filled on the compute stream right before the launch that reads it - no allocation on the launch path (see run_resident)
    struct { bool on = false; float* a2 = nullptr; float* b2 = nullptr; int64_t pitch_a2 = 0, pitch_b2 = 0; size_t a2_slot = 0, b2_slot = 0; } c32;
    if (cl.dtype == TMM_C32 && tmm::c32_math_mode() == TMM_CMATH_TC && tmm::f32_math_mode() != 0 && cl.k > 0 && MB <= INT32_MAX / 2 && kc <= INT32_MAX / 2) {
        c32.pitch_a2 = cl.ta == 'N' ? round_up(2 * MB, 32) : round_up(2 * kc, 32);
        c32.a2_slot = (size_t)c32.pitch_a2 * (size_t)(2 * (cl.ta == 'N' ? kc : MB));   // floats
        c32.pitch_b2 = cl.tb == 'N' ? 2 * pl.pb_slot : round_up(NB, 32);
        c32.b2_slot = cl.tb == 'N' ? 0 : (size_t)c32.pitch_b2 * (size_t)(2 * kc);
        if (ctx->c32_a2.reserve(c32.a2_slot * SLOTS * sizeof(float)) == cudaSuccess &&
            (c32.b2_slot == 0 || ctx->c32_b2.reserve(c32.b2_slot * SLOTS * sizeof(float)) == cudaSuccess)) {
            c32.on = true;
            c32.a2 = static_cast<float*>(ctx->c32_a2.p);
            c32.b2 = static_cast<float*>(ctx->c32_b2.p);
        } else cudaGetLastError();
    }
    std::vector<cudaEvent_t> slot_free(SLOTS, nullptr);
    cudaEvent_t cbuf_free[2] = {nullptr, nullptr};
    int slot = 0, cbuf = 0, nblocks = 0;
    const int64_t nchunks = (cl.k + kc - 1) / kc;
    ctx->stats.k_chunks = (int)nchunks;
    cudaEvent_t ev;

    // The loop nest walks the PLANNED extents (m_plan, n_plan >= m, n): on a GPU grid every rank issues the same sequence of
    // panel exchanges; a rank whose own block ends earlier still contributes its upload share, and only skips the GEMM / C part.
    for (int64_t i0 = 0; i0 < cl.m_plan; i0 += MB) {
        const int64_t mi = std::max<int64_t>(0, std::min(MB, cl.m - i0));
        for (int64_t j0 = 0; j0 < cl.n_plan; j0 += NB) {
            const int64_t nj = std::max<int64_t>(0, std::min(NB, cl.n - j0));
            const bool mine = mi > 0 && nj > 0;
            char* dcb = nullptr;
            int64_t ldcb = 0;
            if (mine) {
                if (c_is_full) { dcb = (char*)dC_full + ((size_t)j0 * ldc_full + i0) * es; ldcb = ldc_full; }
                else { dcb = (char*)ctx->buf_c.p + (size_t)cbuf * pl.pc_blk * NB * es; ldcb = pl.pc_blk; }
                if (!c_is_full && cbuf_free[cbuf]) {
                    // this C buffer's previous contents must have left for the host before it is overwritten
                    CU(cudaStreamWaitEvent(ctx->s_h2d, cbuf_free[cbuf], 0));
                    CU(cudaStreamWaitEvent(cs, cbuf_free[cbuf], 0));
                }
                if (cl.beta_nonzero) {
                    TraceScope ts(ctx, ctx->s_h2d, "h2dC", i0, j0);
                    int rc = h2d_2d(cl, dcb, ldcb, cl.c + ((size_t)j0 * cl.ldc + i0) * es, cl.ldc, mi, nj, ctx->s_h2d);
                    if (rc) return rc;
                }
            }
            for (int64_t ci = 0; ci < nchunks; ++ci) {
                const int64_t p0 = ci * kc, kcc = std::min(kc, cl.k - p0);
                if (slot_free[slot]) {
                    CU(cudaStreamWaitEvent(ctx->s_h2d, slot_free[slot], 0));
                    if (ctx->grid.active()) CU(cudaStreamWaitEvent(ctx->s_comm, slot_free[slot], 0));
                }
                char* da = (char*)ctx->buf_a.p + (size_t)slot * pl.a_slot_bytes;
                char* db = (char*)ctx->buf_b.p + (size_t)slot * pl.b_slot_bytes;
                Sub sa = a_sub(cl, i0, mi, p0, kcc);
                Sub sb = b_sub(cl, p0, kcc, j0, nj);
                {
                    TraceScope ts(ctx, ctx->s_h2d, "h2dAB", p0, kcc, slot);
                    int rc = mi > 0 ? fetch_a(cl, da, pl.pa_slot, sa, slot) : TMM_OK;
                    if (rc) return rc;
                    rc = nj > 0 ? fetch_b(cl, db, pl.pb_slot, sb, slot) : TMM_OK;
                    if (rc) return rc;
                }
                {
                    int rc = panels_ready(cl, cs);
                    if (rc) return rc;
                }
                if (mine && c32.on) {
                    TraceScope ts(ctx, cs, "gemmS", i0, j0, p0);
                    float* a2 = c32.a2 + (size_t)slot * c32.a2_slot;
                    const float* b2 = reinterpret_cast<const float*>(db);  // op N: the slot read as floats
                    cudaError_t e1 = tmm::cgemm_tc_embed_a(cl.ta, (int)mi, (int)kcc, static_cast<const float*>(cl.alpha), da, pl.pa_slot, a2, c32.pitch_a2, cs);
                    if (e1 == cudaSuccess && cl.tb != 'N') {
                        float* out = c32.b2 + (size_t)slot * c32.b2_slot;
                        e1 = tmm::cgemm_tc_split_b(cl.tb, (int)nj, (int)kcc, db, pl.pb_slot, out, c32.pitch_b2, cs);
                        b2 = out;
                    }
                    if (e1 == cudaSuccess)
                        e1 = tmm::cgemm_tc_prepared(cl.ta, cl.tb, (int)mi, (int)nj, (int)kcc, a2, c32.pitch_a2, b2, c32.pitch_b2,
                                                    static_cast<const float*>(ci == 0 ? cl.beta : (const void*)cl.one), dcb, ldcb, cs);
                    if (e1 == cudaErrorInvalidValue) {  // a slot off the 16-byte grid: this launch prepares its own operands
                        cudaGetLastError();
                        int rc = launch_gemm(cl, mi, nj, kcc, da, pl.pa_slot, db, pl.pb_slot, ci == 0 ? cl.beta : (const void*)cl.one, dcb, ldcb, cs);
                        if (rc) return rc;
                    } else if (e1 != cudaSuccess) return cuda_fail(e1, "complex<float> GEMM on prepared operands (streaming)");
                } else if (mine) {
                    TraceScope ts(ctx, cs, "gemmS", i0, j0, p0);
                    int rc = launch_gemm(cl, mi, nj, kcc, da, pl.pa_slot, db, pl.pb_slot, ci == 0 ? cl.beta : (const void*)cl.one, dcb, ldcb, cs);
                    if (rc) return rc;
                }
                CU(ctx->get_event(&slot_free[slot]));
                CU(cudaEventRecord(slot_free[slot], cs));
                if (ctx->grid.active()) {  // the peers that push into this ring slot may overwrite it from here on
                    int rc = mi > 0 ? tmm::link_ack(ctx, ctx->grid.rowl, cs) : TMM_OK;
                    if (!rc && nj > 0) rc = tmm::link_ack(ctx, ctx->grid.coll, cs);
                    if (rc) return rc;
                }
                slot = (slot + 1) % SLOTS;
            }
            if (!mine) continue;
            if (cl.copy_c_back) {
                CU(ctx->get_event(&ev));
                CU(cudaEventRecord(ev, cs));
                CU(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
                TraceScope ts(ctx, ctx->s_d2h, "d2hC", i0, j0);
                int rc = d2h_2d(cl, cl.c + ((size_t)j0 * cl.ldc + i0) * es, cl.ldc, dcb, ldcb, mi, nj, ctx->s_d2h);
                if (rc) return rc;
                if (!c_is_full) {
                    CU(ctx->get_event(&cbuf_free[cbuf]));
                    CU(cudaEventRecord(cbuf_free[cbuf], ctx->s_d2h));
                }
            }
            if (!c_is_full) cbuf = (cbuf + 1) % pl.n_cbuf;
            ++nblocks;
        }
    }
    ctx->stats.c_blocks = nblocks;
    return TMM_OK;
}

int sync_all(tmm_context* ctx) {
    // the reference ends with a device-wide cudaDeviceSynchronize (tiled_mm.cpp:602-604); syncing our own
    // streams gives the same guarantee for this call without stalling unrelated work on the device
    std::vector<cudaStream_t> all = {ctx->s_h2d, ctx->s_comm, ctx->s_d2h};
    for (int i = 0; i < tmm_context::MAX_COMPUTE; ++i) all.push_back(ctx->s_compute[i]);
    for (int i = 1; i < tmm_context::MAX_P1; ++i) all.push_back(ctx->s_p1[i]);
    if (ctx->grid.active()) {
        // on a GPU grid a stream may be waiting for a peer that failed: poll with a deadline instead of blocking forever
        const char* v = getenv("TMM_DIST_TIMEOUT_S");
        const double limit_s = (v && *v) ? atof(v) : 600.0;
        const auto t0 = std::chrono::steady_clock::now();
        for (cudaStream_t s : all) {
            for (;;) {
                cudaError_t e = cudaStreamQuery(s);
                if (e == cudaSuccess) break;
                if (e != cudaErrorNotReady) return cuda_fail(e, "cudaStreamQuery");
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s)
                    return fail(TMM_ERR_CUDA, "GPU ERROR: GPU grid call did not finish within %.0f s (a peer rank failed?); the context is unusable", limit_s);
                std::this_thread::yield();
            }
        }
        return TMM_OK;
    }
    for (cudaStream_t s : all) CU(cudaStreamSynchronize(s));
    return TMM_OK;
}

int pin(tmm_context* ctx, const void* p, size_t bytes, std::vector<const void*>& pinned_now) {
    if (!p || bytes == 0) return TMM_OK;
    if (ctx->pin_cache) {
        auto it = ctx->pinned.find(p);
        if (it != ctx->pinned.end()) {
            if (it->second >= bytes) return TMM_OK;
            cudaHostUnregister(const_cast<void*>(p));
            ctx->pinned.erase(it);
        }
    }
    // memory that is already page-locked (cudaHostAlloc / gpu::malloc_pinned, or registered by the caller) needs nothing:
    // the reference would fail in cudaHostRegister here (tiled_mm.cpp:532-549); accepting it is a superset
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) {
        if (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return TMM_OK;  // nothing to page-lock
    }
    else cudaGetLastError();
    // (Registering a large range in several pieces from several threads was tried in round 2 and removed: cudaHostRegister does not get
    //  faster with threads, and a DMA copy that spans two adjacent registrations fails - profiles/r2_pin_probe.txt.)
    cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return TMM_OK; }  // already DMA-able: nothing to do, nothing to undo
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
    if (ctx->pin_cache) ctx->pinned[p] = bytes; else pinned_now.push_back(p);
    return TMM_OK;
}

}  // namespace

extern "C" {

const char* tmm_last_error(void) { return tmm::last_error_cstr(); }
const char* tmm_version(void) { return "tiled_mm_b200 0.1 (sm_100a)"; }
uint64_t tmm_total_kernel_launches(void) { return tmm::launch_count(); }
int tmm_set_f32_math(int mode) {
    if (mode != TMM_MATH_FP32 && mode != TMM_MATH_TF32 && mode != TMM_MATH_SIMT) return tmm::fail(TMM_ERR_INVALID, "tmm_set_f32_math: unknown mode %d", mode);
    tmm::set_f32_math_mode(mode);
    return TMM_OK;
}
int tmm_get_f32_math(void) { return tmm::f32_math_mode(); }
int tmm_set_c32_math(int mode) {
    if (mode != TMM_CMATH_SIMT && mode != TMM_CMATH_TC) return tmm::fail(TMM_ERR_INVALID, "tmm_set_c32_math: unknown mode %d", mode);
    tmm::set_c32_math_mode(mode);
    return TMM_OK;
}
int tmm_get_c32_math(void) { return tmm::c32_math_mode(); }

int tmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tmm_context_create(int dtype, int n_streams, int max_tile_m, int max_tile_n, int max_tile_k, tmm_context** out) {
    if (!out) return fail(TMM_ERR_INVALID, "out is null");
    *out = nullptr;
    if (dtype < TMM_F32 || dtype > TMM_C64) return fail(TMM_ERR_INVALID, "bad dtype %d", dtype);
    if (n_streams < 1 || max_tile_m < 1 || max_tile_n < 1 || max_tile_k < 1) return fail(TMM_ERR_INVALID, "streams and tile sizes must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TMM_ERR_NOGPU, "no CUDA device: tiled_mm_b200 has no CPU fallback (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    tmm_context* ctx = new tmm_context();
    ctx->dtype = dtype; ctx->n_streams = n_streams;
    ctx->max_tile_m = max_tile_m; ctx->max_tile_n = max_tile_n; ctx->max_tile_k = max_tile_k;
    ctx->tile_m = max_tile_m; ctx->tile_n = max_tile_n; ctx->tile_k = max_tile_k;
    if ((e = cudaGetDevice(&ctx->device)) != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaGetDevice"); }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaGetDeviceProperties"); }
    if (prop.major != 10) { delete ctx; return fail(TMM_ERR_NOGPU, "device %d is sm_%d%d; tiled_mm_b200 kernels are built for sm_100a only", ctx->device, prop.major, prop.minor); }
    const char* pc = getenv("TMM_PIN_CACHE");
    ctx->pin_cache = pc && pc[0] == '1';
    const char* tr = getenv("TMM_TRACE");
    ctx->trace = tr && tr[0] == '1';
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);  // numerically lower = higher priority
    cudaStream_t* all[] = {&ctx->s_h2d, &ctx->s_d2h, &ctx->s_compute[0], &ctx->s_compute[1], &ctx->s_compute[2], &ctx->s_compute[3], &ctx->s_comm};
    // phase-2 streams share ONE lower priority: among equal priorities the block scheduler drains kernels in launch
    // order, so column blocks finish in order (their D2H copies queue in that order) while still back-filling tails
    const int low = std::min(prio_least, prio_greatest + 1);
    const int prio[] = {prio_greatest, prio_greatest, prio_greatest, low, low, low, prio_greatest};
    for (int i = 0; i < 7; ++i)
        if ((e = cudaStreamCreateWithPriority(all[i], cudaStreamNonBlocking, prio[i])) != cudaSuccess) { tmm_context_destroy(ctx); return cuda_fail(e, "cudaStreamCreateWithPriority"); }
    ctx->s_p1[0] = ctx->s_compute[0];  // phase-1 stripe chains: [0] is the main high-priority compute stream
    for (int i = 1; i < tmm_context::MAX_P1; ++i)
        if ((e = cudaStreamCreateWithPriority(&ctx->s_p1[i], cudaStreamNonBlocking, prio_greatest)) != cudaSuccess) { tmm_context_destroy(ctx); return cuda_fail(e, "cudaStreamCreateWithPriority"); }
    *out = ctx;
    // TMM_DEVICES=n: every context created by the application drives the first n GPUs of the box (tmm_context_set_devices) - a caller
    // written for the reference gets the whole box without touching its code.  Not applied to the contexts the library creates itself.
    static const int auto_devices = [] { const char* v = getenv("TMM_DEVICES"); return (v && *v) ? atoi(v) : 0; }();
    if (auto_devices > 1 && !tmm::creating_internal_context()) {
        const int use = std::min(auto_devices, ndev);
        if (use > 1 && tmm_context_set_devices(ctx, use, nullptr) != TMM_OK)
            fprintf(stderr, "tiled_mm_b200: TMM_DEVICES=%d could not be honoured (%s); this context stays on device %d\n", auto_devices, tmm_last_error(), ctx->device);
    }
    return TMM_OK;
}

void tmm_context_destroy(tmm_context* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    for (auto& kv : ctx->pinned) cudaHostUnregister(const_cast<void*>(kv.first));
    for (tmm_context* ch : ctx->children) tmm_context_destroy(ch);
    ctx->children.clear();
    if (ctx->solo) { tmm_context_destroy(ctx->solo); ctx->solo = nullptr; }
    cudaStream_t all[] = {ctx->s_h2d, ctx->s_d2h, ctx->s_compute[0], ctx->s_compute[1], ctx->s_compute[2], ctx->s_compute[3], ctx->s_comm};
    for (cudaStream_t s : all) if (s) cudaStreamSynchronize(s);
    for (int i = 1; i < tmm_context::MAX_P1; ++i) if (ctx->s_p1[i]) { cudaStreamSynchronize(ctx->s_p1[i]); cudaStreamDestroy(ctx->s_p1[i]); }
    tmm::dist_release(ctx);
    for (cudaStream_t s : all) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    ctx->buf_a.release(); ctx->buf_b.release(); ctx->buf_c.release(); ctx->buf_cs.release(); ctx->full_c.release(); ctx->c32_a2.release(); ctx->c32_b2.release(); ctx->i8_q.release(); ctx->i8_e.release();
    for (void* old : ctx->retired) cudaFree(old);
    delete ctx;
}

int tmm_context_dtype(tmm_context* ctx) { return ctx ? ctx->dtype : TMM_ERR_INVALID; }
int tmm_context_get_num_streams(tmm_context* ctx) { return ctx ? ctx->n_streams : TMM_ERR_INVALID; }

int tmm_context_get_max_tile_sizes(tmm_context* ctx, int* tm, int* tn, int* tk) {
    if (!ctx) return fail(TMM_ERR_INVALID, "null context");
    if (tm) *tm = ctx->max_tile_m;
    if (tn) *tn = ctx->max_tile_n;
    if (tk) *tk = ctx->max_tile_k;
    return TMM_OK;
}

int tmm_context_set_streams_and_tiles(tmm_context* ctx, int n_streams, int tile_m, int tile_n, int tile_k) {
    if (!ctx) return fail(TMM_ERR_INVALID, "null context");
    if (n_streams < 1 || tile_m < 1 || tile_n < 1 || tile_k < 1) return fail(TMM_ERR_INVALID, "streams and tile sizes must be >= 1");
    // the reference asserts tile <= max (mm_handle.cpp:136-138); here the hint is clamped and the maxima stay what make_context fixed
    ctx->n_streams = n_streams;
    ctx->tile_m = std::min(tile_m, ctx->max_tile_m); ctx->tile_n = std::min(tile_n, ctx->max_tile_n); ctx->tile_k = std::min(tile_k, ctx->max_tile_k);
    for (tmm_context* ch : ctx->children) { ch->n_streams = n_streams; ch->tile_m = ctx->tile_m; ch->tile_n = ctx->tile_n; ch->tile_k = ctx->tile_k; }
    if (ctx->solo) { ctx->solo->n_streams = n_streams; ctx->solo->tile_m = ctx->tile_m; ctx->solo->tile_n = ctx->tile_n; ctx->solo->tile_k = ctx->tile_k; }
    return TMM_OK;
}

int tmm_context_optimal_tile_sizes(tmm_context* ctx, int m, int n, int k, int* tm, int* tn, int* tk) {
    if (!ctx) return fail(TMM_ERR_INVALID, "null context");
    if (m < 1 || n < 1 || k < 1) return fail(TMM_ERR_INVALID, "dimensions must be >= 1");
    if (tm) *tm = tmm::optimal_tile_size(m, ctx->max_tile_m);
    if (tn) *tn = tmm::optimal_tile_size(n, ctx->max_tile_n);
    if (tk) *tk = tmm::optimal_tile_size(k, ctx->max_tile_k);
    return TMM_OK;
}

namespace {
// the context that holds the device-resident C: the context itself, or - when it drives several devices - the plain context on the first one
tmm_context* device_c_owner(tmm_context* ctx) {
    if (!ctx || ctx->children.empty()) return ctx;
    return ctx->children[0]->grid.active() ? ctx->solo : ctx->children[0];
}
}  // namespace
void* tmm_context_device_c(tmm_context* ctx) { ctx = device_c_owner(ctx); return ctx ? ctx->full_c.p : nullptr; }
size_t tmm_context_device_c_size(tmm_context* ctx) { ctx = device_c_owner(ctx); return ctx ? ctx->full_c_elems : 0; }

int tmm_context_reserve_device_c(tmm_context* ctx, int64_t m, int64_t n) {
    ctx = device_c_owner(ctx);
    if (!ctx) return fail(TMM_ERR_INVALID, "null context");
    if (m < 1 || n < 1) return fail(TMM_ERR_INVALID, "set_full_sizes: dimensions must be >= 1");  // asserts in the reference, mm_handle.cpp:75-77
    DeviceGuard guard(ctx->device);
    const size_t bytes = (size_t)m * (size_t)n * tmm::dtype_size(ctx->dtype);
    if (bytes > ctx->full_c.cap) ctx->budget_cached = 0;
    cudaError_t e = ctx->full_c.reserve(bytes, 1.2);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(full C)");
    ctx->full_c_elems = (size_t)m * (size_t)n;
    return TMM_OK;
}

void* tmm_context_stream(tmm_context* ctx, int kind, int index) {
    if (!ctx) return nullptr;
    switch (kind) {
    // the reference serves stream ids [0, n_streams) (gpu_context.cpp:26-33); this context has at most four compute streams whatever the hint, so
    // larger ids wrap around instead of failing
    case TMM_STREAM_COMPUTE: return (index >= 0 && index < std::max(ctx->n_streams, ctx->n_compute())) ? (void*)ctx->s_compute[index % ctx->n_compute()] : nullptr;
    case TMM_STREAM_H2D: return index == 0 ? (void*)ctx->s_h2d : nullptr;
    case TMM_STREAM_D2H: return index == 0 ? (void*)ctx->s_d2h : nullptr;
    default: return nullptr;
    }
}

int tmm_context_last_stats(tmm_context* ctx, tmm_call_stats* out) {
    if (!ctx || !out) return fail(TMM_ERR_INVALID, "null argument");
    *out = ctx->stats;
    return TMM_OK;
}
// (both setters reach the child contexts of a multi-device parent and its single-device stand-in as well, whenever they are called - ADVICE r1)
int tmm_context_set_profiling(tmm_context* ctx, int on) {
    if (!ctx) return TMM_ERR_INVALID;
    ctx->profiling = on != 0;
    for (tmm_context* ch : ctx->children) ch->profiling = ctx->profiling;
    if (ctx->solo) ctx->solo->profiling = ctx->profiling;
    return TMM_OK;
}
int tmm_context_set_device_budget(tmm_context* ctx, size_t bytes) {
    if (!ctx) return TMM_ERR_INVALID;
    ctx->budget_override = bytes;
    for (tmm_context* ch : ctx->children) ch->budget_override = bytes;
    if (ctx->solo) ctx->solo->budget_override = bytes;
    return TMM_OK;
}

static int gemm_on_context(tmm_context* ctx, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t ld_a,
                           const void* b, int64_t ld_b, const void* beta, void* c, int64_t ld_c, int pin_host_buffers, int copy_c_back);

int tmm_gemm(tmm_context* ctx, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t ld_a,
             const void* b, int64_t ld_b, const void* beta, void* c, int64_t ld_c, int pin_host_buffers, int copy_c_back) {
    if (!ctx) return fail(TMM_ERR_INVALID, "null context");
    if (ctx->children.empty() && ctx->grid.active()) {
        // One rank of a GPU grid.  Whatever happens locally, every rank takes part in exactly one agreement round per call: a rank that
        // fails before it gets there (bad argument, failed page-locking or allocation) still plays the round and reports its failure, so
        // its peers give the call up with it instead of waiting for panel shares that will never come.
        ctx->grid_round_entered = false;
        const int rc = gemm_on_context(ctx, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, pin_host_buffers, copy_c_back);
        if (rc && !ctx->grid_round_entered) { DeviceGuard guard(ctx->device); tmm::dist_abort(ctx); }
        return rc;
    }
    return gemm_on_context(ctx, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, pin_host_buffers, copy_c_back);
}

static int gemm_on_context(tmm_context* ctx, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t ld_a,
                           const void* b, int64_t ld_b, const void* beta, void* c, int64_t ld_c, int pin_host_buffers, int copy_c_back) {
    if (!ctx->children.empty())  // single-process multi-GPU: C blocks over the child contexts, one host thread each
        return tmm::multi_gemm(ctx, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, pin_host_buffers, copy_c_back);
    const auto t_begin = std::chrono::steady_clock::now();
    Call cl;
    cl.ctx = ctx; cl.dtype = ctx->dtype; cl.es = tmm::dtype_size(ctx->dtype);
    cl.ta = (char)std::toupper((unsigned char)trans_a);  // reference tiled_mm.cpp:503-504
    cl.tb = (char)std::toupper((unsigned char)trans_b);
    if ((cl.ta != 'N' && cl.ta != 'T' && cl.ta != 'C') || (cl.tb != 'N' && cl.tb != 'T' && cl.tb != 'C'))
        return fail(TMM_ERR_INVALID, "trans must be one of N, T, C (got '%c','%c')", trans_a, trans_b);
    if (m < 0 || n < 0 || k < 0) return fail(TMM_ERR_INVALID, "negative dimension");
    if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return fail(TMM_ERR_INVALID, "dimension exceeds 2^31-1");
    if (!alpha || !beta) return fail(TMM_ERR_INVALID, "alpha/beta null");
    {   // byte counts are formed in size_t: refuse shapes whose matrices could not be addressed rather than let a product wrap
        const double limit = 4.0e18, es_d = (double)tmm::dtype_size(ctx->dtype);
        const double rows_a = (double)std::max<int64_t>(ld_a, 1), rows_b = (double)std::max<int64_t>(ld_b, 1), rows_c = (double)std::max<int64_t>(ld_c, 1);
        if (rows_a * (double)std::max(m, k) * es_d > limit || rows_b * (double)std::max(n, k) * es_d > limit || rows_c * (double)n * es_d > limit)
            return fail(TMM_ERR_INVALID, "matrix too large to address (more than 4e18 bytes)");
    }
    cl.m = m; cl.n = n; cl.k = k; cl.alpha = alpha; cl.beta = beta;
    cl.a = (const char*)a; cl.b = (const char*)b; cl.c = (char*)c;
    cl.lda = ld_a; cl.ldb = ld_b; cl.ldc = ld_c;
    cl.copy_c_back = copy_c_back != 0;
    cl.beta_nonzero = !scalar_is_zero(ctx->dtype, beta);  // std::abs(beta) > 0, tiled_mm.cpp:325
    cl.a_rows = cl.ta == 'N' ? m : k; cl.a_cols = cl.ta == 'N' ? k : m;   // tiled_mm.cpp:507-511
    cl.b_rows = cl.tb == 'N' ? k : n; cl.b_cols = cl.tb == 'N' ? n : k;
    make_one(ctx->dtype, cl.one);
    // the reference builds these errors but never throws them (tiled_mm.cpp:516-527); cuBLAS would reject them
    if (ld_a < std::max<int64_t>(1, cl.a_rows)) return fail(TMM_ERR_INVALID, "ld_a (%lld) < rows of stored A (%lld)", (long long)ld_a, (long long)cl.a_rows);
    if (ld_b < std::max<int64_t>(1, cl.b_rows)) return fail(TMM_ERR_INVALID, "ld_b (%lld) < rows of stored B (%lld)", (long long)ld_b, (long long)cl.b_rows);
    if (ld_c < std::max<int64_t>(1, m)) return fail(TMM_ERR_INVALID, "ld_c (%lld) < m (%lld)", (long long)ld_c, (long long)m);

    ctx->stats = tmm_call_stats{};
    ctx->ev_next = 0; ctx->tev_next = 0; ctx->trace_ops.clear(); ctx->gemm_events.clear();
    cudaEvent_t trace_t0 = nullptr;
    const uint64_t launches_before = tmm::launch_count();
    if ((m == 0 || n == 0) && ctx->grid.active()) return fail(TMM_ERR_INVALID, "on a GPU grid every rank must own a non-empty block of C");
    if (m == 0 || n == 0) return TMM_OK;  // BLAS quick return (the reference divides by zero here)
    const bool alpha_zero = scalar_is_zero(ctx->dtype, alpha);
    const bool need_ab = k > 0 && !alpha_zero;
    const bool c_touched = cl.copy_c_back || cl.beta_nonzero;
    if ((need_ab && (!a || !b)) || (c_touched && !c)) return fail(TMM_ERR_INVALID, "null matrix pointer");

    DeviceGuard guard(ctx->device);
    // Device-pointer operands (additive; the reference takes host pointers only).  All operands already on this context's
    // device: there is nothing to stream - one launch of the device GEMM on them (any ld / alignment, csrc/tmm_common.cu), the result in
    // c itself or, with copy_c_back = false, in the context's device C.  A mix of host and device operands goes through the scheduler
    // with copies that infer their direction.
    {
        auto where = [&](const void* p, int* dev) {
            cudaPointerAttributes attr;
            *dev = -1;
            if (!p) return 0;
            if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return 0; }
            if (attr.type == cudaMemoryTypeDevice) { *dev = attr.device; return 1; }
            return 0;
        };
        int da = -1, db = -1, dc = -1;
        const int na = where(a, &da), nb = where(b, &db), nc = where(c, &dc);
        cl.device_operands = (na + nb + nc) > 0;
        if (cl.device_operands && ctx->grid.active())
            return fail(TMM_ERR_INVALID, "device-pointer operands are not supported on a GPU grid (the panel exchange uploads from host memory)");
        const bool all_here = need_ab && na && nb && (nc || !c_touched) && da == ctx->device && db == ctx->device && (!nc || dc == ctx->device);
        if (all_here) {
            cudaStream_t st = ctx->s_compute[0];
            void* dC = c;
            int64_t ldd = ld_c;
            cudaError_t e = cudaSuccess;
            int rc0 = TMM_OK;
            if (!cl.copy_c_back) {
                if ((size_t)m * n * cl.es > ctx->full_c.cap) ctx->budget_cached = 0;
                if ((e = ctx->full_c.reserve((size_t)m * n * cl.es, 1.2)) != cudaSuccess) rc0 = cuda_fail(e, "cudaMalloc(full C)");
                ctx->full_c_elems = (size_t)m * n;
                dC = ctx->full_c.p; ldd = m;
                if (!rc0 && cl.beta_nonzero &&
                    (e = cudaMemcpy2DAsync(dC, (size_t)ldd * cl.es, c, (size_t)ld_c * cl.es, (size_t)m * cl.es, (size_t)n, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
                    rc0 = cuda_fail(e, "cudaMemcpy2DAsync(C)");
            }
            if (!rc0 && (e = tmm::device_gemm(cl.dtype, cl.ta, cl.tb, m, n, k, alpha, a, ld_a, b, ld_b, beta, dC, ldd, st)) != cudaSuccess) rc0 = cuda_fail(e, "device_gemm");
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess && !rc0) rc0 = cuda_fail(e, "cudaStreamSynchronize");
            ctx->stats.kernel_launches = tmm::launch_count() - launches_before;
            ctx->stats.c_blocks = 1; ctx->stats.k_chunks = 1;
            ctx->stats.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
            return rc0;
        }
    }
    if (ctx->trace && ctx->get_timing_event(&trace_t0) == cudaSuccess) cudaEventRecord(trace_t0, ctx->s_h2d);
    std::vector<const void*> pinned_now;
    int rc = TMM_OK;
    if (pin_host_buffers) {  // reference tiled_mm.cpp:529-554
        if (need_ab) {
            rc = pin(ctx, a, (size_t)ld_a * cl.a_cols * cl.es, pinned_now);
            if (!rc) rc = pin(ctx, b, (size_t)ld_b * cl.b_cols * cl.es, pinned_now);
        }
        if (!rc && c_touched) rc = pin(ctx, c, (size_t)ld_c * n * cl.es, pinned_now);
    }

    // Planning and enqueue.  A device allocation can fail although the plan fitted the budget (the budget is a cached reading of free
    // HBM, and somebody else may have allocated since): before anything has been enqueued that is recoverable - give back this
    // context's panel / ring / C storage, re-read the free memory and plan again (typically into the streaming regime).  Once.
    for (int attempt = 0; !rc && attempt < 2; ++attempt) {
        const int64_t align = 128 / (int64_t)cl.es;
        void* dC = nullptr;
        int64_t ldc_dev = 0;
        cudaError_t e = cudaSuccess;
        size_t c_bytes = 0;
        if (!cl.copy_c_back) {
            // device-resident C, always compact ld = m (README.md:102-103, tests/test-multiply.cpp:339)
            if ((size_t)m * n * cl.es > ctx->full_c.cap) ctx->budget_cached = 0;
            e = ctx->full_c.reserve((size_t)m * n * cl.es, 1.2);  // device_vector keeps 1.2x slack (device_vector.hpp:92-107)
            ctx->full_c_elems = (size_t)m * n;
            dC = ctx->full_c.p; ldc_dev = m;
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(full C)");
        } else {
            ldc_dev = round_up(m, align);
            c_bytes = (size_t)ldc_dev * n * cl.es;
        }
        const size_t budget = device_budget(ctx);  // after full C is in place
        // on a GPU grid all ranks plan for the same (largest) block and the smallest budget, so the exchanges line up; the round also
        // checks that k, the ops and the modes (beta == 0, copy_c_back, "nothing to multiply") agree, and spreads a local failure
        int64_t m_plan = m, n_plan = n;
        size_t plan_budget = budget;
        int rc_before_agree = TMM_OK;
        if (ctx->grid.active()) {
            const int flags = (cl.ta << 16) | (cl.tb << 8) | (need_ab ? 4 : 0) | (cl.beta_nonzero ? 2 : 0) | (cl.copy_c_back ? 1 : 0);
            const int rc_agree = tmm::dist_agree(ctx, m, n, k, flags, budget, &m_plan, &n_plan, &plan_budget, rc);
            if (!rc) rc = rc_agree;
            rc_before_agree = rc;
            if (!rc) tmm::dist_set_shares(ctx, m_plan, n_plan, k, cl.es, cl.beta_nonzero, cl.copy_c_back);  // who uploads how much of a shared panel
        }
        if (!rc) {
            if (!need_ab) {
                // C = beta * C
                if (cl.copy_c_back) {
                    if ((e = ctx->buf_c.reserve(c_bytes)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(C)");
                    dC = ctx->buf_c.p;
                }
                if (!rc && cl.beta_nonzero) rc = h2d_2d(cl, dC, ldc_dev, cl.c, cl.ldc, m, n, ctx->s_h2d);
                if (!rc) {
                    cudaEvent_t ev = nullptr;
                    if ((e = ctx->get_event(&ev)) != cudaSuccess || (e = cudaEventRecord(ev, ctx->s_h2d)) != cudaSuccess ||
                        (e = cudaStreamWaitEvent(ctx->s_compute[0], ev, 0)) != cudaSuccess ||
                        (e = tmm::device_scale(cl.dtype, m, n, beta, dC, ldc_dev, ctx->s_compute[0])) != cudaSuccess)
                        rc = cuda_fail(e, "scale C");
                    if (!rc && cl.copy_c_back) {
                        if ((e = cudaEventRecord(ev, ctx->s_compute[0])) != cudaSuccess || (e = cudaStreamWaitEvent(ctx->s_d2h, ev, 0)) != cudaSuccess) rc = cuda_fail(e, "event");
                        if (!rc) rc = d2h_2d(cl, cl.c, cl.ldc, dC, ldc_dev, m, n, ctx->s_d2h);
                    }
                }
            } else {
                const tmm::Grid& gr = ctx->grid;
                bool need_stage = false;  // only links that could not map peer memory stage their shares through NCCL
                if (gr.active()) {
                    need_stage = (gr.rowl.active() && !gr.rowl.direct) || (gr.coll.active() && !gr.coll.direct);
                    if (!rc && need_stage) {
                        // staging rings for the all-gathers: shares of at most one k-chunk of A / one column block of B
                        const int64_t kcap = std::min<int64_t>(k, 2048), ncap = std::min<int64_t>(n_plan, 8192);
                        size_t share = 0;
                        if (gr.pc > 1) share = std::max(share, (size_t)(cl.ta == 'N' ? m_plan * ((kcap + gr.pc - 1) / gr.pc) : kcap * ((m_plan + gr.pc - 1) / gr.pc)) * cl.es);
                        if (gr.pr > 1) share = std::max(share, (size_t)(cl.tb == 'N' ? k * ((ncap + gr.pr - 1) / gr.pr) : ncap * ((k + gr.pr - 1) / gr.pr)) * cl.es);
                        const size_t stage = tmm::dist_stage_bytes(share, std::max(gr.pr, gr.pc));
                        const size_t held = ctx->stage_send.cap + ctx->stage_recv.cap;  // already counted as used by cudaMemGetInfo
                        const size_t extra = stage > held ? stage - held : 0;
                        if (!ctx->budget_override) plan_budget = plan_budget > extra ? plan_budget - extra : 0;  // an explicit budget bounds panel storage only
                    }
                }
                cl.m_plan = m_plan; cl.n_plan = n_plan;
                tmm::PlanInput pin_;
                pin_.dtype = cl.dtype; pin_.ta = cl.ta; pin_.tb = cl.tb; pin_.m = m_plan; pin_.n = n_plan; pin_.k = k;
                pin_.beta_nonzero = cl.beta_nonzero; pin_.copy_c_back = cl.copy_c_back; pin_.budget = plan_budget;
                pin_.n_streams = ctx->n_streams; pin_.tile_m = ctx->tile_m; pin_.tile_n = ctx->tile_n; pin_.tile_k = ctx->tile_k;
                pin_.sm_count = tmm::sm_count();
                pin_.parts_a = gr.pc; pin_.parts_b = gr.pr;
                if (gr.active() && ctx->link_h2d_gbs > 0 && ctx->link_d2h_gbs > 0) { pin_.h2d_bw = ctx->link_h2d_gbs * 1e9; pin_.d2h_bw = ctx->link_d2h_gbs * 1e9; }
                if (cl.dtype == TMM_C32 && tmm::c32_math_mode() == TMM_CMATH_TC) pin_.flops = 140e12;  // complex<float> on the tcgen05 kernel (8mnk real flops)
#ifndef TMM_EMULATED
                if (cl.dtype == TMM_F64 && tmm::f64_i8_slices() > 0) {  // opt-in FP64 emulation: a faster GEMM wants a wider first block (the call becomes upload-bound)
                    const char* fv = getenv("TMM_PLAN_F64_FLOPS");
                    pin_.flops = (fv && *fv) ? atof(fv) : (tmm::f64_i8_slices() <= 7 ? 55e12 : 42e12);
                }
#endif
                tmm::Plan pl;
                if (!rc) pl = tmm::make_plan(pin_);
                ctx->stats.regime = pl.regime;
                if (!rc && need_stage) {
                    // exact staging need of this plan
                    size_t share = 0;
                    auto upd = [&](int64_t rows, int64_t cols, int parts) { if (parts > 1) share = std::max(share, (size_t)rows * (size_t)((cols + parts - 1) / parts) * cl.es); };
                    if (pl.regime == tmm::REGIME_RESIDENT) {
                        for (int64_t kc : pl.chunks) {
                            Sub sa = a_sub(cl, 0, m, 0, kc), sb = b_sub(cl, 0, kc, 0, std::min(pl.n1, n));
                            upd(sa.rows, sa.cols, gr.pc); upd(sb.rows, sb.cols, gr.pr);
                        }
                        for (int64_t nb : pl.blocks) { Sub sb = b_sub(cl, 0, k, 0, std::min(nb, n)); upd(sb.rows, sb.cols, gr.pr); }
                    } else if (pl.error.empty()) {
                        Sub sa = a_sub(cl, 0, std::min(pl.MB, m), 0, std::min(pl.kc, k)), sb = b_sub(cl, 0, std::min(pl.kc, k), 0, std::min(pl.NB, n));
                        upd(sa.rows, sa.cols, gr.pc); upd(sb.rows, sb.cols, gr.pr);
                    }
                    rc = tmm::dist_reserve_stage(ctx, share, std::max(gr.pr, gr.pc));
                }
                const bool agreed = gr.active() && !rc_before_agree && pl.error.empty();  // every rank of the grid got this far with the same plan
                if (rc) { if (agreed) tmm::grid_bind(ctx, rc); }  // (staging ring) tell the peers we are out: they are about to bind
                else if (!pl.error.empty()) rc = fail(TMM_ERR_NOMEM, "%s (budget %zu B)", pl.error.c_str(), plan_budget);
                else if (pl.regime == tmm::REGIME_RESIDENT) {
                    if (cl.copy_c_back) {
                        if ((e = ctx->buf_c.reserve(pl.bytes_c)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(C)");
                        dC = ctx->buf_c.p;
                    }
                    void* dCs = nullptr;  // staging copy of the caller's C (beta != 0): same pitch as dC
                    if (!rc && pl.bytes_c_stage) {
                        const size_t need = (size_t)(cl.copy_c_back ? pl.pitch_c : ldc_dev) * n * cl.es;
                        if ((e = ctx->buf_cs.reserve(need)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(C staging)");
                        dCs = ctx->buf_cs.p;
                    }
                    if (!rc) rc = run_resident(cl, pl, dC, cl.copy_c_back ? pl.pitch_c : ldc_dev, dCs);
                    else if (agreed) tmm::grid_bind(ctx, rc);
                } else {
                    rc = run_streaming(cl, pl, cl.copy_c_back ? nullptr : dC, ldc_dev);
                }
            }
        }
        const bool nothing_enqueued = ctx->stats.h2d_copies == 0 && ctx->stats.d2h_copies == 0 && tmm::launch_count() == launches_before;
        if (rc == TMM_ERR_NOMEM && attempt == 0 && nothing_enqueued && !ctx->grid.active() && !ctx->budget_override) {
            fprintf(stderr, "tiled_mm_b200: device allocation failed (%s); releasing staging storage and re-planning with the current free memory\n", tmm_last_error());
            ctx->buf_a.release(); ctx->buf_b.release(); ctx->buf_c.release(); ctx->buf_cs.release();
            ctx->budget_cached = 0;
            rc = TMM_OK;
            continue;
        }
        break;
    }
    const double t_enqueued = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    TMM_DBG("dev %d enqueued rc %d regime %d, syncing", ctx->device, rc, ctx->stats.regime);
    int rc_sync = sync_all(ctx);
    TMM_DBG("dev %d synced rc %d", ctx->device, rc_sync);
    if (!rc) rc = rc_sync;
    for (const void* p : pinned_now) cudaHostUnregister(const_cast<void*>(p));  // tiled_mm.cpp:606-618
    if (!rc) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "kernel execution");
    }
    double kernel_ms_total = 0;
    if (!rc && ctx->profiling) {
        for (auto& pr : ctx->gemm_events) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) kernel_ms_total += ms;
        }
    }
    if (ctx->trace && trace_t0) {
        fprintf(stderr, "[tmm trace] host: enqueue done at %.3f ms, all streams idle at %.3f ms after entry\n", t_enqueued,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        fprintf(stderr, "[tmm trace] %-28s %10s %10s %9s\n", "op", "start_ms", "end_ms", "dur_ms");
        // TMM_TRACE_FILE=<path>: the same timeline as a Chrome / Perfetto trace (one row per stream, one process per device), appended
        const char* trace_path = getenv("TMM_TRACE_FILE");
        FILE* tf = (trace_path && *trace_path) ? fopen(trace_path, "a") : nullptr;
        auto stream_name = [&](cudaStream_t st) -> std::string {
            if (st == ctx->s_h2d) return "H2D";
            if (st == ctx->s_d2h) return "D2H";
            if (st == ctx->s_comm) return "NVLink";
            for (int i = 0; i < tmm_context::MAX_P1; ++i) if (st == ctx->s_p1[i]) return "phase-1 stripe chain " + std::to_string(i);
            for (int i = 0; i < tmm_context::MAX_COMPUTE; ++i) if (st == ctx->s_compute[i]) return "column blocks " + std::to_string(i);
            return "stream";
        };
        for (auto& op : ctx->trace_ops) {
            float t_start = 0, t_end = 0;
            cudaEventElapsedTime(&t_start, trace_t0, op.e0);
            cudaEventElapsedTime(&t_end, trace_t0, op.e1);
            fprintf(stderr, "[tmm trace] %-28s %10.3f %10.3f %9.3f\n", op.name.c_str(), t_start, t_end, t_end - t_start);
            if (tf) fprintf(tf, "{\"name\":\"%s\",\"ph\":\"X\",\"ts\":%.1f,\"dur\":%.1f,\"pid\":%d,\"tid\":\"%s\"},\n", op.name.c_str(), t_start * 1e3, (t_end - t_start) * 1e3,
                            ctx->device, stream_name(op.stream).c_str());
        }
        if (tf) fclose(tf);
    }
    if (rc == TMM_ERR_NOMEM) ctx->budget_cached = 0;  // free memory changed under us: re-query next time
    ctx->stats.kernel_ms = kernel_ms_total;
    ctx->stats.kernel_launches = tmm::launch_count() - launches_before;
    ctx->stats.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return rc;
}

int tmm_optimal_tile_size(int dim, int max_tile) {
    if (dim < 1 || max_tile < 1) return TMM_ERR_INVALID;
    return tmm::optimal_tile_size(dim, max_tile);
}

int tmm_plan_describe(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, int beta_nonzero, int copy_c_back, size_t budget_bytes,
                      int n_streams, int tile_m, int tile_n, int tile_k, int sm_count, char* out, size_t out_size) {
    if (dtype < TMM_F32 || dtype > TMM_C64 || m < 1 || n < 1 || k < 1 || !out || out_size == 0) return fail(TMM_ERR_INVALID, "plan_describe: bad argument");
    tmm::PlanInput in;
    in.dtype = dtype; in.ta = (char)std::toupper((unsigned char)trans_a); in.tb = (char)std::toupper((unsigned char)trans_b);
    in.m = m; in.n = n; in.k = k; in.beta_nonzero = beta_nonzero != 0; in.copy_c_back = copy_c_back != 0; in.budget = budget_bytes;
    in.n_streams = n_streams; in.tile_m = tile_m; in.tile_n = tile_n; in.tile_k = tile_k; in.sm_count = sm_count > 0 ? sm_count : 148;
    const tmm::Plan pl = tmm::make_plan(in);
    const std::string js = tmm::plan_to_json(in, pl);
    if (js.size() + 1 > out_size) return fail(TMM_ERR_INVALID, "plan_describe: buffer too small (%zu needed)", js.size() + 1);
    memcpy(out, js.c_str(), js.size() + 1);
    return TMM_OK;
}

}  // extern "C"

namespace {
// NUMA node of a CUDA device from sysfs (-1: unknown / single node)
int device_numa_node(int dev) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, (int)sizeof bdf, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* q = bdf; *q; ++q) *q = (char)std::tolower((unsigned char)*q);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
// While alive, this thread's page allocations prefer `node` (MPOL_PREFERRED: soft, falls back when the node is full); the previous
// policy is restored afterwards.  Raw syscalls: no libnuma dependency; any failure (seccomp, no NUMA) simply leaves the policy alone.
// nodes listed in /sys/devices/system/node/online ("0-1", "0,2-3"), as a bit mask; empty when there is one node or none
std::vector<unsigned long> online_numa_nodes(unsigned long maxnode, int* count) {
    std::vector<unsigned long> mask(maxnode / 64, 0ul);
    *count = 0;
    FILE* f = fopen("/sys/devices/system/node/online", "r");
    if (!f) return mask;
    char buf[256] = {0};
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    char* save = nullptr;
    for (char* tok = strtok_r(buf, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
        int lo = 0, hi = 0;
        const int got = sscanf(tok, "%d-%d", &lo, &hi);
        if (got < 1) continue;
        if (got == 1) hi = lo;
        for (int nd = lo; nd <= hi && nd < (int)maxnode; ++nd) { mask[nd / 64] |= 1ul << (nd % 64); ++*count; }
    }
    return mask;
}
// While alive, this thread's page allocations prefer `node` (MPOL_PREFERRED: soft, falls back when the node is full) or, with
// node == INTERLEAVE, are spread page by page over all nodes (for buffers that every GPU of a two-socket box reads); the previous
// policy is restored afterwards.  Raw syscalls: no libnuma dependency; any failure (seccomp, no NUMA) simply leaves the policy alone.
struct NumaPreference {
    static constexpr unsigned long MAXNODE = 1024;
    static constexpr int INTERLEAVE = -2;
    int old_mode = 0;
    unsigned long old_mask[MAXNODE / 64] = {0};
    bool active = false;
    explicit NumaPreference(int node) {
#if defined(__linux__) && defined(SYS_set_mempolicy) && defined(SYS_get_mempolicy)
        if (node == -1 || node >= (int)MAXNODE) return;
        if (syscall(SYS_get_mempolicy, &old_mode, old_mask, MAXNODE + 1, nullptr, 0ul) != 0) return;
        if (node == INTERLEAVE) {
            int count = 0;
            std::vector<unsigned long> all = online_numa_nodes(MAXNODE, &count);
            if (count < 2) return;
            active = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, all.data(), MAXNODE + 1) == 0;
            return;
        }
        unsigned long mask[MAXNODE / 64] = {0};
        mask[node / 64] |= 1ul << (node % 64);
        active = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, MAXNODE + 1) == 0;
#else
        (void)node;
#endif
    }
    ~NumaPreference() {
#if defined(__linux__) && defined(SYS_set_mempolicy)
        if (active) syscall(SYS_set_mempolicy, old_mode, old_mode == 0 ? nullptr : old_mask, old_mode == 0 ? 0ul : MAXNODE + 1);
#endif
    }
};
}  // namespace

extern "C" {

int tmm_malloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(TMM_ERR_INVALID, "out is null");
    *out = nullptr;
    // Pinned pages on the NUMA node of the calling thread's current device: a panel that has to cross the socket interconnect on
    // its way to the PCIe root port shares that link with every other GPU of the far socket ("NUMA-local pinned pages").
    // TMM_PINNED_NUMA=0 leaves placement to the caller (numactl, first touch).
    // TMM_PINNED_NUMA=interleave spreads the pages over all nodes instead: for buffers that GPUs on both sockets read (one process driving
    // the whole box through tmm_context_set_devices).
    const char* mode = getenv("TMM_PINNED_NUMA");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = -1; }
    int node = -1;
    if (mode && (mode[0] == 'i' || mode[0] == 'I')) node = NumaPreference::INTERLEAVE;
    else if (!(mode && mode[0] == '0') && dev >= 0) node = device_numa_node(dev);
    NumaPreference pref(node);
    CU(cudaHostAlloc(out, bytes ? bytes : 1, 0));  // flags 0, reference util.hpp:67
    return TMM_OK;
}

namespace {
std::mutex g_large_mu;
std::map<void*, size_t> g_large;  // allocations of tmm_malloc_pinned_large: base -> mapped length
}  // namespace

// Hundreds of GB of pinned host memory (the out-of-core configs): cudaHostAlloc page-locks at ~2 GB/s on this pool whatever the thread
// count (240 GB: two minutes), while anonymous memory on 2 MiB pages, first touched by all cores and registered in ONE cudaHostRegister,
// reaches 26 GB/s and copies at the same 55 GB/s (profiles/r2_pin_probe.txt).  Additive: gpu::malloc_pinned stays cudaHostAlloc, whose
// result callers may release with cudaFreeHost; memory from here is released with tmm_free_pinned only.
int tmm_malloc_pinned_large(size_t bytes, void** out) {
    if (!out) return fail(TMM_ERR_INVALID, "out is null");
    *out = nullptr;
#if defined(__linux__) && !defined(TMM_EMULATED)
    const size_t huge = (size_t)2 << 20;
    const size_t len = (std::max<size_t>(bytes, 1) + huge - 1) / huge * huge;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return fail(TMM_ERR_NOMEM, "mmap of %zu bytes of host memory failed", len);
    madvise(p, len, MADV_HUGEPAGE);  // best effort: with 4 KiB pages the registration is ~3x slower, not wrong
    {
        const unsigned n_threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t)
            pool.emplace_back([=] {
                char* q = static_cast<char*>(p);
                const size_t lo = len / huge * t / n_threads * huge, hi = len / huge * (t + 1) / n_threads * huge;
                for (size_t off = lo; off < hi; off += 4096) q[off] = 0;  // first touch: the kernel zero-fills the page
            });
        for (auto& th : pool) th.join();
    }
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) { munmap(p, len); return cuda_fail(e, "cudaHostRegister(large pinned allocation)"); }
    {
        std::lock_guard<std::mutex> lk(g_large_mu);
        g_large[p] = len;
    }
    *out = p;
    return TMM_OK;
#else
    return tmm_malloc_pinned(bytes, out);
#endif
}

int tmm_free_pinned(void* p) {
    if (!p) return TMM_OK;
#if defined(__linux__) && !defined(TMM_EMULATED)
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_large_mu);
        auto it = g_large.find(p);
        if (it != g_large.end()) { len = it->second; g_large.erase(it); }
    }
    if (len) {
        cudaError_t e = cudaHostUnregister(p);
        munmap(p, len);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
        return TMM_OK;
    }
#endif
    CU(cudaFreeHost(p));
    return TMM_OK;
}
int tmm_malloc_device(size_t bytes, void** out) {
    if (!out) return fail(TMM_ERR_INVALID, "out is null");
    *out = nullptr;
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return TMM_OK;
}
int tmm_free_device(void* p) { if (p) CU(cudaFree(p)); return TMM_OK; }
int tmm_copy_to_device(const void* from, void* to, size_t bytes) { CU(cudaMemcpy(to, from, bytes, cudaMemcpyHostToDevice)); return TMM_OK; }
int tmm_copy_to_host(const void* from, void* to, size_t bytes) { CU(cudaMemcpy(to, from, bytes, cudaMemcpyDeviceToHost)); return TMM_OK; }

int tmm_device_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t ld_a,
                    const void* b, int64_t ld_b, const void* beta, void* c, int64_t ld_c, void* stream) {
    if (dtype < TMM_F32 || dtype > TMM_C64) return fail(TMM_ERR_INVALID, "bad dtype %d", dtype);
    cudaError_t e = tmm::device_gemm(dtype, trans_a, trans_b, m, n, k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) { cudaGetLastError(); return fail(TMM_ERR_INVALID, "device_gemm: invalid argument (trans, sizes, or A/B not 16-byte aligned)"); }
    if (e != cudaSuccess) return cuda_fail(e, "device_gemm");
    return TMM_OK;
}

int tmm_device_gemm_bf16(char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, float alpha, const void* a, int64_t ld_a, const void* b, int64_t ld_b,
                         float beta, float* c, int64_t ld_c, void* stream) {
    const char ta = (char)std::toupper((unsigned char)trans_a), tb = (char)std::toupper((unsigned char)trans_b);
    if ((ta != 'N' && ta != 'T' && ta != 'C') || (tb != 'N' && tb != 'T' && tb != 'C')) return fail(TMM_ERR_INVALID, "trans must be one of N, T, C");
    if (m < 0 || n < 0 || k < 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return fail(TMM_ERR_INVALID, "bad dimension");
    if (ld_a < std::max<int64_t>(1, ta == 'N' ? m : k) || ld_b < std::max<int64_t>(1, tb == 'N' ? k : n) || ld_c < std::max<int64_t>(1, m))
        return fail(TMM_ERR_INVALID, "leading dimension too small");
    if (m == 0 || n == 0) return TMM_OK;
    if ((k > 0 && alpha != 0.f && (!a || !b)) || !c) return fail(TMM_ERR_INVALID, "null matrix pointer");
    cudaError_t e = (k == 0 || alpha == 0.f) ? tmm::device_scale(TMM_F32, m, n, &beta, c, ld_c, (cudaStream_t)stream)
                                             : tmm::bgemm_tc_launch(ta, tb, (int)m, (int)n, (int)k, alpha, a, ld_a, b, ld_b, beta, c, ld_c, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "device_gemm_bf16");
    return TMM_OK;
}

}  // extern "C"
