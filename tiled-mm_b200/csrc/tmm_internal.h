// Internal definitions shared by the scheduler (tmm_context.cu) and the multi-GPU layer (tmm_dist.cu).
#pragma once
#include "../../include/tiled_mm_b200.h"
#include "tmm_blas.h"
#include "tmm_nccl.h"
#include "tmm_plan.h"

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <string>
#include <array>
#include <vector>

namespace tmm {

// error plumbing: every failure sets the thread-local message returned by tmm_last_error()
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);  // also prints the CUDA error string to stderr (reference util.hpp:13-19)
int nccl_fail(int rc, const char* what);

#define TMM_CU(x)                                                \
    do {                                                         \
        cudaError_t e__ = (x);                                   \
        if (e__ != cudaSuccess) return ::tmm::cuda_fail(e__, #x); \
    } while (0)

// developer tracing of the multi-GPU path: TMM_DEBUG=1 prints one line per step to stderr
bool debug_on();
double debug_ms();
#define TMM_DBG(...)                                  \
    do {                                              \
        if (::tmm::debug_on()) { fprintf(stderr, "[tmm dbg %9.3f] ", ::tmm::debug_ms()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } \
    } while (0)

inline int64_t round_up64(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;  // bytes
    // When peers of a GPU grid have this buffer mapped (CUDA IPC), it must not be freed before they have closed their mapping
    // ("cudaFree on an exported region before cudaIpcCloseMemHandle in the importing context is undefined behaviour"): a buffer that is
    // outgrown is parked here instead and freed after the next link_bind, whose closing collective comes after every peer has re-mapped.
    std::vector<void*>* retire = nullptr;
    cudaError_t reserve(size_t bytes, double slack = 1.0) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { if (retire) retire->push_back(p); else cudaFree(p); p = nullptr; cap = 0; }
        size_t want = (size_t)std::ceil((double)bytes * slack);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && slack > 1.0) { want = bytes; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); switched = true; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// Position of a context in a p_r x p_c grid of GPUs over the C blocks.  The grid row shares the A
// row-panel, the grid column shares the B column-panel; each member uploads a distinct 1/p share of a shared
// panel over its own PCIe link and the shares are all-gathered over NVLink (NCCL) - no k split, no reduction.
// One communicator of the grid: my grid row (shares the A row-panel) or my grid column (shares the B column-panel).
// Control plane: NCCL (agreement, handle exchange).  Data plane, when the peers' buffers can be mapped (same process, or
// CUDA IPC across processes): every rank DMA-pushes its upload share straight into the peers' panels with the copy engines
// (no SMs, no staging) and raises a per-peer arrival counter that consumers wait on with stream memory operations.
// If mapping fails the shares are all-gathered by NCCL through a staging ring instead (needs SMs).
struct Link {
    nccl::Comm comm = nullptr;
    int parts = 1, me = 0;
    bool direct = false;
    uint32_t* flags = nullptr;             // my flag block (device): arrive[parts] | ack[parts] | word[2]
    std::vector<uint32_t*> peer_flags;     // the peers' flag blocks, mapped
    std::vector<bool> peer_flags_ipc;      // peer_flags[g] came from cudaIpcOpenMemHandle (closed at teardown)
    std::vector<char*> peer_base;          // the peers' panel buffers for the current call, mapped ([me] = mine)
    std::vector<std::string> peer_key;     // what peer_base[g] was mapped from (pid, pointer, IPC handle)
    std::vector<bool> peer_ipc;            // peer_base[g] came from cudaIpcOpenMemHandle (close when replaced)
    char* local_base = nullptr;
    uint32_t sent = 0;                     // exchanges issued on this link since attach (the same number on all its ranks)
    uint32_t ring_sent = 0, acked = 0;     // streaming ring: exchanges into ring slots / exchanges consumed here
    uint32_t slot_last[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // ring exchange number that last filled each ring slot
    cudaEvent_t slot_pushed[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // per call: my push out of each ring slot has finished
    // Control board: a few KB of POSIX shared memory that all ranks of the link (threads or processes of one box) have mapped.  The per-call
    // agreement and handle exchange are rounds of "publish my 128 bytes, read everybody's" on it - microseconds of host time, no kernel
    // launch, no stream - instead of blocking NCCL collectives.  nullptr: the board could not be set up and the rounds go through NCCL.
    void* board = nullptr;
    size_t board_bytes = 0;
    uint64_t board_round = 0;              // rounds played on the board (the same number on all ranks of the link)
    uint64_t board_key = 0;                // hash of the link's unique id (name of the segment)
    bool broken = false;                   // a peer missed a round's deadline: the link is out of step for good
    // Upload shares by link speed (DMA-push plane): the ranks of a link do not sit behind equally fast host links (8-GPU node of this pool: 8 vs
    // 11 GB/s each way with all links busy), so the rank behind the slower link uploads the smaller part of a shared panel.
    std::vector<int32_t> rate_up, rate_down;  // MB/s of every rank of the link, measured at attach with the whole grid copying (empty: unknown)
    std::vector<int64_t> cut;                 // per call: share boundaries in millionths, cut[0] = 0 .. cut[parts] = 1000000 (empty: equal shares)
    bool active() const { return parts > 1; }
};

// Position of a context in a p_r x p_c grid of GPUs over the C blocks.  No k split, no reduction.
struct Grid {
    int pr = 1, pc = 1, row = 0, col = 0;
    Link rowl;  // the p_c ranks of my grid row    (my rank = col): exchanges A
    Link coll;  // the p_r ranks of my grid column (my rank = row): exchanges B
    bool active() const { return pr * pc > 1; }
};

constexpr int STAGE_SLOTS = 3;

}  // namespace tmm

struct tmm_context {
    int dtype = TMM_F64;
    int n_streams = 2;
    int max_tile_m = 5000, max_tile_n = 5000, max_tile_k = 5000;  // fixed at creation (reference mm_handle.cpp:10-16)
    int tile_m = 5000, tile_n = 5000, tile_k = 5000;              // current staging hints, <= the maxima (mm_handle.cpp:57-66,135-145)
    int device = 0;
    static constexpr int MAX_COMPUTE = 4, MAX_P1 = 4;
    cudaStream_t s_p1[MAX_P1] = {nullptr, nullptr, nullptr, nullptr};  // phase-1 stripe chains ([0] aliases s_compute[0])
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comm = nullptr, s_compute[MAX_COMPUTE] = {nullptr, nullptr, nullptr, nullptr};
    tmm::DevBuf buf_a, buf_b, buf_c;  // panel / ring / staged-C storage (grow-only, reused across calls)
    tmm::DevBuf buf_cs;               // beta != 0, resident regime: staging copy of the caller's C (same pitch as the device C of the call)
    tmm::DevBuf c32_a2, c32_b2;       // complex<float> on the tcgen05 kernel: the real embedding A' of the resident A panel (2 x |A|) and, for op(B) = T / C, B'^T (|B|)
    tmm::DevBuf i8_q, i8_e;           // opt-in FP64 emulation (TMM_F64_MATH=i8): int8 slices and row scales of the panels, carved per call
    tmm::DevBuf full_c;               // API-visible device C (copy_c_back = false), column-major ld = m
    size_t full_c_elems = 0;
    std::vector<cudaEvent_t> events;
    size_t ev_next = 0;
    std::vector<cudaEvent_t> timing_events;
    size_t tev_next = 0;
    size_t budget_override = 0;
    size_t budget_cached = 0;
    bool profiling = false;
    bool pin_cache = false;
    bool trace = false;
    struct TraceOp { std::string name; cudaEvent_t e0, e1; cudaStream_t stream; };
    std::vector<TraceOp> trace_ops;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
    std::map<const void*, size_t> pinned;
    tmm_call_stats stats{};

    // ---- multi-GPU (tmm_dist.cu) ----
    tmm::Grid grid;                         // this context's place in a GPU grid (per-process or child of a parent)
    tmm::DevBuf stage_send, stage_recv, dist_scratch;  // all-gather staging rings (STAGE_SLOTS slots each)
    size_t stage_slot_bytes = 0;            // bytes of one share (send slot); a recv slot holds `parts` of them
    int stage_next = 0;
    cudaEvent_t stage_send_free[tmm::STAGE_SLOTS] = {nullptr, nullptr, nullptr};
    std::vector<void*> retired;             // outgrown panel buffers that peers may still have mapped (DevBuf::retire)
    bool grid_round_entered = false;        // this call has taken part in the grid's agreement round (tmm_gemm tells the peers when it fails before)
    double link_h2d_gbs = 0, link_d2h_gbs = 0;  // host-link rates of THIS GPU while every GPU of the grid moves data (measured at attach; 0 = unknown)
    std::vector<tmm_context*> children;     // single-process multi-GPU: one child context per device, driven by host threads
    tmm_context* solo = nullptr;            // plain context on the first device for shapes too small to split

    // compute streams: [0] carries the phase-1 / streaming chain at the highest priority, the others (lower priorities)
    // carry independent column blocks, whose CTAs then only back-fill SM slots the chain leaves free
    int n_compute() const { int v = n_streams + 1; if (v > MAX_COMPUTE) v = MAX_COMPUTE; return v < 3 ? 3 : v; }

    cudaError_t get_event(cudaEvent_t* out) {
        if (ev_next == events.size()) {
            cudaEvent_t e;
            cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (r != cudaSuccess) return r;
            events.push_back(e);
        }
        *out = events[ev_next++];
        return cudaSuccess;
    }
    cudaError_t get_timing_event(cudaEvent_t* out) {
        if (tev_next == timing_events.size()) {
            cudaEvent_t e;
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
            timing_events.push_back(e);
        }
        *out = timing_events[tev_next++];
        return cudaSuccess;
    }
};

namespace tmm {

// ---- optional timeline (TMM_TRACE=1): every op gets a begin/end event; printed after the call ----
struct TraceScope {
    tmm_context* ctx; cudaStream_t st; size_t idx = (size_t)-1;
    TraceScope(tmm_context* c, cudaStream_t s, const char* name, int64_t a = 0, int64_t b = 0, int64_t d = 0) : ctx(c), st(s) {
        if (!ctx->trace) return;
        cudaEvent_t e0, e1;
        if (ctx->get_timing_event(&e0) != cudaSuccess || ctx->get_timing_event(&e1) != cudaSuccess) return;
        cudaEventRecord(e0, st);
        char buf[96];
        snprintf(buf, sizeof buf, "%s(%lld,%lld,%lld)", name, (long long)a, (long long)b, (long long)d);
        ctx->trace_ops.push_back({buf, e0, e1, st});
        idx = ctx->trace_ops.size() - 1;
    }
    ~TraceScope() { if (idx != (size_t)-1) cudaEventRecord(ctx->trace_ops[idx].e1, st); }
};


// true while the library itself is creating a context (children of a multi-device context): TMM_DEVICES must not recurse into them
bool creating_internal_context();
struct InternalContextScope { InternalContextScope(); ~InternalContextScope(); };

// ---- multi-GPU layer (tmm_dist.cu) ----
// Agree on the planning inputs across the grid (max block dims, min budget) and check that k / flags match everywhere.
// `local_rc` != 0 (this rank cannot run the call: bad argument, failed registration / allocation) is spread to all ranks, which then give
// the call up together; the call returns non-zero on every rank in that case.
int dist_agree(tmm_context* ctx, int64_t m, int64_t n, int64_t k, int flags, size_t budget, int64_t* m_plan, int64_t* n_plan, size_t* budget_min, int local_rc = 0);
// A rank that failed before it reached dist_agree takes part in that round all the same, so that its peers are not left waiting.
int dist_abort(tmm_context* ctx);
// Device bytes the staging rings need for shares of at most `share_bytes` gathered from up to `parts` ranks.
size_t dist_stage_bytes(size_t share_bytes, int parts);
int dist_reserve_stage(tmm_context* ctx, size_t share_bytes, int parts);
// All-gather a stored rows x cols sub-block (host src, leading dimension spitch elements) into dst (device, pitch dpitch
// elements): this rank uploads columns [lo, hi) of it (its share) on s_h2d, the shares are gathered on s_comm over `comm`
// and unpacked into place.  On return the tail of s_comm marks "dst complete".
int dist_exchange(tmm_context* ctx, Link& link, size_t es, const char* src, int64_t spitch, int64_t rows, int64_t cols, char* dst, int64_t dpitch,
                  int ring_slot = -1);
// Per call, after the panel buffer of this link is allocated: (re)map the peers' buffers (collective over the link).
int link_bind(tmm_context* ctx, Link& link, DevBuf& buf, bool local_ok = true);
// Once per call on every rank of the grid, before anything is enqueued: binds both links and spreads a local failure (local_rc != 0) to
// all ranks, so that the grid gives a call up together.  Returns 0 only if every rank is ready.
int grid_bind(tmm_context* ctx, int local_rc);
// Make `stream` wait until every peer's share of all exchanges issued so far on the link has arrived (direct links only).
int link_wait(tmm_context* ctx, Link& link, cudaStream_t stream);
// Streaming ring: tell the peers (after the work queued on `stream`) that one more ring exchange has been consumed here;
// dist_exchange(ring_slot >= 0) waits for the peers' acknowledgement of the exchange that last filled that slot.
int link_ack(tmm_context* ctx, Link& link, cudaStream_t stream);
void dist_release(tmm_context* ctx);
// Host-link probe (tmm_probe.cu): `bytes` up and `bytes` down at the same time on the current device, twice (the first pass warms up);
// `go` is called right before the timed pass so that several callers (ranks / devices) start together.  GB/s, 0 = failed.
struct LinkRates { double h2d = 0, d2h = 0; };
LinkRates probe_host_link(size_t bytes, const std::function<void()>& go);
// share g of `parts` over an extent: balanced split, [lo, hi)
inline void share_range(int64_t extent, int parts, int g, int64_t* lo, int64_t* hi) {
    const int64_t base = extent / parts, rem = extent % parts;
    *lo = g * base + (g < rem ? g : rem);
    *hi = *lo + base + (g < rem ? 1 : 0);
}
// upload share of rank g of a link over `cols` stored columns: by the link's per-call boundaries if it has any, else the balanced split
inline void upload_share(const Link& link, int64_t cols, int g, int64_t* lo, int64_t* hi) {
    if ((int)link.cut.size() != link.parts + 1) { share_range(cols, link.parts, g, lo, hi); return; }
    *lo = (int64_t)((__int128)cols * link.cut[g] / 1000000);
    *hi = (int64_t)((__int128)cols * link.cut[g + 1] / 1000000);
}
// Per call, after the grid has agreed on (m_plan, n_plan, k): set both links' share boundaries from the measured link rates (identical on all
// ranks of a link: integer arithmetic on agreed inputs).
void dist_set_shares(tmm_context* ctx, int64_t m_plan, int64_t n_plan, int64_t k, size_t es, bool c_up, bool c_down);
// single-process multi-GPU: run one call over the children of a parent context
int multi_gemm(tmm_context* parent, char ta, char tb, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t lda, const void* b,
               int64_t ldb, const void* beta, void* c, int64_t ldc, int pin, int copy_c_back);

}  // namespace tmm
