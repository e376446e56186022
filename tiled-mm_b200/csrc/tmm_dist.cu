// Multi-GPU layer: C tile-blocks are assigned to the GPUs of one box as a p_r x p_c grid (the reference is
// single-GPU, tiled_mm.cpp never selects a device).  Nothing is reduced across GPUs: k is never split, so the
// summation structure of every C element is the one-GPU one.  The only exchange is of read-only panels:
//   * the p_c GPUs of a grid row need the same A row-panel, the p_r GPUs of a grid column the same B column-panel;
//   * each GPU uploads a distinct share of every shared panel chunk over its OWN PCIe link (so aggregate host-link bandwidth scales with the
//     GPU count) and DMA-pushes it with the copy engines into the same place of every peer's panel over NVLink 5 / NVSwitch (mapped peer
//     memory: peer access inside a process, CUDA IPC across processes); arrival and acknowledgement counters are stream memory operations -
//     no SMs, no staging, no host on the data path.  Fallback data plane: ncclAllGather through a staging ring.
//   * control plane per call: a few rounds on a shared-memory board (microseconds of host time, with a deadline); NCCL at attach and as fallback.
//   * the shares follow the link rates measured at attach with every rank copying at once (a box's host links are neither independent nor alike).
// Two ways in:
//   (1) one process per GPU (torchrun): tmm_context_attach_grid() on each rank's context; tmm_gemm() then means "my block";
//   (2) one process, many GPUs: tmm_context_set_devices() turns a context into a parent whose tmm_gemm() partitions C
//       over child contexts, one host thread per GPU - the drop-in gpu::gemm then uses the whole box unchanged.
#include "tmm_internal.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

namespace tmm {
const char* last_error_cstr();
static thread_local int t_internal_context = 0;
bool creating_internal_context() { return t_internal_context > 0; }
InternalContextScope::InternalContextScope() { ++t_internal_context; }
InternalContextScope::~InternalContextScope() { --t_internal_context; }
bool debug_on() {
    static const bool on = [] { const char* v = getenv("TMM_DEBUG"); return v && v[0] == '1'; }();
    return on;
}

double debug_ms() {
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

namespace nccl {

const Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* override_path = getenv("TMM_NCCL_LIB");  // explicit path to an NCCL build (default: whatever libnccl.so.2 resolves to)
        for (const char* name : {override_path, "libnccl.so.2", "libnccl.so"}) {
            if (!name || !*name) continue;
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { a.why = "libnccl.so.2 not found (dlopen)"; return; }
        auto sym = [&](const char* n) { return dlsym(h, n); };
        a.GetUniqueId = reinterpret_cast<int (*)(UniqueId*)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<int (*)(Comm*, int, UniqueId, int)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<int (*)(Comm)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, Comm, cudaStream_t)>(sym("ncclAllGather"));
        a.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t)>(sym("ncclAllReduce"));
        a.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
        a.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GetErrorString;
        if (!a.ok) a.why = "libnccl.so.2 lacks a required entry point";
    });
    return a;
}

}  // namespace nccl

#define NC(x)                                        \
    do {                                             \
        int r__ = (x);                               \
        if (r__ != 0) return nccl_fail(r__, #x);     \
    } while (0)

size_t dist_stage_bytes(size_t share_bytes, int parts) {
    const size_t slot = (size_t)round_up64((int64_t)share_bytes, 512);
    return (size_t)STAGE_SLOTS * slot * (size_t)(1 + parts);
}

int dist_reserve_stage(tmm_context* ctx, size_t share_bytes, int parts) {
    const size_t slot = (size_t)round_up64((int64_t)std::max<size_t>(share_bytes, 512), 512);
    cudaError_t e;
    if ((e = ctx->stage_send.reserve((size_t)STAGE_SLOTS * slot)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(all-gather send ring)");
    if ((e = ctx->stage_recv.reserve((size_t)STAGE_SLOTS * slot * parts)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(all-gather recv ring)");
    ctx->stage_slot_bytes = slot;
    ctx->stage_next = 0;
    for (auto& ev : ctx->stage_send_free) ev = nullptr;  // events are per call (ctx->get_event pool)
    return TMM_OK;
}

// ---- stream memory operations (driver API, resolved at run time) -----------------------------------------------------
namespace {
typedef int (*StreamMemOpFn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
struct MemOps { StreamMemOpFn wait = nullptr, write = nullptr; bool ok = false; };
const MemOps& memops() {
    static MemOps m;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) m.wait = (StreamMemOpFn)p;
        p = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) m.write = (StreamMemOpFn)p;
        m.ok = m.wait && m.write;
    });
    return m;
}
int wait_geq(cudaStream_t st, const uint32_t* addr, uint32_t value) {
    int r = memops().wait(st, (unsigned long long)(uintptr_t)addr, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ: (int32)(*addr - value) >= 0 */);
    return r ? fail(TMM_ERR_CUDA, "GPU ERROR: cuStreamWaitValue32 failed (%d)", r) : TMM_OK;
}
int write_value(cudaStream_t st, uint32_t* addr, uint32_t value) {
    int r = memops().write(st, (unsigned long long)(uintptr_t)addr, value, 0);
    return r ? fail(TMM_ERR_CUDA, "GPU ERROR: cuStreamWriteValue32 failed (%d)", r) : TMM_OK;
}


// ---- control board (Link::board): per-call rounds through POSIX shared memory ----------------------------------------------------------
// Every rank of a link - a thread of this process or another process on the box - maps the same small segment.  A round is "publish my
// payload under the round's number, then read everybody's": an all-gather in a few microseconds of host time, with a deadline, no kernel
// launch and no stream.  Slots are double-buffered by round parity: a rank can only be one round ahead of the slowest reader, because
// finishing round r + 1 needs every peer's r + 1 publication, which a peer makes only after it has read round r.
constexpr int BOARD_MAX_PARTS = 16;
constexpr size_t BOARD_PAYLOAD = 128;
struct alignas(64) BoardSlot {
    std::atomic<uint64_t> seq;
    unsigned char pad[56];
    unsigned char payload[BOARD_PAYLOAD];
};
struct Board { BoardSlot slot[2][BOARD_MAX_PARTS]; };
static_assert(std::atomic<uint64_t>::is_always_lock_free, "board sequence numbers must be lock-free to work across processes");

#ifdef TMM_EMULATED
extern "C" void emul_collective(const void* group, int phase);  // tests/emul: a round orders the host threads of its ranks (race detector)
#define BOARD_HOOK(link, phase) emul_collective(reinterpret_cast<const void*>(static_cast<uintptr_t>((link).board_key)), phase)
#else
#define BOARD_HOOK(link, phase) ((void)0)
#endif

double dist_timeout_s() {
    const char* v = getenv("TMM_DIST_TIMEOUT_S");
    return (v && *v) ? atof(v) : 600.0;
}

uint64_t fnv1a(const void* data, size_t bytes) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < bytes; ++i) { h ^= static_cast<const unsigned char*>(data)[i]; h *= 1099511628211ull; }
    return h;
}

void board_name(const Link& link, char* out, size_t n) { snprintf(out, n, "/tmm_b200_%016llx_%u", (unsigned long long)link.board_key, (unsigned)getuid()); }

bool board_open(Link& link, const nccl::UniqueId& id) {
    const char* off = getenv("TMM_DIST_BOARD");
    if (off && off[0] == '0') return false;
    if (link.parts > BOARD_MAX_PARTS) return false;
    link.board_key = fnv1a(&id, sizeof id) ^ ((uint64_t)link.parts << 56);
    char name[64];
    board_name(link, name, sizeof name);
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return false;
    bool ok = ftruncate(fd, (off_t)sizeof(Board)) == 0;  // a fresh segment reads as zeros: sequence 0 = nothing published
    void* p = ok ? mmap(nullptr, sizeof(Board), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0) : MAP_FAILED;
    close(fd);
    if (p == MAP_FAILED) return false;
    link.board = p; link.board_bytes = sizeof(Board); link.board_round = 0;
    return true;
}

void board_unlink(const Link& link) {
    if (!link.board_key) return;
    char name[64];
    board_name(link, name, sizeof name);
    shm_unlink(name);
}
void board_close(Link& link) {
    if (link.board) munmap(link.board, link.board_bytes);
    link.board = nullptr; link.board_bytes = 0;
}

// one round: all-gather `bytes` (<= 128) per rank; `all` receives parts x bytes
int board_round(tmm_context* ctx, Link& link, const void* mine, size_t bytes, void* all) {
    if (link.broken) return fail(TMM_ERR_CUDA, "GPU ERROR: this GPU grid is out of step after a peer missed an earlier call; destroy the contexts");
    if (bytes > BOARD_PAYLOAD) return fail(TMM_ERR_INVALID, "internal: board payload too large");
    Board* b = static_cast<Board*>(link.board);
    const uint64_t round = ++link.board_round;
    BoardSlot& me = b->slot[round & 1][link.me];
    memcpy(me.payload, mine, bytes);
    BOARD_HOOK(link, 0);
    me.seq.store(round, std::memory_order_release);
    const double limit_s = dist_timeout_s();
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < link.parts; ++g) {
        BoardSlot& peer = b->slot[round & 1][g];
        for (uint64_t spins = 0; peer.seq.load(std::memory_order_acquire) < round; ++spins) {
            if (spins < 4096) continue;
            std::this_thread::yield();
            if ((spins & 1023) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
                link.broken = true;
                return fail(TMM_ERR_CUDA, "GPU ERROR: rank %d of this GPU grid link did not reach the call within %.0f s (device %d gives the call up; the grid is unusable)",
                            g, limit_s, ctx->device);
            }
        }
        memcpy(static_cast<char*>(all) + (size_t)g * bytes, peer.payload, bytes);
    }
    BOARD_HOOK(link, 1);
    return TMM_OK;
}

// a stream of the control plane must go idle, but a peer may have died: poll with the grid's deadline instead of blocking forever
int sync_with_deadline(tmm_context* ctx, cudaStream_t st) {
    const double limit_s = dist_timeout_s();
    const auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spins = 0;; ++spins) {
        cudaError_t e = cudaStreamQuery(st);
        if (e == cudaSuccess) return TMM_OK;
        if (e != cudaErrorNotReady) return cuda_fail(e, "cudaStreamQuery");
        if (spins > 4096) std::this_thread::yield();
        if ((spins & 1023) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s)
            return fail(TMM_ERR_CUDA, "GPU ERROR: a control collective of the GPU grid did not finish within %.0f s on device %d (a peer rank failed?)", limit_s, ctx->device);
    }
}

// what one rank tells the others about a buffer it owns
struct BufMsg {
    int64_t pid;
    int32_t dev, ok;
    uint64_t ptr, bytes;
    cudaIpcMemHandle_t handle;  // 64 bytes
};
static_assert(sizeof(BufMsg) == 96, "BufMsg layout");

// all-gather one BufMsg per rank over the link (tiny NCCL collective + host sync)
int gather_msgs(tmm_context* ctx, Link& link, const BufMsg& mine, std::vector<BufMsg>& all) {
    if (link.board) { all.resize(link.parts); return board_round(ctx, link, &mine, sizeof mine, all.data()); }
    const nccl::Api& nc = nccl::api();
    cudaError_t e;
    if ((e = ctx->dist_scratch.reserve(4096)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(dist scratch)");
    if ((size_t)(link.parts + 1) * sizeof(BufMsg) + 256 > ctx->dist_scratch.cap) return fail(TMM_ERR_INVALID, "grid link too wide");
    char* d = static_cast<char*>(ctx->dist_scratch.p) + 256;
    TMM_CU(cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->s_comm));
    NC(nc.AllGather(d, d + sizeof(BufMsg), sizeof(BufMsg), nccl::Int8, link.comm, ctx->s_comm));
    all.resize(link.parts);
    TMM_CU(cudaMemcpyAsync(all.data(), d + sizeof(BufMsg), sizeof(BufMsg) * link.parts, cudaMemcpyDeviceToHost, ctx->s_comm));
    return sync_with_deadline(ctx, ctx->s_comm);
}

BufMsg describe(tmm_context* ctx, void* p, size_t bytes) {
    BufMsg m;
    memset(&m, 0, sizeof m);
    m.pid = (int64_t)getpid(); m.dev = ctx->device; m.ptr = (uint64_t)(uintptr_t)p; m.bytes = bytes; m.ok = 1;
    if (p && cudaIpcGetMemHandle(&m.handle, p) != cudaSuccess) { cudaGetLastError(); m.ok = 0; }
    return m;
}

// map a peer's buffer into this process / device; returns nullptr on failure
void* map_peer(tmm_context* ctx, const BufMsg& msg, bool* via_ipc) {
    *via_ipc = false;
    if (!msg.ptr) return nullptr;
    // (TMM_DIST_FORCE_IPC=1 sends same-process peers through the IPC branch as well: only meaningful on the emulated runtime of tests/emul,
    //  where it lets the bookkeeping of imported handles be exercised without a second process; real CUDA cannot import its own export)
    static const bool force_ipc = [] { const char* v = getenv("TMM_DIST_FORCE_IPC"); return v && v[0] == '1'; }();
    if (msg.pid == (int64_t)getpid() && !force_ipc) {  // same process: peer access is enough
        if (msg.dev != ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(msg.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return nullptr; }
            cudaGetLastError();
        }
        return (void*)(uintptr_t)msg.ptr;
    }
    if (!msg.ok) return nullptr;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, msg.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *via_ipc = true;
    return p;
}

std::string key_of(const BufMsg& m) { return std::string(reinterpret_cast<const char*>(&m), sizeof m); }

// do all ranks of the link agree that a step worked?  (min-reduce of a flag; keeps the ranks on the same path)
int all_ok(tmm_context* ctx, Link& link, bool mine, bool* everyone) {
    if (link.board) {
        int32_t flag = mine ? 1 : 0, got[BOARD_MAX_PARTS];
        int rc = board_round(ctx, link, &flag, sizeof flag, got);
        if (rc) return rc;
        *everyone = true;
        for (int g = 0; g < link.parts; ++g) *everyone = *everyone && got[g] == 1;
        return TMM_OK;
    }
    const nccl::Api& nc = nccl::api();
    int32_t v = mine ? 1 : 0;
    int32_t* d = static_cast<int32_t*>(ctx->dist_scratch.p);
    TMM_CU(cudaMemcpyAsync(d, &v, sizeof v, cudaMemcpyHostToDevice, ctx->s_comm));
    NC(nc.AllReduce(d, d, 1, nccl::Int32, nccl::Min, link.comm, ctx->s_comm));
    TMM_CU(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, ctx->s_comm));
    { int rc = sync_with_deadline(ctx, ctx->s_comm); if (rc) return rc; }
    *everyone = v == 1;
    return TMM_OK;
}

void link_unmap(Link& link) {
    for (size_t g = 0; g < link.peer_base.size(); ++g)
        if (link.peer_ipc[g] && link.peer_base[g]) cudaIpcCloseMemHandle(link.peer_base[g]);
    link.peer_base.clear(); link.peer_key.clear(); link.peer_ipc.clear();
}

// at attach: my flag block, the peers' flag blocks, and the decision DMA push vs NCCL staging
int link_setup(tmm_context* ctx, Link& link, const nccl::UniqueId& id) {
    if (!link.active()) return TMM_OK;
    cudaError_t e;
    if ((e = ctx->dist_scratch.reserve(4096)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(dist scratch)");
    {
        // control board: used only if EVERY rank of the link could map it (that question itself is still settled over NCCL)
        const bool opened = board_open(link, id);
        void* mapped = link.board;
        link.board = nullptr;
        bool everyone = false;
        int rc = all_ok(ctx, link, opened, &everyone);
        link.board = mapped;
        if (link.me == 0) board_unlink(link);  // every rank that could open the segment has it mapped by now: the name can go
        if (rc || !everyone) board_close(link);
        if (rc) return rc;
        TMM_DBG("dev %d link of %d ranks (me %d): control plane %s", ctx->device, link.parts, link.me, link.board ? "shared-memory board" : "NCCL collectives");
    }
    const char* force = getenv("TMM_DIST_NCCL");
    bool ok = memops().ok && !(force && force[0] == '1');
    void* fl = nullptr;
    if ((e = cudaMalloc(&fl, 1024)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(flags)");
    TMM_CU(cudaMemset(fl, 0, 1024));
    link.flags = static_cast<uint32_t*>(fl);
    std::vector<BufMsg> all;
    int rc = gather_msgs(ctx, link, describe(ctx, fl, 1024), all);
    if (rc) return rc;
    link.peer_flags.assign(link.parts, nullptr);
    link.peer_flags_ipc.assign(link.parts, false);
    for (int g = 0; g < link.parts && ok; ++g) {
        if (g == link.me) { link.peer_flags[g] = link.flags; continue; }
        bool via = false;
        link.peer_flags[g] = static_cast<uint32_t*>(map_peer(ctx, all[g], &via));
        link.peer_flags_ipc[g] = via;
        if (!link.peer_flags[g]) ok = false;
    }
    bool everyone = false;
    if ((rc = all_ok(ctx, link, ok, &everyone))) return rc;
    link.direct = everyone;
    TMM_DBG("dev %d link of %d ranks (me %d): %s", ctx->device, link.parts, link.me, link.direct ? "DMA push over mapped peer memory" : "NCCL all-gather staging");
    link.peer_base.assign(link.parts, nullptr); link.peer_key.assign(link.parts, std::string()); link.peer_ipc.assign(link.parts, false);
    return TMM_OK;
}

void link_teardown(Link& link) {
    const nccl::Api& nc = nccl::api();
    link_unmap(link);
    for (size_t g = 0; g < link.peer_flags.size(); ++g)  // imported flag blocks of the peers
        if (g < link.peer_flags_ipc.size() && link.peer_flags_ipc[g] && link.peer_flags[g]) cudaIpcCloseMemHandle(link.peer_flags[g]);
    if (link.flags) cudaFree(link.flags);
    if (link.comm && nc.ok) nc.CommDestroy(link.comm);
    board_close(link);
    link = Link{};
}
}  // namespace


// ---- host-link probe ------------------------------------------------------------------------------------------------------------------
// What a GPU's PCIe link delivers depends on how many of the box's GPUs move data at the same time: on the 8-GPU node of this pool one
// B200 alone copies 55 GB/s each way, eight at once get 8 - 12 GB/s each way (profiles/r2_probe_8gpu.txt: shared uplinks / host memory).
// The scheduler's chunk and block sizes are a function of these rates, so a grid measures them once, with all its ranks copying at the
// same time, instead of assuming the single-GPU figure.
// Collective over the link, once per call and on EVERY rank of it, whatever happened locally: `local_ok == false` (an allocation failed
// here, or another link already reported a failure) is spread to the peers, so that all ranks give the call up together instead of
// some of them enqueueing work that waits for shares which will never come.
int link_bind(tmm_context* ctx, Link& link, DevBuf& buf, bool local_ok) {
    if (!link.active()) return local_ok ? TMM_OK : TMM_ERR_CUDA;
    for (auto& ev : link.slot_pushed) ev = nullptr;  // events come from the per-call pool
    bool ok = local_ok;
    int rc;
    if (link.direct) {
        std::vector<BufMsg> all;
        rc = gather_msgs(ctx, link, describe(ctx, local_ok ? buf.p : nullptr, local_ok ? buf.cap : 0), all);
        if (rc) return rc;
        for (int g = 0; g < link.parts; ++g) {
            if (g == link.me) { link.peer_base[g] = static_cast<char*>(buf.p); continue; }
            if (!all[g].ptr) { ok = false; continue; }  // that peer could not allocate its panel
            if (!local_ok) continue;
            const std::string key = key_of(all[g]);
            if (key == link.peer_key[g] && link.peer_base[g]) continue;  // same allocation as last call: mapping still valid
            if (link.peer_ipc[g] && link.peer_base[g]) cudaIpcCloseMemHandle(link.peer_base[g]);
            bool via = false;
            link.peer_base[g] = static_cast<char*>(map_peer(ctx, all[g], &via));
            link.peer_ipc[g] = via; link.peer_key[g] = key;
            if (!link.peer_base[g]) ok = false;
        }
        link.local_base = static_cast<char*>(buf.p);
    }
    bool everyone = false;
    if ((rc = all_ok(ctx, link, ok, &everyone))) return rc;
    if (!everyone)
        return local_ok ? fail(TMM_ERR_CUDA, "GPU ERROR: a peer GPU of the grid could not allocate or map its panel buffer for this call (CUDA IPC / peer access / "
                                             "out of memory); set TMM_DIST_NCCL=1 to use NCCL staging if mapping is the problem")
                        : TMM_ERR_CUDA;
    return TMM_OK;
}

// Both links, always both, rows first: a failure anywhere in the grid reaches every rank (the failing rank's row learns it in the first
// round and passes it on to all columns in the second).  Returns 0 only if every rank of the grid is ready to enqueue.
int grid_bind(tmm_context* ctx, int local_rc) {
    const int r1 = link_bind(ctx, ctx->grid.rowl, ctx->buf_a, local_rc == TMM_OK);
    const int r2 = link_bind(ctx, ctx->grid.coll, ctx->buf_b, local_rc == TMM_OK && r1 == TMM_OK);
    if (local_rc) return local_rc;
    return r1 ? r1 : r2;
}

int link_wait(tmm_context* ctx, Link& link, cudaStream_t stream) {
    if (!link.active() || !link.direct || link.sent == 0) return TMM_OK;
    for (int g = 0; g < link.parts; ++g) {
        if (g == link.me) continue;
        int rc = wait_geq(stream, link.flags + g, link.sent);
        if (rc) return rc;
    }
    return TMM_OK;
}

int link_ack(tmm_context* ctx, Link& link, cudaStream_t stream) {
    if (!link.active() || !link.direct) return TMM_OK;
    ++link.acked;
    uint32_t* word = link.flags + 2 * link.parts + 1;
    int rc = write_value(stream, word, link.acked);
    if (rc) return rc;
    for (int g = 0; g < link.parts; ++g) {
        if (g == link.me) continue;
        TMM_CU(cudaMemcpyAsync(link.peer_flags[g] + link.parts + link.me, word, 4, cudaMemcpyDeviceToDevice, stream));
    }
    return TMM_OK;
}

static void reduce_max8(int64_t* v, const int64_t* all, int parts) {
    for (int g = 0; g < parts; ++g)
        for (int i = 0; i < 8; ++i) v[i] = std::max(v[i], all[g * 8 + i]);
}

// element-wise maximum of 8 words over the whole grid (columns, then rows: after both every rank holds the grid-wide maxima); a rendezvous
static int grid_reduce_max8(tmm_context* ctx, int64_t* v) {
    const nccl::Api& nc = nccl::api();
    Grid& gr = ctx->grid;
    for (Link* link : {&gr.coll, &gr.rowl}) {
        if (!link->active()) continue;
        if (link->board) {  // (a link's ranks agree on having the board: link_setup)
            int64_t all[BOARD_MAX_PARTS * 8];
            int rc = board_round(ctx, *link, v, 8 * sizeof(int64_t), all);
            if (rc) return rc;
            reduce_max8(v, all, link->parts);
            continue;
        }
        if (!nc.ok) return fail(TMM_ERR_CUDA, "GPU ERROR: NCCL unavailable: %s", nc.why);
        cudaError_t e;
        if ((e = ctx->dist_scratch.reserve(4096)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(dist scratch)");
        int64_t* d = static_cast<int64_t*>(ctx->dist_scratch.p);
        TMM_CU(cudaMemcpyAsync(d, v, 8 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->s_comm));
        NC(nc.AllReduce(d, d, 8, nccl::Int64, nccl::Max, link->comm, ctx->s_comm));
        TMM_CU(cudaMemcpyAsync(v, d, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->s_comm));
        int rc = sync_with_deadline(ctx, ctx->s_comm);
        if (rc) return rc;
    }
    return TMM_OK;
}

int dist_agree(tmm_context* ctx, int64_t m, int64_t n, int64_t k, int flags, size_t budget, int64_t* m_plan, int64_t* n_plan, size_t* budget_min, int local_rc) {
    ctx->grid_round_entered = true;
    // max-reduce {m, n, k, -k, flags, -flags, -budget, failed}: maxima give the planning block and the smallest budget, the +/- pairs show
    // (on every rank alike, so all ranks fail together instead of hanging) whether k and the flags agree, and the last word spreads a
    // rank-local failure.  Being a rendezvous it is also the barrier that keeps a fast rank from pushing panels of the next call into a
    // peer that is still computing the previous one (tmm_gemm is synchronous: a rank that arrives here has finished its previous call).
    int64_t v[8] = {m, n, k, -k, (int64_t)flags, -(int64_t)flags, -(int64_t)std::min<size_t>(budget, (size_t)INT64_MAX), local_rc ? 1 : 0};
    TMM_DBG("dev %d agree: m %lld n %lld k %lld status %d", ctx->device, (long long)m, (long long)n, (long long)k, local_rc);
    {
        int rc = grid_reduce_max8(ctx, v);
        if (rc) return rc;
    }
    if (local_rc) return local_rc;
    if (v[7] != 0) return fail(TMM_ERR_CUDA, "GPU grid: another rank could not start this call (bad argument, failed registration or allocation there); the call is given up on every rank");
    if (v[2] != -v[3]) return fail(TMM_ERR_INVALID, "GPU grid: ranks disagree on k (%lld vs %lld)", (long long)v[2], (long long)-v[3]);
    if (v[4] != -v[5]) return fail(TMM_ERR_INVALID, "GPU grid: ranks disagree on trans / beta==0 / copy_c_back / alpha==0");
    *m_plan = v[0]; *n_plan = v[1]; *budget_min = (size_t)(-v[6]);
    TMM_DBG("dev %d agreed: m_plan %lld n_plan %lld budget %zu", ctx->device, (long long)*m_plan, (long long)*n_plan, *budget_min);
    return TMM_OK;
}

int dist_abort(tmm_context* ctx) {
    int64_t mp, np;
    size_t b;
    const std::string keep = last_error_cstr();  // the caller reports ITS error, not the round's
    dist_agree(ctx, 0, 0, 0, 0, 0, &mp, &np, &b, TMM_ERR_INVALID);
    fail(TMM_ERR_INVALID, "%s", keep.c_str());
    return TMM_OK;
}

// DMA push: my share goes host -> my panel (in place), then panel -> the same place in every peer's panel over NVLink.
static int dist_push(tmm_context* ctx, Link& link, size_t es, const char* src, int64_t spitch, int64_t rows, int64_t cols, char* dst, int64_t dpitch,
                     int ring_slot) {
    const int parts = link.parts, me = link.me;
    int64_t lo, hi;
    upload_share(link, cols, me, &lo, &hi);
    if (ring_slot >= 0) {
        // the peers must have consumed the exchange that last filled this ring slot before it is overwritten
        const uint32_t last = link.slot_last[ring_slot & 7];
        if (last)
            for (int g = 0; g < parts; ++g)
                if (g != me) { int rc = wait_geq(ctx->s_comm, link.flags + parts + g, last); if (rc) return rc; }
        link.slot_last[ring_slot & 7] = ++link.ring_sent;
        // ... and my own previous push out of this slot must have read my share before the next upload overwrites it (found by the
        // race detector of tests/emul: the GEMM that frees the slot waits for the peers' shares, not for my outgoing copies)
        if (link.slot_pushed[ring_slot & 7]) TMM_CU(cudaStreamWaitEvent(ctx->s_h2d, link.slot_pushed[ring_slot & 7], 0));
    }
    const uint32_t x = ++link.sent;
    TMM_DBG("dev %d push #%u: %lld x %lld, my columns [%lld,%lld), %d peers, slot %d", ctx->device, x, (long long)rows, (long long)cols, (long long)lo, (long long)hi, parts - 1, ring_slot);
    char* mine = dst + (size_t)lo * dpitch * es;
    if (hi > lo) {
        TMM_CU(cudaMemcpy2DAsync(mine, (size_t)dpitch * es, src + (size_t)lo * spitch * es, (size_t)spitch * es, (size_t)rows * es, (size_t)(hi - lo),
                                 cudaMemcpyHostToDevice, ctx->s_h2d));
        ctx->stats.h2d_bytes += (uint64_t)rows * (hi - lo) * es;
        ctx->stats.h2d_copies++;
    }
    cudaEvent_t up;
    TMM_CU(ctx->get_event(&up));
    TMM_CU(cudaEventRecord(up, ctx->s_h2d));
    TMM_CU(cudaStreamWaitEvent(ctx->s_comm, up, 0));
    const size_t off = (size_t)(mine - link.local_base);
    uint32_t* word = link.flags + 2 * parts;
    {
        TraceScope ts(ctx, ctx->s_comm, "push", rows, hi - lo, parts - 1);
        for (int g = 0; g < parts; ++g) {
            if (g == me) continue;
            if (hi > lo) {
                char* pd = link.peer_base[g] + off;
                if (rows == dpitch) TMM_CU(cudaMemcpyAsync(pd, mine, (size_t)rows * (hi - lo) * es, cudaMemcpyDeviceToDevice, ctx->s_comm));
                else TMM_CU(cudaMemcpy2DAsync(pd, (size_t)dpitch * es, mine, (size_t)dpitch * es, (size_t)rows * es, (size_t)(hi - lo), cudaMemcpyDeviceToDevice, ctx->s_comm));
                ctx->stats.peer_bytes += (uint64_t)rows * (hi - lo) * es;
            }
        }
        // arrival counter: raised in every peer after the data (stream order on s_comm)
        int rc = write_value(ctx->s_comm, word, x);
        if (rc) return rc;
        for (int g = 0; g < parts; ++g)
            if (g != me) TMM_CU(cudaMemcpyAsync(link.peer_flags[g] + me, word, 4, cudaMemcpyDeviceToDevice, ctx->s_comm));
    }
    if (ring_slot >= 0) {
        TMM_CU(ctx->get_event(&link.slot_pushed[ring_slot & 7]));
        TMM_CU(cudaEventRecord(link.slot_pushed[ring_slot & 7], ctx->s_comm));
    }
    return TMM_OK;
}

int dist_exchange(tmm_context* ctx, Link& link, size_t es, const char* src, int64_t spitch, int64_t rows, int64_t cols, char* dst, int64_t dpitch,
                  int ring_slot) {
    if (rows <= 0 || cols <= 0) return TMM_OK;
    if (link.direct) return dist_push(ctx, link, es, src, spitch, rows, cols, dst, dpitch, ring_slot);
    const nccl::Api& nc = nccl::api();
    const int parts = link.parts, me = link.me;
    const int64_t max_cols = (cols + parts - 1) / parts;
    const size_t share_bytes = (size_t)rows * (size_t)max_cols * es;
    if (share_bytes > ctx->stage_slot_bytes) return fail(TMM_ERR_INVALID, "internal: all-gather share %zu B exceeds the staging slot %zu B", share_bytes, ctx->stage_slot_bytes);
    const int slot = ctx->stage_next;
    TMM_DBG("dev %d exchange: %lld x %lld over %d ranks (me %d), share %zu B, slot %d", ctx->device, (long long)rows, (long long)cols, parts, me, share_bytes, slot);
    ctx->stage_next = (slot + 1) % STAGE_SLOTS;
    char* send = static_cast<char*>(ctx->stage_send.p) + (size_t)slot * ctx->stage_slot_bytes;
    char* recv = static_cast<char*>(ctx->stage_recv.p) + (size_t)slot * ctx->stage_slot_bytes * parts;
    // 1. my share (a range of stored columns: long contiguous runs for the DMA engine) -> send slot, compact
    int64_t lo, hi;
    share_range(cols, parts, me, &lo, &hi);
    if (ctx->stage_send_free[slot]) TMM_CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->stage_send_free[slot], 0));
    if (hi > lo) {
        TMM_CU(cudaMemcpy2DAsync(send, (size_t)rows * es, src + (size_t)lo * spitch * es, (size_t)spitch * es, (size_t)rows * es, (size_t)(hi - lo),
                                 cudaMemcpyHostToDevice, ctx->s_h2d));
        ctx->stats.h2d_bytes += (uint64_t)rows * (hi - lo) * es;
        ctx->stats.h2d_copies++;
    }
    cudaEvent_t up;
    TMM_CU(ctx->get_event(&up));
    TMM_CU(cudaEventRecord(up, ctx->s_h2d));
    TMM_CU(cudaStreamWaitEvent(ctx->s_comm, up, 0));
    // 2. all-gather the shares over NVLink (equal counts: the largest share; shorter shares carry padding that is never unpacked)
    {
        TraceScope ts(ctx, ctx->s_comm, "allgather", rows, cols, parts);
        NC(nc.AllGather(send, recv, share_bytes, nccl::Int8, link.comm, ctx->s_comm));
    }
    TMM_CU(ctx->get_event(&ctx->stage_send_free[slot]));
    TMM_CU(cudaEventRecord(ctx->stage_send_free[slot], ctx->s_comm));
    // 3. unpack into the panel (device-to-device 2-D copies re-pitch for free; stream order protects the recv slot)
    TraceScope ts_unpack(ctx, ctx->s_comm, "unpack", rows, cols, parts);
    for (int g = 0; g < parts; ++g) {
        share_range(cols, parts, g, &lo, &hi);
        if (hi <= lo) continue;
        TMM_CU(cudaMemcpy2DAsync(dst + (size_t)lo * dpitch * es, (size_t)dpitch * es, recv + (size_t)g * share_bytes, (size_t)rows * es, (size_t)rows * es,
                                 (size_t)(hi - lo), cudaMemcpyDeviceToDevice, ctx->s_comm));
        if (g != me) ctx->stats.peer_bytes += (uint64_t)rows * (hi - lo) * es;
    }
    return TMM_OK;
}

// all-gather four words per rank over a link (attach time only)
static int link_gather4(tmm_context* ctx, Link& link, const int64_t* mine, std::vector<int64_t>& all) {
    all.assign((size_t)link.parts * 4, 0);
    if (!link.active()) return TMM_OK;
    if (link.board) return board_round(ctx, link, mine, 4 * sizeof(int64_t), all.data());
    const nccl::Api& nc = nccl::api();
    char* d = static_cast<char*>(ctx->dist_scratch.p) + 256;
    TMM_CU(cudaMemcpyAsync(d, mine, 32, cudaMemcpyHostToDevice, ctx->s_comm));
    NC(nc.AllGather(d, d + 32, 32, nccl::Int8, link.comm, ctx->s_comm));
    TMM_CU(cudaMemcpyAsync(all.data(), d + 32, (size_t)32 * link.parts, cudaMemcpyDeviceToHost, ctx->s_comm));
    return sync_with_deadline(ctx, ctx->s_comm);
}

// Shares that equalise the ranks' host-link time for this call.  Rank g of a link carries, besides its share f_g of this link's panel (P bytes),
// a fixed load: its (roughly equal) share of the other link's panel and its C block, up and / or down.  With measured rates u_g (up) and d_g (down),
//     T = (fixed_up_g + f_g P) / u_g + fixed_down_g / d_g   for all g,   sum f_g = 1
// has the closed form below; shares are floored at 5 % (every rank keeps a part in the exchange) and renormalised.  Millionths, integer inputs.
static void balance_link(Link& link, double panel_bytes, double other_up_bytes, double c_up_bytes, double c_down_bytes) {
    link.cut.clear();
    const int parts = link.parts;
    if (!link.active() || !link.direct || (int)link.rate_up.size() != parts || (int)link.rate_down.size() != parts || panel_bytes <= 0) return;
    for (int g = 0; g < parts; ++g)
        if (link.rate_up[g] <= 0 || link.rate_down[g] <= 0) return;
    double sum_u = 0, sum_fixed = 0;
    for (int g = 0; g < parts; ++g) {
        const double u = link.rate_up[g], d = link.rate_down[g];
        sum_u += u;
        sum_fixed += (other_up_bytes + c_up_bytes) + u * c_down_bytes / d;
    }
    const double T = (panel_bytes + sum_fixed) / sum_u;  // (bytes per MB/s: the unit cancels in f)
    std::vector<double> f(parts);
    double total = 0;
    for (int g = 0; g < parts; ++g) {
        const double u = link.rate_up[g], d = link.rate_down[g];
        f[g] = ((T - c_down_bytes / d) * u - (other_up_bytes + c_up_bytes)) / panel_bytes;
        if (!(f[g] > 0.05)) f[g] = 0.05;
        total += f[g];
    }
    link.cut.assign(parts + 1, 0);
    double acc = 0;
    for (int g = 0; g < parts; ++g) {
        acc += f[g] / total;
        link.cut[g + 1] = g == parts - 1 ? 1000000 : std::min<int64_t>(1000000, (int64_t)(acc * 1e6 + 0.5));
    }
}

void dist_set_shares(tmm_context* ctx, int64_t m_plan, int64_t n_plan, int64_t k, size_t es, bool c_up, bool c_down) {
    Grid& g = ctx->grid;
    g.rowl.cut.clear(); g.coll.cut.clear();
    static const bool off = [] { const char* v = getenv("TMM_DIST_BALANCE"); return v && v[0] == '0'; }();
    if (off || !g.active()) return;
    const double pa = (double)m_plan * (double)k * (double)es, pb = (double)k * (double)n_plan * (double)es, pc = (double)m_plan * (double)n_plan * (double)es;
    // the row link shares the A panel (its ranks also upload ~1/p_r of their B panel), the column link the B panel (~1/p_c of their A panel)
    balance_link(g.rowl, pa, pb / std::max(1, g.pr), c_up ? pc : 0.0, c_down ? pc : 0.0);
    balance_link(g.coll, pb, pa / std::max(1, g.pc), c_up ? pc : 0.0, c_down ? pc : 0.0);
    if (debug_on()) {
        for (Link* l : {&g.rowl, &g.coll})
            if (!l->cut.empty()) {
                std::string sline;
                for (int i = 0; i < l->parts; ++i) sline += " " + std::to_string((l->cut[i + 1] - l->cut[i]) / 10000) + "%";
                TMM_DBG("dev %d %s link upload shares:%s", ctx->device, l == &g.rowl ? "row (A)" : "column (B)", sline.c_str());
            }
    }
}

void dist_release(tmm_context* ctx) {
    ctx->buf_a.retire = ctx->buf_b.retire = nullptr;
    link_teardown(ctx->grid.rowl);
    link_teardown(ctx->grid.coll);
    ctx->grid = Grid{};
    ctx->stage_send.release(); ctx->stage_recv.release(); ctx->dist_scratch.release();
}

static int attach(tmm_context* ctx, int pr, int pc, int row, int col, const nccl::UniqueId* row_id, const nccl::UniqueId* col_id) {
    const nccl::Api& nc = nccl::api();
    if (!nc.ok) return fail(TMM_ERR_CUDA, "GPU ERROR: NCCL unavailable: %s", nc.why);
    DeviceGuard guard(ctx->device);
    dist_release(ctx);
    Grid& g = ctx->grid;
    g.pr = pr; g.pc = pc; g.row = row; g.col = col;
    g.rowl.parts = pc; g.rowl.me = col;
    g.coll.parts = pr; g.coll.me = row;
    TMM_DBG("dev %d attach %dx%d at (%d,%d)", ctx->device, pr, pc, row, col);
    if (pc > 1) NC(nc.CommInitRank(&g.rowl.comm, pc, *row_id, col));
    if (pr > 1) NC(nc.CommInitRank(&g.coll.comm, pr, *col_id, row));
    int rc = link_setup(ctx, g.rowl, *row_id);
    if (!rc) rc = link_setup(ctx, g.coll, *col_id);
    // panel buffers that a peer may have mapped are retired, not freed, when they are outgrown (freed after the next link_bind)
    ctx->buf_a.retire = (g.rowl.active() && g.rowl.direct) ? &ctx->retired : nullptr;
    ctx->buf_b.retire = (g.coll.active() && g.coll.direct) ? &ctx->retired : nullptr;
    ctx->budget_cached = 0;
    ctx->link_h2d_gbs = ctx->link_d2h_gbs = 0;
    {
        int64_t mine[4] = {0, 0, 0, 0};  // this rank's own link figures in MB/s (0 = unknown)
        bool have = false;
#ifndef TMM_EMULATED
        // host-link rates with every rank of the grid copying at once; all ranks then plan with the slowest link's figures (TMM_DIST_PROBE=0: skip)
        const char* pv = getenv("TMM_DIST_PROBE");
        if (!rc && !(pv && pv[0] == '0')) {
            int64_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int rc_go = TMM_OK;
            const LinkRates lr = probe_host_link((size_t)64 << 20, [&] { rc_go = grid_reduce_max8(ctx, v); });
            const bool valid = rc_go == TMM_OK && lr.h2d > 0 && lr.d2h > 0;
            int64_t w[8] = {valid ? -(int64_t)(lr.h2d * 1e3) : 0, valid ? -(int64_t)(lr.d2h * 1e3) : 0, valid ? 0 : 1, 0, 0, 0, 0, 0};
            rc = rc_go ? rc_go : grid_reduce_max8(ctx, w);
            if (!rc && w[2] == 0) { ctx->link_h2d_gbs = (double)(-w[0]) * 1e-3; ctx->link_d2h_gbs = (double)(-w[1]) * 1e-3; }
            if (valid) { mine[0] = (int64_t)(lr.h2d * 1e3); mine[1] = (int64_t)(lr.d2h * 1e3); }
            have = true;
            TMM_DBG("dev %d host link with the whole grid active: %.1f GB/s up, %.1f GB/s down here; grid minimum %.1f / %.1f", ctx->device, lr.h2d, lr.d2h,
                    ctx->link_h2d_gbs, ctx->link_d2h_gbs);
        }
#else
        // the emulated runtime has no links to measure; TMM_EMUL_LINK_RATES=1 gives every device a different made-up figure so that the
        // rate-balanced upload shares are exercised by the CPU suite
        const char* fv = getenv("TMM_EMUL_LINK_RATES");
        if (!rc && fv && fv[0] == '1') { mine[0] = 8000 + 5000 * (ctx->device % 3); mine[1] = 9000 + 2500 * ((ctx->device + 1) % 4); have = true; }
#endif
        // every rank's own figures, per link: the upload shares of a call follow them (dist_set_shares)
        for (Link* l : {&g.rowl, &g.coll}) {
            l->rate_up.clear(); l->rate_down.clear(); l->cut.clear();
            if (rc || !have || !l->active()) continue;
            std::vector<int64_t> all;
            rc = link_gather4(ctx, *l, mine, all);
            if (!rc)
                for (int i = 0; i < l->parts; ++i) { l->rate_up.push_back((int32_t)all[(size_t)i * 4]); l->rate_down.push_back((int32_t)all[(size_t)i * 4 + 1]); }
        }
    }
    TMM_DBG("dev %d attached rc %d", ctx->device, rc);
    return rc;
}

// ---- single process, many GPUs -------------------------------------------------------------------------------------
int multi_gemm(tmm_context* parent, char ta, char tb, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t lda, const void* b,
               int64_t ldb, const void* beta, void* c, int64_t ldc, int pin, int copy_c_back) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int nd = (int)parent->children.size();
    const Grid& g0 = parent->children[0]->grid;
    const int pr = g0.pr, pc = g0.pc;
    const size_t es = dtype_size(parent->dtype);
    const char TA = (char)std::toupper((unsigned char)ta), TB = (char)std::toupper((unsigned char)tb);
    parent->stats = tmm_call_stats{};
    if (!copy_c_back) {
        // The result has to be ONE column-major device matrix (get_full_device_buffer_c, ld = m): such a call runs on the plain context
        // of the first device, and the parent's device-C accessors follow it there.  The grid is for results that go back to the host.
        tmm_context* one = parent->children[0]->grid.active() ? parent->solo : parent->children[0];
        const int rc1 = tmm_gemm(one, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin, copy_c_back);
        parent->stats = one->stats;
        return rc1;
    }
    if (m < 0 || n < 0 || k < 0) return fail(TMM_ERR_INVALID, "negative dimension");
    if (m == 0 || n == 0) return TMM_OK;
    if ((TA != 'N' && TA != 'T' && TA != 'C') || (TB != 'N' && TB != 'T' && TB != 'C')) return fail(TMM_ERR_INVALID, "trans must be one of N, T, C (got '%c','%c')", ta, tb);
    if (m < pr || n < pc) {  // fewer rows / columns than grid rows / columns: not worth a grid, run on the first device
        return tmm_gemm(parent->children[0]->grid.active() ? parent->solo : parent->children[0], ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin, copy_c_back);
    }
    // page-lock once for all devices (portable), not once per child
    std::vector<void*> pinned;
    if (pin) {
        const int64_t a_cols = TA == 'N' ? k : m, b_cols = TB == 'N' ? n : k;
        struct { const void* p; size_t bytes; } regs[3] = {{a, (size_t)lda * a_cols * es}, {b, (size_t)ldb * b_cols * es}, {c, (size_t)ldc * n * es}};
        for (auto& r : regs) {
            if (!r.p || !r.bytes) continue;
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, r.p) == cudaSuccess && attr.type == cudaMemoryTypeHost) continue;
            cudaGetLastError();
            cudaError_t e = cudaHostRegister(const_cast<void*>(r.p), r.bytes, cudaHostRegisterPortable);
            if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); continue; }
            if (e != cudaSuccess) { for (void* q : pinned) cudaHostUnregister(q); return cuda_fail(e, "cudaHostRegister"); }
            pinned.push_back(const_cast<void*>(r.p));
        }
    }
    std::vector<int> rcs(nd, TMM_OK);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> threads;
    for (int d = 0; d < nd; ++d) {
        threads.emplace_back([&, d] {
            tmm_context* ch = parent->children[d];
            cudaSetDevice(ch->device);
            int64_t i0, i1, j0, j1;
            share_range(m, pr, ch->grid.row, &i0, &i1);
            share_range(n, pc, ch->grid.col, &j0, &j1);
            const char* ap = static_cast<const char*>(a) + (TA == 'N' ? (size_t)i0 : (size_t)i0 * lda) * es;
            const char* bp = static_cast<const char*>(b) + (TB == 'N' ? (size_t)j0 * ldb : (size_t)j0) * es;
            char* cp = static_cast<char*>(c) + ((size_t)j0 * ldc + i0) * es;
            TMM_DBG("dev %d child gemm block rows [%lld,%lld) cols [%lld,%lld)", ch->device, (long long)i0, (long long)i1, (long long)j0, (long long)j1);
            rcs[d] = tmm_gemm(ch, ta, tb, i1 - i0, j1 - j0, k, alpha, ap, lda, bp, ldb, beta, cp, ldc, 0, 1);
            TMM_DBG("dev %d child gemm done rc %d", ch->device, rcs[d]);
            if (rcs[d]) errs[d] = last_error_cstr();
        });
    }
    for (auto& t : threads) t.join();
    for (void* q : pinned) cudaHostUnregister(q);
    int rc = TMM_OK;
    for (int d = 0; d < nd; ++d) {
        const tmm_call_stats& s = parent->children[d]->stats;
        parent->stats.h2d_bytes += s.h2d_bytes; parent->stats.d2h_bytes += s.d2h_bytes; parent->stats.peer_bytes += s.peer_bytes;
        parent->stats.kernel_launches += s.kernel_launches; parent->stats.h2d_copies += s.h2d_copies; parent->stats.d2h_copies += s.d2h_copies;
        parent->stats.kernel_ms = std::max(parent->stats.kernel_ms, s.kernel_ms);
        parent->stats.regime = s.regime; parent->stats.c_blocks += s.c_blocks; parent->stats.k_chunks = s.k_chunks;
        if (rcs[d] && !rc) { rc = rcs[d]; fail(rc, "device %d: %s", parent->children[d]->device, errs[d].c_str()); }
    }
    parent->stats.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return rc;
}

}  // namespace tmm

extern "C" {

int tmm_grid_shape(int n_gpus, int* grid_rows, int* grid_cols) {
    if (n_gpus < 1 || !grid_rows || !grid_cols) return tmm::fail(TMM_ERR_INVALID, "grid_shape: bad argument");
    int pr = 1;
    for (int d = 1; d * d <= n_gpus; ++d)
        if (n_gpus % d == 0) pr = d;  // the most square factorisation with p_r <= p_c: 1->1x1, 2->1x2, 4->2x2, 8->2x4
    *grid_rows = pr; *grid_cols = n_gpus / pr;
    return TMM_OK;
}

int tmm_share_range(int64_t extent, int parts, int index, int64_t* lo, int64_t* hi) {
    if (extent < 0 || parts < 1 || index < 0 || index >= parts || !lo || !hi) return tmm::fail(TMM_ERR_INVALID, "share_range: bad argument");
    tmm::share_range(extent, parts, index, lo, hi);
    return TMM_OK;
}

int tmm_dist_unique_id(void* out128) {
    if (!out128) return tmm::fail(TMM_ERR_INVALID, "out is null");
    const tmm::nccl::Api& nc = tmm::nccl::api();
    if (!nc.ok) return tmm::fail(TMM_ERR_CUDA, "GPU ERROR: NCCL unavailable: %s", nc.why);
    tmm::nccl::UniqueId id;
    int r = nc.GetUniqueId(&id);
    if (r) return tmm::nccl_fail(r, "ncclGetUniqueId");
    memcpy(out128, &id, sizeof id);
    return TMM_OK;
}

int tmm_context_attach_grid(tmm_context* ctx, int grid_rows, int grid_cols, int my_row, int my_col, const void* row_id128, const void* col_id128) {
    if (!ctx) return tmm::fail(TMM_ERR_INVALID, "null context");
    if (grid_rows < 1 || grid_cols < 1 || my_row < 0 || my_row >= grid_rows || my_col < 0 || my_col >= grid_cols)
        return tmm::fail(TMM_ERR_INVALID, "attach_grid: bad grid position (%d,%d) in %dx%d", my_row, my_col, grid_rows, grid_cols);
    if ((grid_cols > 1 && !row_id128) || (grid_rows > 1 && !col_id128)) return tmm::fail(TMM_ERR_INVALID, "attach_grid: missing NCCL unique id");
    if (!ctx->children.empty()) return tmm::fail(TMM_ERR_INVALID, "attach_grid: context already drives several devices");
    tmm::nccl::UniqueId rid{}, cid{};
    if (row_id128) memcpy(&rid, row_id128, sizeof rid);
    if (col_id128) memcpy(&cid, col_id128, sizeof cid);
    return tmm::attach(ctx, grid_rows, grid_cols, my_row, my_col, &rid, &cid);
}

int tmm_context_grid(tmm_context* ctx, int* grid_rows, int* grid_cols, int* my_row, int* my_col) {
    if (!ctx) return tmm::fail(TMM_ERR_INVALID, "null context");
    const tmm::Grid& g = ctx->children.empty() ? ctx->grid : ctx->children[0]->grid;
    if (grid_rows) *grid_rows = g.pr;
    if (grid_cols) *grid_cols = g.pc;
    if (my_row) *my_row = ctx->children.empty() ? g.row : -1;
    if (my_col) *my_col = ctx->children.empty() ? g.col : -1;
    return TMM_OK;
}

int tmm_context_set_devices(tmm_context* ctx, int n_devices, const int* device_ids) {
    if (!ctx) return tmm::fail(TMM_ERR_INVALID, "null context");
    if (n_devices < 1) return tmm::fail(TMM_ERR_INVALID, "set_devices: n_devices must be >= 1");
    int ndev = 0;
    TMM_CU(cudaGetDeviceCount(&ndev));
    std::vector<int> ids(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        ids[i] = device_ids ? device_ids[i] : i;
        if (ids[i] < 0 || ids[i] >= ndev) return tmm::fail(TMM_ERR_INVALID, "set_devices: device %d not present (%d devices)", ids[i], ndev);
        for (int j = 0; j < i; ++j)
            if (ids[j] == ids[i]) return tmm::fail(TMM_ERR_INVALID, "set_devices: device %d listed twice", ids[i]);
    }
    for (tmm_context* ch : ctx->children) tmm_context_destroy(ch);
    ctx->children.clear();
    if (ctx->solo) { tmm_context_destroy(ctx->solo); ctx->solo = nullptr; }
    if (n_devices == 1 && ids[0] == ctx->device) return TMM_OK;  // back to the plain single-GPU context
    int pr, pc;
    tmm_grid_shape(n_devices, &pr, &pc);
    const tmm::nccl::Api& nc = tmm::nccl::api();
    if (n_devices > 1 && !nc.ok) return tmm::fail(TMM_ERR_CUDA, "GPU ERROR: NCCL unavailable: %s", nc.why);
    std::vector<tmm::nccl::UniqueId> row_ids(pr), col_ids(pc);
    if (n_devices > 1) {
        for (auto& id : row_ids) { int r = nc.GetUniqueId(&id); if (r) return tmm::nccl_fail(r, "ncclGetUniqueId"); }
        for (auto& id : col_ids) { int r = nc.GetUniqueId(&id); if (r) return tmm::nccl_fail(r, "ncclGetUniqueId"); }
    }
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = TMM_OK;
    tmm::InternalContextScope internal;  // the contexts created below are the library's own
    for (int i = 0; i < n_devices && !rc; ++i) {
        cudaSetDevice(ids[i]);
        tmm_context* ch = nullptr;
        rc = tmm_context_create(ctx->dtype, ctx->n_streams, ctx->max_tile_m, ctx->max_tile_n, ctx->max_tile_k, &ch);
        if (!rc) { ch->budget_override = ctx->budget_override; ch->profiling = ctx->profiling; ctx->children.push_back(ch); }
    }
    if (!rc && n_devices > 1) {
        // ncclCommInitRank blocks until every rank of the communicator has called it: one thread per device
        std::vector<int> rcs(n_devices, TMM_OK);
        std::vector<std::string> errs(n_devices);
        std::vector<std::thread> threads;
        for (int i = 0; i < n_devices; ++i)
            threads.emplace_back([&, i] {
                cudaSetDevice(ids[i]);
                const int row = i / pc, col = i % pc;
                rcs[i] = tmm::attach(ctx->children[i], pr, pc, row, col, &row_ids[row], &col_ids[col]);
                if (rcs[i]) errs[i] = tmm::last_error_cstr();
            });
        for (auto& t : threads) t.join();
        for (int i = 0; i < n_devices; ++i)
            if (rcs[i] && !rc) rc = tmm::fail(rcs[i], "device %d: %s", ids[i], errs[i].c_str());
        if (!rc) {  // shapes too small for the grid fall back to one plain context on the first device
            cudaSetDevice(ids[0]);
            rc = tmm_context_create(ctx->dtype, ctx->n_streams, ctx->max_tile_m, ctx->max_tile_n, ctx->max_tile_k, &ctx->solo);
            if (!rc && ctx->solo) { ctx->solo->budget_override = ctx->budget_override; ctx->solo->profiling = ctx->profiling; }
        }
    }
    cudaSetDevice(prev);
    if (rc) {
        for (tmm_context* ch : ctx->children) tmm_context_destroy(ch);
        ctx->children.clear();
    }
    return rc;
}

int tmm_context_num_devices(tmm_context* ctx) { return ctx ? (ctx->children.empty() ? 1 : (int)ctx->children.size()) : TMM_ERR_INVALID; }

tmm_context* tmm_context_child(tmm_context* ctx, int index) {
    if (!ctx || index < 0 || index >= (int)ctx->children.size()) return nullptr;
    return ctx->children[index];
}

int tmm_memcpy_2d_async(void* dst, size_t dpitch_bytes, const void* src, size_t spitch_bytes, size_t width_bytes, size_t height, int kind, void* stream) {
    // kind: 1 host->device, 2 device->host, 3 device->device (cudaMemcpyKind values); the copy_tile_* helpers of the
    // reference (tiled_mm.cpp:45-123) are cudaMemcpy2DAsync with exactly these arguments
    if (kind < 1 || kind > 3) return tmm::fail(TMM_ERR_INVALID, "memcpy_2d: bad kind %d", kind);
    if (width_bytes == 0 || height == 0) return TMM_OK;
    TMM_CU(cudaMemcpy2DAsync(dst, dpitch_bytes, src, spitch_bytes, width_bytes, height, (cudaMemcpyKind)kind, (cudaStream_t)stream));
    return TMM_OK;
}

}  // extern "C"
