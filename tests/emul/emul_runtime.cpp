// TEST INFRASTRUCTURE ONLY - never part of the product.
// CPU emulation of the CUDA runtime entry points the product's SCHEDULER uses (csrc/tmm_context.cu, tmm_dist.cu), so that the
// real scheduler and the real multi-GPU layer - unmodified, compiled as plain C++ - execute in a GPU-less container:
//   * several "devices" (TMM_EMUL_DEVICES), a per-thread current device, per-device memory accounting (TMM_EMUL_MEM_MB)
//   * every stream operation executes synchronously at enqueue time: a legal serialisation of the stream/event DAG, because
//     an operation is only ever enqueued after the operations it waits for
//   * cuStreamWaitValue32 / cuStreamWriteValue32 (the arrival / ack counters of the peer DMA push) spin on / store to memory,
//     so the host threads that drive different devices synchronise exactly where the GPUs would
//   * every copy and every GEMM operand is bounds-checked against the allocation it points into ("device" memory is
//     malloc()ed and NaN-poisoned): an index error in the scheduler is a test failure, not silent corruption
//   * a happens-before race detector runs underneath: every stream carries a vector clock, events / host synchronisation /
//     stream memory operations / blocking collectives add the edges CUDA guarantees, and every access to device memory (2-D copy
//     regions, GEMM operands) is checked against earlier conflicting accesses - a missing cudaStreamWaitEvent in the scheduler,
//     which on hardware would be an intermittent wrong result, is a deterministic test failure here
// What it cannot show: timing, real overlap, CUDA IPC between processes (validated on hardware at 2 and 4 GPUs).
#include <cuda_runtime_api.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <vector>
#include <mutex>
#include <thread>

namespace {

struct Block { size_t bytes; int device; int kind; };  // kind: 0 device, 1 pinned host (cudaHostAlloc), 2 registered host
std::mutex g_mu;
std::map<uintptr_t, Block> g_blocks;
std::map<uintptr_t, int> g_ipc_imports;  // exporting allocation -> open CUDA IPC imports (under g_mu)
thread_local int t_device = 0;
std::atomic<uint64_t> g_h2d{0}, g_d2h{0}, g_d2d{0}, g_violations{0}, g_unpinned_async{0};
char g_first_violation[256] = "";

int device_count() {
    static const int n = [] { const char* v = getenv("TMM_EMUL_DEVICES"); int k = v ? atoi(v) : 1; return k < 1 ? 1 : k; }();
    return n;
}
size_t device_total() {
    static const size_t b = [] { const char* v = getenv("TMM_EMUL_MEM_MB"); return (size_t)(v ? atoll(v) : 2048) << 20; }();
    return b;
}

// Dry run (TMM_EMUL_DRY=1): allocations above 64 KiB are address ranges without backing store, copies and launches are checked
// (bounds, ordering, byte counts) but move no data - the full-size out-of-core configurations (100000^3 on 8 x 180 GB) can then be
// walked through the real scheduler in seconds.  Small allocations (flags, scalars) stay real: the protocol dereferences them.
bool dry_run() {
    static const bool on = [] { const char* v = getenv("TMM_EMUL_DRY"); return v && v[0] == '1'; }();
    return on;
}
constexpr size_t DRY_REAL_LIMIT = 64 << 10;
std::atomic<uintptr_t> g_virtual_next{0x100000000000ull};
// address-only memory: device ranges handed out by the dry-run cudaMalloc, and the window the dry-run test uses for its host "buffers"
// (exact ranges, not a heuristic: sanitizer allocators place real heap memory at similar-looking addresses)
bool is_virtual(const void* p) {
    if (!dry_run()) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return (a >= 0x100000000000ull && a < g_virtual_next.load()) || (a >= 0x200000000000ull && a < 0x500000000000ull);
}

// fault injection: the n-th cudaMalloc / n-th 2-D copy / n-th stream-memory wait from now on fails once (0 = off)
std::atomic<long> g_fail_malloc{0}, g_fail_copy{0};
bool take_fault(std::atomic<long>& counter) {
    long v = counter.load();
    while (v > 0) {
        if (counter.compare_exchange_weak(v, v - 1)) return v == 1;
    }
    return false;
}

void violation(const char* what, const void* p, size_t bytes) {
    if (g_violations.fetch_add(1) == 0) snprintf(g_first_violation, sizeof g_first_violation, "%s: %p + %zu", what, p, bytes);
    fprintf(stderr, "[emul] VIOLATION %s: %p + %zu bytes\n", what, p, bytes);
}

// the tracked block containing [p, p + bytes), or nullptr
const Block* find_block(const void* p, size_t bytes, uintptr_t* base_out = nullptr) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_blocks.upper_bound(a);
    if (it == g_blocks.begin()) return nullptr;
    --it;
    if (a + bytes > it->first + it->second.bytes) return nullptr;
    if (base_out) *base_out = it->first;
    return &it->second;
}

// device-side ranges must lie inside one device allocation; host-side ranges inside a tracked host block when there is one
void check_range(const void* p, size_t span, bool device_side, bool async_host) {
    std::lock_guard<std::mutex> lk(g_mu);
    const Block* b = find_block(p, span);
    if (device_side) {
        if (!b || b->kind != 0) violation("device range outside any device allocation", p, span);
    } else {
        const Block* head = find_block(p, 1);
        if (head && !b) {
            // a range may span several registrations that touch (the product registers a large buffer in 2 MiB-aligned pieces)
            uintptr_t at = reinterpret_cast<uintptr_t>(p);
            const uintptr_t end = at + span;
            bool covered = true;
            while (at < end) {
                uintptr_t base = 0;
                const Block* piece = find_block(reinterpret_cast<const void*>(at), 1, &base);
                if (!piece || piece->kind == 0) { covered = false; break; }
                at = base + piece->bytes;
            }
            if (!covered) violation("host range runs past its pinned / registered block", p, span);
        }
        if (async_host && !head) g_unpinned_async.fetch_add(1);
    }
}

size_t span_2d(size_t pitch, size_t width, size_t height) { return height ? (height - 1) * pitch + width : 0; }

// ---- happens-before race detector ----------------------------------------------------------------------------------------
using Clock = std::vector<uint32_t>;
void join(Clock& into, const Clock& from) {
    if (into.size() < from.size()) into.resize(from.size(), 0);
    for (size_t i = 0; i < from.size(); ++i) if (from[i] > into[i]) into[i] = from[i];
}
struct Access { uintptr_t base; size_t pitch, width, height; int stream; uint32_t epoch; bool write; const char* what; };
std::mutex g_det;
std::map<cudaStream_t, int> g_stream_id;      // nullptr (legacy stream) is id 0
std::vector<Clock> g_vc(1);                    // per stream
std::map<cudaEvent_t, Clock> g_event_vc;
std::map<int, Clock> g_host_vc;                // per device: the host thread that drives it
std::map<uintptr_t, std::deque<Access>> g_shadow;  // per device allocation
std::map<uintptr_t, Clock> g_flag_vc;          // per 32-bit word written by a stream memory operation (or a 4-byte D2D copy of one)
std::map<const void*, Clock> g_collective;     // per communicator: joined host clocks of its ranks
std::atomic<uint64_t> g_races{0};
char g_first_race[512] = "";

int sid_locked(cudaStream_t s) {
    if (!s) return 0;
    auto it = g_stream_id.find(s);
    if (it != g_stream_id.end()) return it->second;
    const int id = (int)g_vc.size();
    g_vc.emplace_back();
    g_stream_id[s] = id;
    return id;
}
// an operation enters stream s: it is ordered after everything the enqueuing host thread has synchronised with
int begin_op_locked(cudaStream_t s) {
    const int id = sid_locked(s);
    join(g_vc[id], g_host_vc[t_device]);
    if ((int)g_vc[id].size() <= id) g_vc[id].resize(id + 1, 0);
    ++g_vc[id][id];
    return id;
}
bool rows_overlap(const Access& a, const Access& b) {  // exact test on the 2-D byte sets; iterates the rows of a
    for (size_t r = 0; r < a.height; ++r) {
        const uintptr_t lo = a.base + r * a.pitch, hi = lo + a.width;
        if (hi <= b.base) continue;
        for (size_t rr = lo > b.base ? (lo - b.base) / b.pitch : 0; rr < b.height; ++rr) {
            const uintptr_t blo = b.base + rr * b.pitch;
            if (blo >= hi) break;
            if (blo + b.width > lo) return true;
        }
    }
    return false;
}
bool overlap(const Access& a, const Access& b) {
    const uintptr_t a_end = a.base + span_2d(a.pitch, a.width, a.height), b_end = b.base + span_2d(b.pitch, b.width, b.height);
    if (a_end <= b.base || b_end <= a.base) return false;
    if (a.pitch == b.pitch && a.width <= a.pitch && b.width <= b.pitch) {
        // same pitch (the common case: two sub-blocks of one panel): compare as rectangles on the pitch grid anchored at the lower base
        const uintptr_t org = a.base < b.base ? a.base : b.base;
        const size_t ra = (a.base - org) / a.pitch, ca = (a.base - org) % a.pitch, rb = (b.base - org) / b.pitch, cb = (b.base - org) % b.pitch;
        if (ca + a.width <= a.pitch && cb + b.width <= b.pitch)
            return ra < rb + b.height && rb < ra + a.height && ca < cb + b.width && cb < ca + a.width;
    }
    return a.height <= b.height ? rows_overlap(a, b) : rows_overlap(b, a);
}
void note_locked(int sid, const void* p, size_t pitch, size_t width, size_t height, bool write, const char* what) {
    if (!p || !width || !height) return;
    uintptr_t base = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        const Block* b = find_block(p, 1, &base);
        if (!b || b->bytes <= 4096) return;  // untracked (pageable) host memory, or flag / scalar scratch that is synchronisation state itself
        // tracked host blocks (cudaHostAlloc / cudaHostRegister) are checked too: the H2D read of a block of host C (beta != 0) against
        // the D2H write of the result into the same block
    }
    Access now{reinterpret_cast<uintptr_t>(p), pitch ? pitch : width, width, height, sid, g_vc[sid][sid], write, what};
    std::deque<Access>& log = g_shadow[base];
    const Clock& vc = g_vc[sid];
    for (const Access& prev : log) {
        if (prev.stream == sid || (!prev.write && !write)) continue;
        const uint32_t seen = prev.stream < (int)vc.size() ? vc[prev.stream] : 0;
        if (seen >= prev.epoch || !overlap(prev, now)) continue;
        if (g_races.fetch_add(1) == 0)
            snprintf(g_first_race, sizeof g_first_race, "%s %s on stream %d (op %u) is not ordered after %s %s on stream %d (op %u): %p [%zu x %zu, pitch %zu]",
                     what, write ? "write" : "read", sid, now.epoch, prev.what, prev.write ? "write" : "read", prev.stream, prev.epoch, p, width, height, now.pitch);
        fprintf(stderr, "[emul] RACE: %s %s (stream %d) vs earlier %s %s (stream %d) at %p\n", what, write ? "write" : "read", sid, prev.what,
                prev.write ? "write" : "read", prev.stream, p);
        break;
    }
    log.push_back(now);
    if (log.size() > 4096) log.pop_front();
}

}  // namespace

extern "C" {

#define EMUL_API __attribute__((visibility("default")))

// ---- counters the tests read -------------------------------------------------------------------------------------------
EMUL_API void emul_reset_counters() { g_h2d = 0; g_d2h = 0; g_d2d = 0; g_unpinned_async = 0; }
EMUL_API uint64_t emul_h2d_bytes() { return g_h2d; }
EMUL_API uint64_t emul_d2h_bytes() { return g_d2h; }
EMUL_API uint64_t emul_d2d_bytes() { return g_d2d; }
EMUL_API uint64_t emul_violations() { return g_violations; }
EMUL_API uint64_t emul_unpinned_async_copies() { return g_unpinned_async; }
EMUL_API const char* emul_first_violation() { return g_first_violation; }
EMUL_API uint64_t emul_live_device_bytes(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t s = 0;
    for (auto& kv : g_blocks) if (kv.second.kind == 0 && kv.second.device == device) s += kv.second.bytes;
    return s;
}
EMUL_API uint64_t emul_registered_host_blocks() {
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t n = 0;
    for (auto& kv : g_blocks) if (kv.second.kind == 2) ++n;
    return n;
}
EMUL_API void emul_inject_fault(int kind, long nth) { (kind == 0 ? g_fail_malloc : g_fail_copy).store(nth); }
EMUL_API int emul_set_device(int d) { return (int)cudaSetDevice(d); }  // for test threads that play one rank each
EMUL_API uint64_t emul_open_ipc_imports() {
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t n = 0;
    for (auto& kv : g_ipc_imports) n += (uint64_t)kv.second;
    return n;
}
EMUL_API int emul_dry_run() { return dry_run() ? 1 : 0; }
EMUL_API uint64_t emul_races() { return g_races; }
EMUL_API const char* emul_first_race() { return g_first_race; }
// used by the GEMM double: a launch enters `stream`, then declares its operand regions
EMUL_API int emul_op_begin(void* stream) { std::lock_guard<std::mutex> lk(g_det); return begin_op_locked(static_cast<cudaStream_t>(stream)); }
EMUL_API void emul_op_access(int sid, const void* p, size_t pitch, size_t width, size_t height, int write, const char* what) {
    std::lock_guard<std::mutex> lk(g_det);
    note_locked(sid, p, pitch, width, height, write != 0, what);
}
// used by the NCCL stand-in (wired up by the test worker): a blocking collective orders the host threads of its ranks.
// phase 0 before the rendezvous (contribute), phase 1 after it (everybody has contributed: take the join)
EMUL_API void emul_collective(const void* comm_group, int phase) {
    std::lock_guard<std::mutex> lk(g_det);
    if (phase == 0) join(g_collective[comm_group], g_host_vc[t_device]);
    else join(g_host_vc[t_device], g_collective[comm_group]);
}
// a stream-ordered collective (the NCCL staging data plane): the ranks' streams are joined at the rendezvous, and the collective reads
// its send buffer before / writes its receive buffer after it
EMUL_API void emul_collective_stream(const void* comm_group, void* stream, int phase, const void* buf, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_det);
    static std::map<const void*, Clock> acc;
    const int sid = begin_op_locked(static_cast<cudaStream_t>(stream));
    if (phase == 0) { note_locked(sid, buf, bytes, bytes, 1, false, "collective send"); join(acc[comm_group], g_vc[sid]); }
    else { join(g_vc[sid], acc[comm_group]); note_locked(sid, buf, bytes, bytes, 1, true, "collective recv"); }
}
// used by the GEMM double (emul_blas.cpp) to bounds-check operands
EMUL_API void emul_check_device_range(const void* p, size_t bytes) { check_range(p, bytes, true, false); }

// ---- devices -----------------------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = t_device; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= device_count()) return cudaErrorInvalidDevice; t_device = d; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "emulated sm_100");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    p->totalGlobalMem = device_total();
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned) { return (peer >= 0 && peer < device_count()) ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaDeviceGetPCIBusId(char* id, int len, int dev) { snprintf(id, (size_t)len, "0000:%02x:00.0", 0x10 + dev); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    case cudaErrorHostMemoryAlreadyRegistered: return "part or all of the requested memory range is already mapped";
    case cudaErrorHostMemoryNotRegistered: return "pointer does not correspond to a registered memory region";
    default: return "emulated CUDA error";
    }
}
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) {
    const uint64_t used = emul_live_device_bytes(t_device);
    *total_b = device_total();
    *free_b = used < device_total() ? device_total() - used : 0;
    return cudaSuccess;
}

// ---- memory ------------------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (take_fault(g_fail_malloc)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    if (emul_live_device_bytes(t_device) + bytes > device_total()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    void* q = nullptr;
    if (dry_run() && bytes > DRY_REAL_LIMIT) {
        q = reinterpret_cast<void*>(g_virtual_next.fetch_add((bytes + 0xFFFFF) & ~uintptr_t(0xFFFFF)));  // address range only
    } else {
        if (posix_memalign(&q, 256, bytes ? bytes : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
        const uint64_t nan64 = 0x7ff8dead7fc0beefULL;  // NaN as double, and as two floats: reads of never-written device memory show up
        for (size_t i = 0; i + 8 <= bytes; i += 8) memcpy(static_cast<char*>(q) + i, &nan64, 8);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_blocks[reinterpret_cast<uintptr_t>(q)] = Block{bytes ? bytes : 1, t_device, 0};
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
        if (it == g_blocks.end() || it->second.kind != 0) { violation("cudaFree of a pointer that is not a device allocation", p, 0); return cudaErrorInvalidValue; }
        if (g_ipc_imports.count(reinterpret_cast<uintptr_t>(p)) && !getenv("TMM_EMUL_ALLOW_FREE_WHILE_IMPORTED"))
            violation("cudaFree of an allocation that a peer still has imported through CUDA IPC (undefined behaviour on hardware)", p, it->second.bytes);
        g_blocks.erase(it);
    }
    { std::lock_guard<std::mutex> lk(g_det); g_shadow.erase(reinterpret_cast<uintptr_t>(p)); }
    if (!is_virtual(p)) free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) {
    void* q = nullptr;
    if (posix_memalign(&q, 4096, bytes ? bytes : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    std::lock_guard<std::mutex> lk(g_mu);
    g_blocks[reinterpret_cast<uintptr_t>(q)] = Block{bytes ? bytes : 1, -1, 1};
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
        if (it == g_blocks.end() || it->second.kind != 1) return cudaErrorInvalidValue;
        g_blocks.erase(it);
    }
    { std::lock_guard<std::mutex> lk(g_det); g_shadow.erase(reinterpret_cast<uintptr_t>(p)); }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (find_block(p, 1)) return cudaErrorHostMemoryAlreadyRegistered;
    g_blocks[reinterpret_cast<uintptr_t>(p)] = Block{bytes, -1, 2};
    return cudaSuccess;
}
cudaError_t cudaHostUnregister(void* p) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
        if (it == g_blocks.end() || it->second.kind != 2) return cudaErrorHostMemoryNotRegistered;
        g_blocks.erase(it);
    }
    std::lock_guard<std::mutex> lk2(g_det);  // never nested inside g_mu: the detector takes them in the other order
    g_shadow.erase(reinterpret_cast<uintptr_t>(p));
    return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* attr, const void* p) {
    memset(attr, 0, sizeof *attr);
    std::lock_guard<std::mutex> lk(g_mu);
    const Block* b = find_block(p, 1);
    attr->type = !b ? cudaMemoryTypeUnregistered : (b->kind == 0 ? cudaMemoryTypeDevice : cudaMemoryTypeHost);
    attr->device = b && b->kind == 0 ? b->device : 0;
    return cudaSuccess;
}
// CUDA IPC, emulated inside one address space (the product takes this branch for same-process peers only with TMM_DIST_FORCE_IPC=1):
// a handle names the exporting allocation; every import is counted, and freeing an allocation that is still imported somewhere is
// reported - on hardware that is undefined behaviour ("cudaFree on an exported region before cudaIpcCloseMemHandle in the importer").
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
    if (it == g_blocks.end() || it->second.kind != 0) return cudaErrorInvalidValue;  // only whole device allocations are exported here
    memset(h, 0, sizeof *h);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    memcpy(h->reserved, &a, sizeof a);
    memcpy(h->reserved + 8, "emul-ipc", 8);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    uintptr_t a = 0;
    memcpy(&a, h.reserved, sizeof a);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.find(a);
    if (memcmp(h.reserved + 8, "emul-ipc", 8) != 0 || it == g_blocks.end() || it->second.kind != 0) { *p = nullptr; return cudaErrorInvalidResourceHandle; }
    ++g_ipc_imports[a];
    *p = reinterpret_cast<void*>(a);
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ipc_imports.find(reinterpret_cast<uintptr_t>(p));
    if (it == g_ipc_imports.end() || it->second <= 0) { violation("cudaIpcCloseMemHandle of something that is not imported", p, 0); return cudaErrorInvalidValue; }
    if (--it->second == 0) g_ipc_imports.erase(it);
    return cudaSuccess;
}

static void account(size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) g_h2d += bytes;
    else if (kind == cudaMemcpyDeviceToHost) g_d2h += bytes;
    else if (kind == cudaMemcpyDeviceToDevice) g_d2d += bytes;
}
static cudaError_t copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, bool async,
                           cudaStream_t stream = nullptr, bool host_blocks = false) {
    if (width == 0 || height == 0) return cudaSuccess;
    if (width > dpitch || width > spitch) return cudaErrorInvalidPitchValue;
    {
        std::lock_guard<std::mutex> lk(g_det);
        const int sid = begin_op_locked(stream);
        note_locked(sid, src, spitch, width, height, false, "copy");
        note_locked(sid, dst, dpitch, width, height, true, "copy");
        if (width == 4 && height == 1 && kind == cudaMemcpyDeviceToDevice) g_flag_vc[reinterpret_cast<uintptr_t>(dst)] = g_vc[sid];  // a counter forwarded to a peer
        if (host_blocks) join(g_host_vc[t_device], g_vc[sid]);
    }
    bool dst_dev = kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice;
    bool src_dev = kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice;
    if (kind == cudaMemcpyDefault) {  // unified addressing: the direction follows from where the pointers live
        std::lock_guard<std::mutex> lk(g_mu);
        const Block* bd = find_block(dst, 1);
        const Block* bs = find_block(src, 1);
        dst_dev = bd && bd->kind == 0;
        src_dev = bs && bs->kind == 0;
        kind = dst_dev ? (src_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice) : (src_dev ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    }
    check_range(dst, span_2d(dpitch, width, height), dst_dev, async);
    check_range(src, span_2d(spitch, width, height), src_dev, async);
    if (!dry_run() || (width * height <= 4096 && !is_virtual(dst) && !is_virtual(src)))  // dry run: only the protocol's own words move
        for (size_t r = 0; r < height; ++r) memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    std::atomic_thread_fence(std::memory_order_seq_cst);
    account(width * height, kind);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) { return copy_2d(dst, bytes, src, bytes, bytes, 1, kind, false, nullptr, true); }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) { return copy_2d(dst, bytes, src, bytes, bytes, 1, kind, false, s); }
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t s) {
    if (take_fault(g_fail_copy)) return cudaErrorLaunchFailure;
    return copy_2d(dst, dpitch, src, spitch, width, height, kind, true, s);
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) {
    check_range(p, bytes, true, false);
    {
        std::lock_guard<std::mutex> lk(g_det);
        const int sid = begin_op_locked(nullptr);
        note_locked(sid, p, bytes, bytes, 1, true, "memset");
        join(g_host_vc[t_device], g_vc[sid]);
    }
    if (!is_virtual(p)) memset(p, v, bytes);
    return cudaSuccess;
}

// ---- streams and events: everything already happened when it was enqueued -----------------------------------------------
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    { std::lock_guard<std::mutex> lk(g_det); g_stream_id.erase(s); }  // the id (and its clock slot) is retired, never reused
    free(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {  // the host now knows everything enqueued on s has happened
    std::lock_guard<std::mutex> lk(g_det);
    join(g_host_vc[t_device], g_vc[sid_locked(s)]);
    return cudaSuccess;
}
cudaError_t cudaStreamQuery(cudaStream_t s) { return cudaStreamSynchronize(s); }  // always "done" here, which is the same knowledge
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    // self-test of the detector (tests/test_scheduler_emulated.py): with TMM_EMUL_DROP_WAITS=1 every event wait is ignored, i.e.
    // the schedule loses its cross-stream dependencies, and the detector must report races
    static const bool drop = [] { const char* v = getenv("TMM_EMUL_DROP_WAITS"); return v && v[0] == '1'; }();
    if (drop) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_det);
    auto it = g_event_vc.find(e);
    if (it != g_event_vc.end()) join(g_vc[sid_locked(s)], it->second);  // an event that was never recorded orders nothing (CUDA semantics)
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    { std::lock_guard<std::mutex> lk(g_det); g_event_vc.erase(e); }
    free(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_det);
    const int id = sid_locked(s);
    join(g_vc[id], g_host_vc[t_device]);
    g_event_vc[e] = g_vc[id];
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- stream memory operations (driver API, handed out through cudaGetDriverEntryPoint) -----------------------------------
static int emul_wait_value32(cudaStream_t stream, unsigned long long addr, unsigned value, unsigned /*flags: GEQ*/) {
    volatile uint32_t* p = reinterpret_cast<volatile uint32_t*>(static_cast<uintptr_t>(addr));
    const auto t0 = std::chrono::steady_clock::now();
    while ((int32_t)(*p - value) < 0) {
        std::this_thread::yield();
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) {
            violation("cuStreamWaitValue32 never satisfied (deadlock in the peer exchange protocol)", (const void*)p, value);
            return 999;
        }
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    {   // acquire: later work on this stream is ordered after whatever raised the counter
        std::lock_guard<std::mutex> lk(g_det);
        const int sid = begin_op_locked(stream);
        auto it = g_flag_vc.find(static_cast<uintptr_t>(addr));
        if (it != g_flag_vc.end()) join(g_vc[sid], it->second);
    }
    return 0;
}
static int emul_write_value32(cudaStream_t stream, unsigned long long addr, unsigned value, unsigned) {
    {   // release: the counter carries everything enqueued on this stream so far
        std::lock_guard<std::mutex> lk(g_det);
        const int sid = begin_op_locked(stream);
        g_flag_vc[static_cast<uintptr_t>(addr)] = g_vc[sid];
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    *reinterpret_cast<volatile uint32_t*>(static_cast<uintptr_t>(addr)) = value;
    return 0;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    *fn = nullptr;
    if (!strcmp(symbol, "cuStreamWaitValue32")) *fn = reinterpret_cast<void*>(&emul_wait_value32);
    else if (!strcmp(symbol, "cuStreamWriteValue32")) *fn = reinterpret_cast<void*>(&emul_write_value32);
    if (q) *q = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}

}  // extern "C"
