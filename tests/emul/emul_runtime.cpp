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
// What it cannot show: timing, real overlap, CUDA IPC between processes (validated on hardware at 2 and 4 GPUs).
#include <cuda_runtime_api.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>

namespace {

struct Block { size_t bytes; int device; int kind; };  // kind: 0 device, 1 pinned host (cudaHostAlloc), 2 registered host
std::mutex g_mu;
std::map<uintptr_t, Block> g_blocks;
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
        if (head && !b) violation("host range runs past its pinned / registered block", p, span);
        if (async_host && !head) g_unpinned_async.fetch_add(1);
    }
}

size_t span_2d(size_t pitch, size_t width, size_t height) { return height ? (height - 1) * pitch + width : 0; }

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
    if (emul_live_device_bytes(t_device) + bytes > device_total()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    void* q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    const uint64_t nan64 = 0x7ff8dead7fc0beefULL;  // NaN as double, and as two floats: reads of never-written device memory show up
    for (size_t i = 0; i + 8 <= bytes; i += 8) memcpy(static_cast<char*>(q) + i, &nan64, 8);
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
        g_blocks.erase(it);
    }
    free(p);
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
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.find(reinterpret_cast<uintptr_t>(p));
    if (it == g_blocks.end() || it->second.kind != 2) return cudaErrorHostMemoryNotRegistered;
    g_blocks.erase(it);
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
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }  // one process: peer access is enough
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

static void account(size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) g_h2d += bytes;
    else if (kind == cudaMemcpyDeviceToHost) g_d2h += bytes;
    else if (kind == cudaMemcpyDeviceToDevice) g_d2d += bytes;
}
static cudaError_t copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, bool async) {
    if (width == 0 || height == 0) return cudaSuccess;
    if (width > dpitch || width > spitch) return cudaErrorInvalidPitchValue;
    const bool dst_dev = kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice;
    const bool src_dev = kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice;
    check_range(dst, span_2d(dpitch, width, height), dst_dev, async);
    check_range(src, span_2d(spitch, width, height), src_dev, async);
    for (size_t r = 0; r < height; ++r) memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    std::atomic_thread_fence(std::memory_order_seq_cst);
    account(width * height, kind);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) { return copy_2d(dst, bytes, src, bytes, bytes, 1, kind, false); }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) { return copy_2d(dst, bytes, src, bytes, bytes, 1, kind, false); }
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t) {
    return copy_2d(dst, dpitch, src, spitch, width, height, kind, true);
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) { check_range(p, bytes, true, false); memset(p, v, bytes); return cudaSuccess; }

// ---- streams and events: everything already happened when it was enqueued -----------------------------------------------
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- stream memory operations (driver API, handed out through cudaGetDriverEntryPoint) -----------------------------------
static int emul_wait_value32(cudaStream_t, unsigned long long addr, unsigned value, unsigned /*flags: GEQ*/) {
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
    return 0;
}
static int emul_write_value32(cudaStream_t, unsigned long long addr, unsigned value, unsigned) {
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
