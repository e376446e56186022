// TEST INFRASTRUCTURE ONLY - never part of the product.
// In-process stand-in for libnccl.so.2 (the handful of entry points csrc/tmm_nccl.h resolves with dlopen): communicators are
// rendezvous groups of host threads, collectives are blocking barriers + memcpy over the emulated "device" memory.  Enough
// to run the product's control plane (agreement all-reduce, handle all-gather) and its NCCL-staging data plane on a CPU.
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Group {
    int nranks = 0, joined = 0, arrived = 0;
    uint64_t generation = 0;
    std::vector<const void*> send;
    std::vector<void*> recv;
    std::mutex mu;
    std::condition_variable cv;
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t gen = generation;
        if (++arrived == nranks) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};
struct Comm { Group* group; int rank; };
std::mutex g_mu;
std::map<std::string, Group*> g_groups;
// optional: the runtime emulation's race detector wants to know that a blocking collective orders the host threads of its ranks
void (*g_sync_hook)(const void* group, int phase) = nullptr;
void hook(const void* g, int phase) { if (g_sync_hook) g_sync_hook(g, phase); }
void (*g_stream_hook)(const void* group, void* stream, int phase, const void* buf, size_t bytes) = nullptr;
uint64_t g_next_id = 1;

size_t type_size(int dt) { return dt <= 1 ? 1 : (dt <= 3 ? 4 : 8); }  // Int8 0, Uint8 1, Int32 2, Uint32 3, Int64 4, Uint64 5, ... Float64 8

template <typename T>
void reduce_into(std::vector<char>& out, const std::vector<const void*>& src, size_t count, int op) {
    T* o = reinterpret_cast<T*>(out.data());
    for (size_t i = 0; i < count; ++i) {
        T acc = static_cast<const T*>(src[0])[i];
        for (size_t r = 1; r < src.size(); ++r) {
            const T v = static_cast<const T*>(src[r])[i];
            acc = op == 0 ? (T)(acc + v) : op == 1 ? (T)(acc * v) : op == 2 ? (v > acc ? v : acc) : (v < acc ? v : acc);
        }
        o[i] = acc;
    }
}
}  // namespace

extern "C" {
#define STUB_API __attribute__((visibility("default")))

STUB_API void nccl_stub_set_sync_hook(void (*fn)(const void*, int)) { g_sync_hook = fn; }
STUB_API void nccl_stub_set_stream_hook(void (*fn)(const void*, void*, int, const void*, size_t)) { g_stream_hook = fn; }
STUB_API int ncclGetVersion(int* v) { *v = 29999; return 0; }
STUB_API const char* ncclGetErrorString(int rc) { return rc == 0 ? "no error" : "emulated NCCL error"; }
STUB_API int ncclGetUniqueId(void* id128) {
    std::lock_guard<std::mutex> lk(g_mu);
    memset(id128, 0, 128);
    const uint64_t v = g_next_id++;
    memcpy(id128, &v, sizeof v);
    memcpy(static_cast<char*>(id128) + 8, "tmm-emul", 8);
    return 0;
}
struct UniqueId { char internal[128]; };
STUB_API int ncclCommInitRank(void** comm, int nranks, UniqueId id, int rank) {
    Group* g;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Group*& slot = g_groups[std::string(id.internal, 128)];
        if (!slot) { slot = new Group(); slot->nranks = nranks; slot->send.resize(nranks); slot->recv.resize(nranks); }
        g = slot;
        if (g->nranks != nranks || rank < 0 || rank >= nranks) return 4;  // ncclInvalidArgument
    }
    *comm = new Comm{g, rank};
    hook(g, 0);
    g->barrier();  // like NCCL: returns once every rank has joined
    hook(g, 1);
    return 0;
}
STUB_API int ncclCommDestroy(void* comm) { delete static_cast<Comm*>(comm); return 0; }
STUB_API int ncclGroupStart() { return 0; }
STUB_API int ncclGroupEnd() { return 0; }

STUB_API int ncclAllGather(const void* send, void* recv, size_t count, int dtype, void* comm, void* stream) {
    Comm* c = static_cast<Comm*>(comm);
    Group* g = c->group;
    const size_t bytes = count * type_size(dtype);
    g->send[c->rank] = send;
    if (g_stream_hook) g_stream_hook(g, stream, 0, send, bytes);
    g->barrier();
    if (g_stream_hook) g_stream_hook(g, stream, 1, recv, bytes * g->nranks);
    for (int r = 0; r < g->nranks; ++r) memcpy(static_cast<char*>(recv) + (size_t)r * bytes, g->send[r], bytes);
    g->barrier();  // nobody reuses its send buffer before everyone has read it
    return 0;
}
STUB_API int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void*) {
    Comm* c = static_cast<Comm*>(comm);
    Group* g = c->group;
    g->send[c->rank] = send;
    hook(g, 0);
    g->barrier();
    hook(g, 1);
    std::vector<char> out(count * type_size(dtype));
    if (dtype == 2) reduce_into<int32_t>(out, g->send, count, op);
    else if (dtype == 4) reduce_into<int64_t>(out, g->send, count, op);
    else if (dtype == 8) reduce_into<double>(out, g->send, count, op);
    else return 4;
    g->barrier();  // everyone has read every send buffer (in-place all-reduce is allowed)
    memcpy(recv, out.data(), out.size());
    g->barrier();
    return 0;
}
}  // extern "C"
