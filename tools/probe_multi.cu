// Box probe for the multi-GPU path: what do the host links and NVLink deliver when SEVERAL GPUs move data at once?
//   (1) pinned H2D / D2H / duplex bandwidth per GPU for growing sets of concurrently active GPUs (the aggregate host-link roofline)
//   (2) peer-to-peer copy-engine pushes: one pair, one-to-all, all-to-all
// One process, one host thread per GPU, copies released together by a spin barrier; timed with CUDA events on each GPU.
// Measurement tool, not part of the product path.   build/probe_multi [MiB per copy = 512]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

struct Dev {
    int id;
    char *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr, *d_peer = nullptr;
    cudaStream_t s_up, s_down, s_p2p[8];
    cudaEvent_t e0, e1, f0, f1;
};

static std::atomic<int> g_arrived{0};
static std::atomic<int> g_phase{0};
static void spin_barrier(int n) {
    const int ph = g_phase.load();
    if (g_arrived.fetch_add(1) + 1 == n) { g_arrived.store(0); g_phase.fetch_add(1); }
    else while (g_phase.load() == ph) {}
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 512) << 20;
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd > 8) nd = 8;
    printf("probe_multi: %d GPUs, %zu MiB per copy\n", nd, bytes >> 20);
    std::vector<Dev> dev(nd);
    {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; ++i)
            th.emplace_back([&, i] {
                Dev& d = dev[i];
                d.id = i;
                CK(cudaSetDevice(i));
                CK(cudaHostAlloc((void**)&d.h_in, bytes, cudaHostAllocPortable));
                CK(cudaHostAlloc((void**)&d.h_out, bytes, cudaHostAllocPortable));
                memset(d.h_in, 1, bytes); memset(d.h_out, 2, bytes);
                CK(cudaMalloc((void**)&d.d_in, bytes)); CK(cudaMalloc((void**)&d.d_out, bytes)); CK(cudaMalloc((void**)&d.d_peer, bytes));
                CK(cudaStreamCreateWithFlags(&d.s_up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&d.s_down, cudaStreamNonBlocking));
                for (auto& s : d.s_p2p) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
                CK(cudaEventCreate(&d.e0)); CK(cudaEventCreate(&d.e1)); CK(cudaEventCreate(&d.f0)); CK(cudaEventCreate(&d.f1));
                for (int j = 0; j < nd; ++j)
                    if (j != i) { cudaError_t e = cudaDeviceEnablePeerAccess(j, 0); if (e != cudaSuccess) { cudaGetLastError(); } }
            });
        for (auto& t : th) t.join();
    }
    // ---- (1) host links ----
    std::vector<std::vector<int>> sets = {{0}};
    if (nd >= 2) sets.push_back({0, 1});
    if (nd >= 4) { sets.push_back({0, 2}); sets.push_back({0, 1, 2, 3}); }
    if (nd >= 8) { sets.push_back({0, 4}); sets.push_back({0, 2, 4, 6}); sets.push_back({4, 5, 6, 7}); sets.push_back({0, 1, 2, 3, 4, 5, 6, 7}); }
    const char* modes[] = {"H2D", "D2H", "duplex"};
    for (auto& set : sets) {
        for (int mode = 0; mode < 3; ++mode) {
            const int n = (int)set.size();
            std::vector<float> up(n, 0.f), down(n, 0.f);
            std::vector<std::thread> th;
            const auto w0 = std::chrono::steady_clock::now();
            for (int t = 0; t < n; ++t)
                th.emplace_back([&, t] {
                    Dev& d = dev[set[t]];
                    CK(cudaSetDevice(d.id));
                    const int reps = 2;
                    for (int warm = 0; warm < 2; ++warm) {  // pass 0 = warm-up, pass 1 = timed
                        spin_barrier(n);
                        if (mode != 1) CK(cudaEventRecord(d.e0, d.s_up));
                        if (mode != 0) CK(cudaEventRecord(d.f0, d.s_down));
                        for (int r = 0; r < reps; ++r) {
                            if (mode != 1) CK(cudaMemcpyAsync(d.d_in, d.h_in, bytes, cudaMemcpyHostToDevice, d.s_up));
                            if (mode != 0) CK(cudaMemcpyAsync(d.h_out, d.d_out, bytes, cudaMemcpyDeviceToHost, d.s_down));
                        }
                        if (mode != 1) CK(cudaEventRecord(d.e1, d.s_up));
                        if (mode != 0) CK(cudaEventRecord(d.f1, d.s_down));
                        CK(cudaStreamSynchronize(d.s_up)); CK(cudaStreamSynchronize(d.s_down));
                    }
                    if (mode != 1) { float ms; CK(cudaEventElapsedTime(&ms, d.e0, d.e1)); up[t] = (float)(2.0 * bytes / ms * 1e-6); }
                    if (mode != 0) { float ms; CK(cudaEventElapsedTime(&ms, d.f0, d.f1)); down[t] = (float)(2.0 * bytes / ms * 1e-6); }
                });
            for (auto& t : th) t.join();
            (void)w0;
            std::string who;
            for (int g : set) who += std::to_string(g);
            float su = 0, sd = 0;
            printf("host-link %-6s GPUs {%s}:", modes[mode], who.c_str());
            for (int t = 0; t < n; ++t) { printf(" %5.1f/%5.1f", up[t], down[t]); su += up[t]; sd += down[t]; }
            printf("   | sum up %6.1f down %6.1f GB/s\n", su, sd);
            fflush(stdout);
        }
    }
    // ---- (2) peer pushes ----
    if (nd >= 2) {
        auto push = [&](const std::vector<std::pair<int, int>>& pairs, const char* label) {
            // all pushes released together; each source uses one stream per destination
            std::vector<int> srcs;
            for (auto& p : pairs) { bool seen = false; for (int s : srcs) seen |= s == p.first; if (!seen) srcs.push_back(p.first); }
            const int n = (int)srcs.size();
            std::vector<float> ms(n, 0.f);
            std::vector<size_t> sent(n, 0);
            std::vector<std::thread> th;
            for (int t = 0; t < n; ++t)
                th.emplace_back([&, t] {
                    Dev& d = dev[srcs[t]];
                    CK(cudaSetDevice(d.id));
                    for (int warm = 0; warm < 2; ++warm) {
                        spin_barrier(n);
                        CK(cudaEventRecord(d.e0, d.s_up));
                        int q = 0;
                        sent[t] = 0;
                        for (auto& p : pairs) {
                            if (p.first != d.id) continue;
                            CK(cudaStreamWaitEvent(d.s_p2p[q], d.e0, 0));
                            CK(cudaMemcpyAsync(dev[p.second].d_peer, d.d_in, bytes, cudaMemcpyDeviceToDevice, d.s_p2p[q]));
                            CK(cudaEventRecord(d.f0, d.s_p2p[q]));
                            CK(cudaStreamWaitEvent(d.s_up, d.f0, 0));
                            sent[t] += bytes;
                            ++q;
                        }
                        CK(cudaEventRecord(d.e1, d.s_up));
                        CK(cudaStreamSynchronize(d.s_up));
                    }
                    CK(cudaEventElapsedTime(&ms[t], d.e0, d.e1));
                });
            for (auto& t : th) t.join();
            printf("peer push %-28s:", label);
            double sum = 0;
            for (int t = 0; t < n; ++t) { const double gbs = sent[t] / ms[t] * 1e-6; printf(" %6.1f", gbs); sum += gbs; }
            printf("   | sum %7.1f GB/s sent\n", sum);
            fflush(stdout);
        };
        push({{0, 1}}, "0->1");
        {
            std::vector<std::pair<int, int>> v;
            for (int j = 1; j < nd && j < 4; ++j) v.push_back({0, j});
            push(v, "0->{1..3} (parallel streams)");
        }
        {
            std::vector<std::pair<int, int>> v;
            for (int i = 0; i < nd; ++i) v.push_back({i, (i + 1) % nd});
            push(v, "ring i->i+1, all GPUs");
        }
        {
            std::vector<std::pair<int, int>> v;
            for (int i = 0; i < nd; ++i) for (int j = 0; j < nd; ++j) if (i != j && i / 4 == j / 4) v.push_back({i, j});
            push(v, "all-to-all within groups of 4");
        }
        // duplex host links WHILE every GPU also pushes to a neighbour (the grid's steady state)
        {
            const int n = nd;
            std::vector<float> up(n), down(n), pp(n);
            std::vector<std::thread> th;
            for (int t = 0; t < n; ++t)
                th.emplace_back([&, t] {
                    Dev& d = dev[t];
                    CK(cudaSetDevice(d.id));
                    for (int warm = 0; warm < 2; ++warm) {
                        spin_barrier(n);
                        CK(cudaEventRecord(d.e0, d.s_up)); CK(cudaEventRecord(d.f0, d.s_down));
                        cudaEvent_t p0, p1; CK(cudaEventCreate(&p0)); CK(cudaEventCreate(&p1));
                        CK(cudaEventRecord(p0, d.s_p2p[0]));
                        for (int r = 0; r < 2; ++r) {
                            CK(cudaMemcpyAsync(d.d_in, d.h_in, bytes, cudaMemcpyHostToDevice, d.s_up));
                            CK(cudaMemcpyAsync(d.h_out, d.d_out, bytes, cudaMemcpyDeviceToHost, d.s_down));
                            CK(cudaMemcpyAsync(dev[(t + 1) % n].d_peer, d.d_out, bytes, cudaMemcpyDeviceToDevice, d.s_p2p[0]));
                        }
                        CK(cudaEventRecord(d.e1, d.s_up)); CK(cudaEventRecord(d.f1, d.s_down)); CK(cudaEventRecord(p1, d.s_p2p[0]));
                        CK(cudaStreamSynchronize(d.s_up)); CK(cudaStreamSynchronize(d.s_down)); CK(cudaStreamSynchronize(d.s_p2p[0]));
                        float ms;
                        CK(cudaEventElapsedTime(&ms, d.e0, d.e1)); up[t] = (float)(2.0 * bytes / ms * 1e-6);
                        CK(cudaEventElapsedTime(&ms, d.f0, d.f1)); down[t] = (float)(2.0 * bytes / ms * 1e-6);
                        CK(cudaEventElapsedTime(&ms, p0, p1)); pp[t] = (float)(2.0 * bytes / ms * 1e-6);
                        CK(cudaEventDestroy(p0)); CK(cudaEventDestroy(p1));
                    }
                });
            for (auto& t : th) t.join();
            printf("duplex + peer push, all GPUs (up/down/peer GB/s):");
            for (int t = 0; t < n; ++t) printf(" %4.0f/%4.0f/%4.0f", up[t], down[t], pp[t]);
            printf("\n");
        }
    }
    return 0;
}
