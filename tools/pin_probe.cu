// How fast can hundreds of GB of host memory be made DMA-able on this box?  (The out-of-core configs need 240 - 326 GB of pinned host
// memory; at the 1.5 GB/s that three concurrent cudaHostAlloc calls reached, that alone is minutes.)  Measurement tool only.
//   build/pin_probe [GiB per trial = 4]
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const size_t gib = argc > 1 ? (size_t)atoi(argv[1]) : 4;
    const size_t bytes = gib << 30;
    cudaFree(0);
    {
        double t0 = now();
        void* p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, bytes, 0);
        printf("cudaHostAlloc x1 thread: %zu GiB in %.2f s = %.2f GB/s (%s)\n", gib, now() - t0, bytes / (now() - t0) * 1e-9, cudaGetErrorString(e));
        if (p) cudaFreeHost(p);
    }
    for (int nt : {4, 16}) {
        double t0 = now();
        std::vector<void*> ps(nt, nullptr);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { cudaSetDevice(0); cudaHostAlloc(&ps[t], bytes / nt, 0); });
        for (auto& x : th) x.join();
        printf("cudaHostAlloc x%d threads (%zu GiB in all): %.2f s = %.2f GB/s\n", nt, gib, now() - t0, bytes / (now() - t0) * 1e-9);
        for (void* p : ps) if (p) cudaFreeHost(p);
    }
    for (int huge : {0, 1})
        for (int nreg : {1, 16}) {
            double t0 = now();
            char* p = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) { printf("mmap failed\n"); continue; }
            if (huge) madvise(p, bytes, MADV_HUGEPAGE);
            const int nt = 32;
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { memset(p + bytes / nt * t, 1, bytes / nt); });
            for (auto& x : th) x.join();
            const double t_touch = now() - t0;
            double t1 = now();
            std::vector<cudaError_t> res(nreg, cudaSuccess);
            std::vector<std::thread> rt;
            for (int r = 0; r < nreg; ++r)
                rt.emplace_back([&, r] { cudaSetDevice(0); res[r] = cudaHostRegister(p + bytes / nreg * r, bytes / nreg, cudaHostRegisterPortable); });
            for (auto& x : rt) x.join();
            const double t_reg = now() - t1;
            printf("mmap%s + first touch (32 threads) %.2f s + cudaHostRegister in %d piece(s) %.2f s (%s): %.2f GB/s overall\n", huge ? " + MADV_HUGEPAGE" : "", t_touch, nreg, t_reg,
                   cudaGetErrorString(res[0]), bytes / (now() - t0) * 1e-9);
            // a copy across the whole range (spans the adjacent registrations) must run at pinned speed
            void* d = nullptr;
            cudaMalloc(&d, (size_t)1 << 30);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const size_t off = bytes / 2 - ((size_t)512 << 20);  // straddles the middle boundary
            cudaMemcpy(d, p + off, (size_t)1 << 30, cudaMemcpyHostToDevice);
            cudaEventRecord(e0); cudaMemcpyAsync(d, p + off, (size_t)1 << 30, cudaMemcpyHostToDevice, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            printf("    1 GiB H2D across the middle of the range: %.1f GB/s\n", (double)((size_t)1 << 30) / ms * 1e-6);
            cudaFree(d);
            for (int r = 0; r < nreg; ++r) cudaHostUnregister(p + bytes / nreg * r);
            munmap(p, bytes);
        }
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) { char buf[128] = {0}; if (fgets(buf, sizeof buf, f)) printf("THP: %s", buf); fclose(f); }
    return 0;
}
