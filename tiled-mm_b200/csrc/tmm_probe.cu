// Box probes behind the C ABI: the two roofline denominators of this path that MEASURED_PEAKS.json does not carry (FP64 tensor issue
// rate, host-link bandwidth per GPU with several GPUs copying at once).  bench.py reports against what these return on the box it runs
// on; a GPU grid uses the host-link figures to size its schedule (tmm_dist.cu) and to pick the GPUs with the best links.
#include "tmm_internal.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace tmm {

LinkRates probe_host_link(size_t bytes, const std::function<void()>& go) {
    LinkRates r;
    char *h_up = nullptr, *h_down = nullptr, *d_up = nullptr, *d_down = nullptr;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ok = cudaHostAlloc((void**)&h_up, bytes, cudaHostAllocDefault) == cudaSuccess && cudaHostAlloc((void**)&h_down, bytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc((void**)&d_up, bytes) == cudaSuccess && cudaMalloc((void**)&d_down, bytes) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& ev : e) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    if (ok) memset(h_up, 0, bytes);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && go) go();  // (called even after a local failure: the other callers wait for this one)
        if (!ok) continue;
        ok = cudaEventRecord(e[0], s_up) == cudaSuccess && cudaEventRecord(e[2], s_down) == cudaSuccess &&
             cudaMemcpyAsync(d_up, h_up, bytes, cudaMemcpyHostToDevice, s_up) == cudaSuccess &&
             cudaMemcpyAsync(h_down, d_down, bytes, cudaMemcpyDeviceToHost, s_down) == cudaSuccess && cudaEventRecord(e[1], s_up) == cudaSuccess &&
             cudaEventRecord(e[3], s_down) == cudaSuccess && cudaStreamSynchronize(s_up) == cudaSuccess && cudaStreamSynchronize(s_down) == cudaSuccess;
    }
    float up_ms = 0, down_ms = 0;
    if (ok && cudaEventElapsedTime(&up_ms, e[0], e[1]) == cudaSuccess && cudaEventElapsedTime(&down_ms, e[2], e[3]) == cudaSuccess && up_ms > 0 && down_ms > 0) {
        r.h2d = (double)bytes / up_ms * 1e-6;
        r.d2h = (double)bytes / down_ms * 1e-6;
    }
    cudaGetLastError();
    for (auto& ev : e) if (ev) cudaEventDestroy(ev);
    if (s_up) cudaStreamDestroy(s_up);
    if (s_down) cudaStreamDestroy(s_down);
    if (d_up) cudaFree(d_up);
    if (d_down) cudaFree(d_down);
    if (h_up) cudaFreeHost(h_up);
    if (h_down) cudaFreeHost(h_down);
    return r;
}

#ifndef TMM_EMULATED
namespace {
// FP64 tensor issue rate: every warp keeps 8 independent DMMA.8x8x4 accumulator chains in flight (mma.sync.m8n8k4.f64, the instruction
// of gemm_f64.cu / gemm_c64.cu); 8 warps x 4 CTAs per SM.  No memory traffic: this is the ceiling a DGEMM kernel can approach.
__global__ void __launch_bounds__(256) dmma_issue_kernel(double* out, int iters, double a0, double b0) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
}  // namespace
#endif

}  // namespace tmm

extern "C" {

int tmm_probe_fp64_peak(double* tflops) {
    if (!tflops) return tmm::fail(TMM_ERR_INVALID, "out is null");
    *tflops = 0;
#ifdef TMM_EMULATED
    return tmm::fail(TMM_ERR_NOGPU, "no device");
#else
    double* out = nullptr;
    TMM_CU(cudaMalloc((void**)&out, 64));
    cudaEvent_t e0, e1;
    TMM_CU(cudaEventCreate(&e0)); TMM_CU(cudaEventCreate(&e1));
    const int sms = tmm::sm_count(), iters = 20000, blocks = sms * 4;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms the clocks up
        TMM_CU(cudaEventRecord(e0, 0));
        tmm::dmma_issue_kernel<<<blocks, 256>>>(out, iters, 1.0, 1e-3);
        TMM_CU(cudaEventRecord(e1, 0));
        TMM_CU(cudaEventSynchronize(e1));
        float ms = 0;
        TMM_CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    TMM_CU(cudaGetLastError());
    // one m8n8k4 MMA = 8 * 8 * 4 FMAs = 512 flop per warp instruction
    const double flop = (double)blocks * 8.0 /* warps */ * (double)iters * 8.0 /* chains */ * 512.0;
    *tflops = flop / ((double)best * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return TMM_OK;
#endif
}

int tmm_probe_host_links(int n_devices, const int* device_ids, size_t bytes, double* h2d_gbs, double* d2h_gbs) {
    if (n_devices < 1 || !h2d_gbs || !d2h_gbs) return tmm::fail(TMM_ERR_INVALID, "probe_host_links: bad argument");
    if (bytes == 0) bytes = (size_t)64 << 20;
    int prev = 0;
    cudaGetDevice(&prev);
    std::atomic<int> arrived{0};
    std::vector<std::thread> pool;
    for (int i = 0; i < n_devices; ++i)
        pool.emplace_back([&, i] {
            cudaSetDevice(device_ids ? device_ids[i] : i);
            const tmm::LinkRates r = tmm::probe_host_link(bytes, [&] {
                arrived.fetch_add(1);
                while (arrived.load() < n_devices) std::this_thread::yield();
            });
            h2d_gbs[i] = r.h2d; d2h_gbs[i] = r.d2h;
        });
    for (auto& t : pool) t.join();
    cudaSetDevice(prev);
    return TMM_OK;
}

}  // extern "C"
