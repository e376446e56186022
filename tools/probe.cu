// Box probe: the two roofline denominators this path needs and MEASURED_PEAKS.json lacks.
//   (1) FP64 peak: raw DMMA / DFMA issue microbenchmarks and device-resident cublasDgemm/Zgemm
//   (2) PCIe: pinned H2D, D2H, duplex; 2-D (pitched) H2D as the tile copies use it
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/probe tools/probe.cu -lcublas
// This is a measurement tool, not part of the product path.
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <thread>
#include <chrono>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

static float time_kernel(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(arg); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(arg); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    return best;
}

struct MB { int blocks, threads, iters; double* out; int variant; };

template <int NACC> static void run_dmma(int nsm, double* out) {
    for (int warps : {4, 8, 16, 32}) {
        int iters = 20000 / NACC * 8 / std::max(1, warps / 4);
        if (iters < 16) iters = 16;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_dmma<NACC><<<nsm, warps * 32>>>(out, iters, 1.0, 1.0); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0)); k_dmma<NACC><<<nsm, warps * 32>>>(out, iters, 1.0, 1.0); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
        }
        double flops = 2.0 * 256 * NACC * (double)iters * warps * nsm;
        printf("DMMA.884 nacc=%2d warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", NACC, warps, best, flops / best * 1e-9);
    }
}
template <int NACC> static void run_dfma(int nsm, double* out) {
    for (int warps : {8, 16, 32}) {
        int iters = 4096;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_dfma<NACC><<<nsm, warps * 32>>>(out, iters, 1.0, 1.0); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0)); k_dfma<NACC><<<nsm, warps * 32>>>(out, iters, 1.0, 1.0); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
        }
        double flops = 2.0 * 32 * NACC * (double)iters * warps * nsm;
        printf("DFMA     nacc=%2d warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", NACC, warps, best, flops / best * 1e-9);
    }
}

static void cublas_bench(cublasHandle_t h, char ta, char tb, int m, int n, int k, bool cplx, int reps) {
    size_t es = cplx ? 16 : 8;
    void *A, *B, *C;
    CK(cudaMalloc(&A, es * (size_t)m * k)); CK(cudaMalloc(&B, es * (size_t)k * n)); CK(cudaMalloc(&C, es * (size_t)m * n));
    CK(cudaMemset(A, 0, es * (size_t)m * k)); CK(cudaMemset(B, 0, es * (size_t)k * n)); CK(cudaMemset(C, 0, es * (size_t)m * n));
    // fill with something non-trivial (power depends on data toggling)
    {
        std::vector<double> hbuf((size_t)1 << 22);
        for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (double)rand() / RAND_MAX - 0.5;
        for (void* p : {A, B}) {
            size_t tot = es * (size_t)m * k; if (p == B) tot = es * (size_t)k * n;
            for (size_t off = 0; off < tot; off += hbuf.size() * 8)
                CK(cudaMemcpy((char*)p + off, hbuf.data(), std::min(hbuf.size() * 8, tot - off), cudaMemcpyHostToDevice));
        }
    }
    auto op = [](char t) { return t == 'N' ? CUBLAS_OP_N : (t == 'T' ? CUBLAS_OP_T : CUBLAS_OP_C); };
    int lda = ta == 'N' ? m : k, ldb = tb == 'N' ? k : n;
    double alpha = 1.0, beta = 0.0; cuDoubleComplex za = {1, 0}, zb = {0, 0};
    auto run = [&]() {
        if (!cplx) cublasDgemm(h, op(ta), op(tb), m, n, k, &alpha, (double*)A, lda, (double*)B, ldb, &beta, (double*)C, m);
        else cublasZgemm(h, op(ta), op(tb), m, n, k, &za, (cuDoubleComplex*)A, lda, (cuDoubleComplex*)B, ldb, &zb, (cuDoubleComplex*)C, m);
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    run(); CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); run(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms); tot += ms;
    }
    double flops = (cplx ? 8.0 : 2.0) * m * (double)n * k;
    printf("cublas%cgemm %c%c %6d %6d %6d  best %.3f ms %.2f TF  avg %.3f ms %.2f TF\n", cplx ? 'Z' : 'D', ta, tb, m, n, k,
           best, flops / best * 1e-9, tot / reps, flops / (tot / reps) * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
}

static void pcie_bench() {
    size_t bytes = (size_t)1 << 30;
    void *h0, *h1, *d0, *d1;
    CK(cudaHostAlloc(&h0, bytes, 0)); CK(cudaHostAlloc(&h1, bytes, 0));
    memset(h0, 1, bytes); memset(h1, 2, bytes);
    CK(cudaMalloc(&d0, bytes)); CK(cudaMalloc(&d1, bytes));
    cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    for (int rep = 0; rep < 3; ++rep) {
        float ms;
        CK(cudaEventRecord(e0, s0)); CK(cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0)); CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("H2D 1GiB pinned      %.2f GB/s\n", bytes / ms * 1e-6);
        CK(cudaEventRecord(e0, s0)); CK(cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s0)); CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("D2H 1GiB pinned      %.2f GB/s\n", bytes / ms * 1e-6);
        // duplex
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s0)); CK(cudaEventRecord(f0, s1));
        CK(cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s1));
        CK(cudaEventRecord(e1, s0)); CK(cudaEventRecord(f1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventSynchronize(f1));
        float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, f0, f1));
        printf("duplex: H2D %.2f GB/s  D2H %.2f GB/s\n", bytes / a * 1e-6, bytes / b * 1e-6);
        // two H2D streams at once (half each)
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s0));
        CK(cudaMemcpyAsync(d0, h0, bytes / 2, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(d1, h1, bytes / 2, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(f1, s1)); CK(cudaStreamWaitEvent(s0, f1, 0)); CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("2xH2D concurrent 2x512MiB  %.2f GB/s aggregate\n", bytes / ms * 1e-6);
    }
    // pitched 2-D H2D: 10000 x 10000 doubles, sub-block 10000 rows x 1000 cols out of ld=10000 -> contiguous really; use 5000-row sub-block
    {
        size_t ld = 10000, rows = 5000, cols = 10000; // 400 MB, host pitch 80000 B, width 40000 B
        float ms;
        for (size_t dpitch : {rows * 8, (size_t)40064}) {
            CK(cudaEventRecord(e0, s0));
            CK(cudaMemcpy2DAsync(d0, dpitch, h0, ld * 8, rows * 8, cols, cudaMemcpyHostToDevice, s0));
            CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("H2D 2D width 40000B spitch 80000B dpitch %zu: %.2f GB/s\n", dpitch, rows * 8 * cols / ms * 1e-6);
        }
        // narrow rows: width 4096 B (512 doubles), 100000 rows... k-panel of a row-major-ish access
        size_t w = 512 * 8, h = 100000;
        CK(cudaEventRecord(e0, s0));
        CK(cudaMemcpy2DAsync(d0, w, h0, 10000 * 8, w, h, cudaMemcpyHostToDevice, s0));
        CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("H2D 2D width 4096B spitch 80000B x100000 rows: %.2f GB/s\n", w * h / ms * 1e-6);
        w = 128 * 8; h = 400000;
        CK(cudaEventRecord(e0, s0));
        CK(cudaMemcpy2DAsync(d0, w, h0, 2000 * 8, w, h, cudaMemcpyHostToDevice, s0));
        CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("H2D 2D width 1024B spitch 16000B x400000 rows: %.2f GB/s\n", w * h / ms * 1e-6);
        // D2H 2D
        CK(cudaEventRecord(e0, s0));
        CK(cudaMemcpy2DAsync(h1, ld * 8, d1, rows * 8, rows * 8, cols, cudaMemcpyDeviceToHost, s0));
        CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("D2H 2D width 40000B dpitch(host) 80000B: %.2f GB/s\n", rows * 8 * cols / ms * 1e-6);
    }
    // small-copy latency: 1 MB and 64 KB H2D
    for (size_t sz : {(size_t)65536, (size_t)1 << 20, (size_t)16 << 20}) {
        float ms; CK(cudaEventRecord(e0, s0));
        for (int i = 0; i < 20; ++i) CK(cudaMemcpyAsync((char*)d0 + i * sz, (char*)h0 + i * sz, sz, cudaMemcpyHostToDevice, s0));
        CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("H2D 20 x %zu B back-to-back: %.2f GB/s (%.1f us each)\n", sz, 20 * sz / ms * 1e-6, ms * 50);
    }
    // cudaHostRegister cost on 800 MB of pageable memory
    {
        size_t sz = (size_t)800 << 20; void* p = aligned_alloc(4096, sz); memset(p, 1, sz);
        cudaEvent_t a; (void)a;
        auto t0 = std::chrono::steady_clock::now();
        CK(cudaHostRegister(p, sz, cudaHostRegisterDefault));
        auto t1 = std::chrono::steady_clock::now();
        CK(cudaHostUnregister(p));
        auto t2 = std::chrono::steady_clock::now();
        printf("cudaHostRegister 800MiB: %.1f ms, unregister %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
               std::chrono::duration<double, std::milli>(t2 - t1).count());
        // pageable (unpinned) H2D
        float ms; CK(cudaEventRecord(e0, s0)); CK(cudaMemcpyAsync(d0, p, sz, cudaMemcpyHostToDevice, s0)); CK(cudaEventRecord(e1, s0)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("H2D pageable 800MiB  %.2f GB/s\n", sz / ms * 1e-6);
        free(p);
    }
    cudaFreeHost(h0); cudaFreeHost(h1); cudaFree(d0); cudaFree(d1);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int ndev; CK(cudaGetDeviceCount(&ndev));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz smem/blk optin=%zu L2=%d MB ndev=%d hw_threads=%u\n", p.name, p.major, p.minor, p.multiProcessorCount,
           p.clockRate, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, ndev, std::thread::hardware_concurrency());
    size_t fr, to; CK(cudaMemGetInfo(&fr, &to)); printf("mem free %.1f GB total %.1f GB\n", fr * 1e-9, to * 1e-9);
    double* out; CK(cudaMalloc(&out, 64));
    int nsm = p.multiProcessorCount;
    run_dmma<8>(nsm, out); run_dmma<16>(nsm, out); run_dmma<32>(nsm, out);
    run_dfma<16>(nsm, out);
    cublasHandle_t h; cublasCreate(&h);
    cublas_bench(h, 'N', 'N', 5000, 5000, 5000, false, 5);
    cublas_bench(h, 'N', 'N', 8192, 8192, 8192, false, 5);
    cublas_bench(h, 'N', 'N', 10000, 10000, 10000, false, 5);
    cublas_bench(h, 'T', 'N', 10000, 10000, 10000, false, 3);
    cublas_bench(h, 'N', 'T', 10000, 10000, 10000, false, 3);
    cublas_bench(h, 'N', 'N', 10000, 10000, 512, false, 5);
    cublas_bench(h, 'N', 'N', 10000, 10000, 256, false, 5);
    cublas_bench(h, 'N', 'N', 10000, 2048, 10000, false, 5);
    cublas_bench(h, 'N', 'N', 6000, 6000, 6000, true, 3);
    cublas_bench(h, 'C', 'N', 6000, 6000, 6000, true, 3);
    pcie_bench();
    return 0;
}
