// Development harness (GPU box): device GEMM vs cuBLAS (correctness + speed), then host-to-host tmm_gemm.
// Not part of the product; cuBLAS appears here only as the comparator.
#include "../include/tiled_mm_b200.h"
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static cublasOperation_t op(char t) { return t == 'N' ? CUBLAS_OP_N : (t == 'T' ? CUBLAS_OP_T : CUBLAS_OP_C); }

static void fill(std::vector<double>& v, unsigned seed) {
    unsigned s = seed * 2654435761u + 12345u;
    for (auto& x : v) { s = s * 1664525u + 1013904223u; x = ((double)(s >> 8) / (1 << 24)) * 2.0 - 1.0; }
}

static double check_dev(cublasHandle_t h, char ta, char tb, int m, int n, int k, double alpha, double beta, int pad) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = (ar + pad + 1) & ~1, ldb = (br + pad + 1) & ~1, ldc = m + pad;
    std::vector<double> A((size_t)lda * ac), B((size_t)ldb * bc), C((size_t)ldc * n);
    fill(A, 1); fill(B, 2); fill(C, 3);
    double *dA, *dB, *dC, *dR;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dR, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dR, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    cublasDgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dR, ldc);
    int rc = tmm_device_gemm(TMM_F64, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr);
    CK(cudaDeviceSynchronize());
    if (rc) { printf("tmm_device_gemm rc=%d %s\n", rc, tmm_last_error()); return 1e30; }
    std::vector<double> R(C.size()), O(C.size());
    CK(cudaMemcpy(R.data(), dR, C.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(O.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0; size_t bad_pad = 0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < ldc; ++i) {
        size_t idx = (size_t)j * ldc + i;
        if (i < m) err = std::max(err, std::fabs(R[idx] - O[idx])); else if (O[idx] != C[idx]) ++bad_pad;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
    printf("dev dgemm %c%c m=%d n=%d k=%d a=%.1f b=%.1f pad=%d: max|diff|=%.3e rel=%.2e pad-clobber=%zu %s\n", ta, tb, m, n, k, alpha, beta, pad, err,
           err / (k ? k : 1), bad_pad, (err / (k ? k : 1) < 1e-15 && !bad_pad) ? "OK" : "FAIL");
    return err;
}

static void bench_dev(cublasHandle_t h, char ta, char tb, int m, int n, int k, double beta) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = (ar + 15) & ~15, ldb = (br + 15) & ~15, ldc = m;
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, (size_t)lda * ac * 8)); CK(cudaMalloc(&dB, (size_t)ldb * bc * 8)); CK(cudaMalloc(&dC, (size_t)ldc * n * 8));
    std::vector<double> hb((size_t)1 << 22); fill(hb, 7);
    for (double* p : {dA, dB, dC}) {
        size_t tot = (p == dA ? (size_t)lda * ac : p == dB ? (size_t)ldb * bc : (size_t)ldc * n) * 8;
        for (size_t off = 0; off < tot; off += hb.size() * 8) CK(cudaMemcpy((char*)p + off, hb.data(), std::min(hb.size() * 8, tot - off), cudaMemcpyHostToDevice));
    }
    double alpha = 1.0;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_c = 1e30f, best_t = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0)); cublasDgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best_c = std::min(best_c, ms);
        CK(cudaEventRecord(e0)); int rc = tmm_device_gemm(TMM_F64, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        if (rc) { printf("rc=%d %s\n", rc, tmm_last_error()); break; }
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best_t = std::min(best_t, ms);
    }
    double fl = 2.0 * m * (double)n * k;
    printf("bench %c%c %6d %6d %6d beta=%.0f: cublas %.3f ms %.2f TF | tmm %.3f ms %.2f TF | ratio %.3f\n", ta, tb, m, n, k, beta, best_c, fl / best_c * 1e-9,
           best_t, fl / best_t * 1e-9, best_c / best_t);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

static void host_gemm(cublasHandle_t h, char ta, char tb, int m, int n, int k, double beta, int reps, int copy_back, int streams = 2) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    size_t na = (size_t)ar * ac, nb = (size_t)br * bc, nc = (size_t)m * n;
    double *A, *B, *C, *C0;
    tmm_malloc_pinned(na * 8, (void**)&A); tmm_malloc_pinned(nb * 8, (void**)&B); tmm_malloc_pinned(nc * 8, (void**)&C); tmm_malloc_pinned(nc * 8, (void**)&C0);
    { std::vector<double> t(1 << 20); fill(t, 11); for (size_t i = 0; i < na; ++i) A[i] = t[i & (t.size() - 1)] ; for (size_t i = 0; i < nb; ++i) B[i] = t[(i * 7 + 3) & (t.size() - 1)]; for (size_t i = 0; i < nc; ++i) C0[i] = t[(i * 13 + 5) & (t.size() - 1)]; }
    tmm_context* ctx; int rc = tmm_context_create(TMM_F64, streams, 5000, 5000, 5000, &ctx);
    if (rc) { printf("ctx create failed %s\n", tmm_last_error()); return; }
    double alpha = 1.0, best = 1e30;
    tmm_call_stats st;
    for (int r = 0; r < reps + 1; ++r) {
        memcpy(C, C0, nc * 8);
        auto t0 = std::chrono::steady_clock::now();
        rc = tmm_gemm(ctx, ta, tb, m, n, k, &alpha, A, ar, B, br, &beta, C, m, 0, copy_back);
        auto t1 = std::chrono::steady_clock::now();
        if (rc) { printf("tmm_gemm rc=%d %s\n", rc, tmm_last_error()); return; }
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (r) best = std::min(best, ms);
        tmm_context_last_stats(ctx, &st);
        printf("  run %d: %.2f ms  (h2d %.1f MB in %llu copies, d2h %.1f MB, %llu launches, regime %d, blocks %d chunks %d)\n", r, ms, st.h2d_bytes / 1e6,
               (unsigned long long)st.h2d_copies, st.d2h_bytes / 1e6, (unsigned long long)st.kernel_launches, st.regime, st.c_blocks, st.k_chunks);
    }
    double fl = 2.0 * m * (double)n * k;
    printf("HOST gemm %c%c %d %d %d beta=%.0f copy_back=%d: best %.2f ms = %.2f TFLOP/s\n", ta, tb, m, n, k, beta, copy_back, best, fl / best * 1e-9);
    // verify against cuBLAS on the device
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, na * 8)); CK(cudaMalloc(&dB, nb * 8)); CK(cudaMalloc(&dC, nc * 8));
    CK(cudaMemcpy(dA, A, na * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, nb * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dC, C0, nc * 8, cudaMemcpyHostToDevice));
    cublasDgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, ar, dB, br, &beta, dC, m);
    CK(cudaMemcpy(C0, dC, nc * 8, cudaMemcpyDeviceToHost));
    if (!copy_back) CK(cudaMemcpy(C, tmm_context_device_c(ctx), nc * 8, cudaMemcpyDeviceToHost));
    double err = 0; for (size_t i = 0; i < nc; ++i) err = std::max(err, std::fabs(C[i] - C0[i]));
    printf("  vs cuBLAS: max|diff| = %.3e  (/k = %.2e) %s\n", err, err / k, err / k < 1e-15 ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    tmm_context_destroy(ctx);
    tmm_free_pinned(A); tmm_free_pinned(B); tmm_free_pinned(C); tmm_free_pinned(C0);
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "all";
    cublasHandle_t h; cublasCreate(&h);
    if (mode == "all" || mode == "check") {
        const char* ops[] = {"NN", "TN", "NT", "TT"};
        for (auto o : ops) {
            check_dev(h, o[0], o[1], 128, 64, 16, 1.0, 0.0, 0);
            check_dev(h, o[0], o[1], 257, 131, 77, 1.5, 0.0, 3);
            check_dev(h, o[0], o[1], 1000, 1000, 1000, 1.0, 1.0, 0);
            check_dev(h, o[0], o[1], 5, 2, 2, 1.0, -0.5, 1);
            check_dev(h, o[0], o[1], 50, 200, 21, 2.0, 0.0, 7);
        }
    }
    if (mode == "all" || mode == "bench") {
        bench_dev(h, 'N', 'N', 10000, 10000, 10000, 0.0);
        bench_dev(h, 'T', 'N', 10000, 10000, 10000, 0.0);
        bench_dev(h, 'N', 'T', 10000, 10000, 10000, 0.0);
        bench_dev(h, 'T', 'T', 10000, 10000, 10000, 0.0);
        bench_dev(h, 'N', 'N', 10000, 4800, 256, 1.0);
        bench_dev(h, 'N', 'N', 10000, 4800, 512, 1.0);
        bench_dev(h, 'N', 'N', 10000, 4800, 2048, 1.0);
        bench_dev(h, 'N', 'N', 10000, 1664, 10000, 0.0);
        bench_dev(h, 'N', 'N', 10000, 512, 10000, 0.0);
        bench_dev(h, 'N', 'N', 8192, 8192, 8192, 0.0);
    }
    if (mode == "hostone") {  // hostone ta tb m n k beta copy_back streams reps
        host_gemm(h, argv[2][0], argv[3][0], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atof(argv[7]), atoi(argv[9]) > 0 ? atoi(argv[10]) : 3, atoi(argv[8]), atoi(argv[9]));
        return 0;
    }
    if (mode == "benchone") {  // benchone ta tb m n k beta
        bench_dev(h, argv[2][0], argv[3][0], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atof(argv[7]));
        return 0;
    }
    if (mode == "all" || mode == "host") {
        host_gemm(h, 'N', 'N', 1000, 1000, 1000, 1.0, 2, 1);
        host_gemm(h, 'N', 'N', 10000, 10000, 10000, 0.0, 4, 1);
        host_gemm(h, 'N', 'N', 10000, 10000, 10000, 0.0, 3, 0);
        host_gemm(h, 'N', 'N', 10000, 10000, 10000, 1.0, 3, 1);
        host_gemm(h, 'T', 'N', 10000, 10000, 10000, 0.0, 3, 1);
        host_gemm(h, 'N', 'N', 10000, 10000, 10000, 0.0, 3, 1, 4);
    }
    return 0;
}
