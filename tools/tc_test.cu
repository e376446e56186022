// Development harness (GPU box) for the tcgen05 SGEMM: correctness against a CPU double reference / cuBLAS SGEMM,
// layout probes (which operand element reached which accumulator position), a sweep over the MN-major descriptor
// fields, and timing against cuBLAS (FP32 and TF32 math).  Not part of the product; cuBLAS is the comparator only.
#include "../include/tiled_mm_b200.h"
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(1);} } while (0)

static cublasOperation_t op(char t) { return t == 'N' ? CUBLAS_OP_N : (t == 'T' ? CUBLAS_OP_T : CUBLAS_OP_C); }

static void fill(std::vector<float>& v, unsigned seed, bool ints) {
    unsigned s = seed * 2654435761u + 12345u;
    for (auto& x : v) {
        s = s * 1664525u + 1013904223u;
        x = ints ? (float)((s >> 16) % 10) : (float)(((double)(s >> 8) / (1 << 24)) * 2.0 - 1.0);
    }
}

static inline float geta(const std::vector<float>& A, char ta, int lda, int i, int kk) { return ta == 'N' ? A[(size_t)kk * lda + i] : A[(size_t)i * lda + kk]; }
static inline float getb(const std::vector<float>& B, char tb, int ldb, int kk, int j) { return tb == 'N' ? B[(size_t)j * ldb + kk] : B[(size_t)kk * ldb + j]; }

// returns normalised error; prints one line
static double check(cublasHandle_t h, char ta, char tb, int m, int n, int k, float alpha, float beta, int pad, bool ints, bool use_cublas_ref, const char* tag) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = (ar + pad + 3) & ~3, ldb = (br + pad + 3) & ~3, ldc = m + pad;
    std::vector<float> A((size_t)lda * ac), B((size_t)ldb * bc), C((size_t)ldc * n);
    fill(A, 1, ints); fill(B, 2, ints); fill(C, 3, ints);
    float *dA, *dB, *dC, *dR;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&dR, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dR, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
    int rc = tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr);
    cudaError_t se = cudaDeviceSynchronize();
    if (rc || se != cudaSuccess) { printf("%s %c%c %d %d %d: rc=%d (%s) sync=%s FAIL\n", tag, ta, tb, m, n, k, rc, tmm_last_error(), cudaGetErrorString(se)); fflush(stdout); exit(2); }
    std::vector<float> O(C.size());
    CK(cudaMemcpy(O.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> R((size_t)m * n);
    if (use_cublas_ref) {
        cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dR, ldc);
        std::vector<float> Rf(C.size());
        CK(cudaMemcpy(Rf.data(), dR, C.size() * 4, cudaMemcpyDeviceToHost));
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) R[(size_t)j * m + i] = Rf[(size_t)j * ldc + i];
    } else {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                double s = 0;
                for (int kk = 0; kk < k; ++kk) s += (double)geta(A, ta, lda, i, kk) * (double)getb(B, tb, ldb, kk, j);
                R[(size_t)j * m + i] = (double)alpha * s + (beta != 0.f ? (double)beta * C[(size_t)j * ldc + i] : 0.0);
            }
    }
    double err = 0; size_t bad_pad = 0, nbad = 0; int fi = -1, fj = -1;
    double amax = ints ? 9 : 1, bmax = ints ? 9 : 1;
    for (int j = 0; j < n; ++j) for (int i = 0; i < ldc; ++i) {
        size_t idx = (size_t)j * ldc + i;
        if (i < m) {
            double d = std::fabs(R[(size_t)j * m + i] - (double)O[idx]);
            if (!(d == d)) d = 1e30;
            if (d > err) { err = d; }
            if (d > 1e-4 * std::max(1, k) * amax * bmax) { if (!nbad) { fi = i; fj = j; } ++nbad; }
        } else if (O[idx] != C[idx]) ++bad_pad;
    }
    double rel = err / (std::max(1, k) * amax * bmax);
    bool ok = rel < (use_cublas_ref ? 2e-6 : 5e-7) && !bad_pad;
    printf("%s %c%c m=%d n=%d k=%d a=%.2f b=%.2f pad=%d %s ref=%s: max|diff|=%.3e rel=%.2e gross=%zu first=(%d,%d) pad-clobber=%zu %s\n", tag, ta, tb, m, n, k, alpha, beta,
           pad, ints ? "ints" : "rand", use_cublas_ref ? "cublas" : "cpu64", err, rel, nbad, fi, fj, bad_pad, ok ? "OK" : "FAIL");
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
    return rel;
}

// Layout probe: one operand carries its own coordinates (value = mn * 64 + kk, exact under 3xTF32), the other is the
// identity in k, so C shows which operand element reached which accumulator position.
static int probe(char ta, char tb, bool probe_a) {
    const int m = 128, n = 128, k = 32;
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = ar, ldb = br, ldc = m;
    std::vector<float> A((size_t)lda * ac, 0.f), B((size_t)ldb * bc, 0.f), C((size_t)ldc * n, 0.f);
    for (int i = 0; i < m; ++i) for (int kk = 0; kk < k; ++kk) {
        float v = probe_a ? (float)(i * 64 + kk) : (i == kk ? 1.f : 0.f);
        if (ta == 'N') A[(size_t)kk * lda + i] = v; else A[(size_t)i * lda + kk] = v;
    }
    for (int j = 0; j < n; ++j) for (int kk = 0; kk < k; ++kk) {
        float v = probe_a ? (j == kk ? 1.f : 0.f) : (float)(j * 64 + kk);
        if (tb == 'N') B[(size_t)j * ldb + kk] = v; else B[(size_t)kk * ldb + j] = v;
    }
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, C.size() * 4));
    float alpha = 1.f, beta = 0.f;
    int rc = tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr);
    cudaError_t se = cudaDeviceSynchronize();
    if (rc || se != cudaSuccess) { printf("probe %c%c: rc=%d sync=%s\n", ta, tb, rc, cudaGetErrorString(se)); fflush(stdout); exit(2); }
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    // expected: probe_a: C[i][j] = i*64 + j for j < k, else 0.   probe_b: C[i][j] = j*64 + i for i < k, else 0
    int wrong = 0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) {
        float e = probe_a ? (j < k ? (float)(i * 64 + j) : 0.f) : (i < k ? (float)(j * 64 + i) : 0.f);
        if (C[(size_t)j * ldc + i] != e) ++wrong;
    }
    printf("probe %s of %c%c: %d wrong of %d\n", probe_a ? "A" : "B", ta, tb, wrong, m * n);
    if (wrong) {
        const int sel[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 64, 127};
        printf("  got (mn,k) at accumulator [row][col] for %s index = row|col below, k index across:\n", probe_a ? "row" : "col");
        for (int s : sel) {
            printf("  %3d:", s);
            for (int q = 0; q < 32; ++q) {
                float v = probe_a ? C[(size_t)q * ldc + s] : C[(size_t)s * ldc + q];
                int iv = (int)v;
                if (v != (float)iv || iv < 0) printf(" (%g)", v); else printf(" %d.%d", iv / 64, iv % 64);
            }
            printf("\n");
        }
    }
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return wrong;
}

// Accuracy against an exact (double) reference with well-mixed inputs: ours at the current TMM_TC_WINDOW vs cuBLAS FP32 (pedantic)
static inline float hrand(uint64_t i, uint64_t seed, bool positive) {
    uint64_t z = (i + seed * 0x9E3779B97F4A7C15ull) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    double u = (double)(z >> 11) / 9007199254740992.0;
    return (float)(positive ? u : 2.0 * u - 1.0);
}
static void precision(cublasHandle_t h, char ta, char tb, int m, int n, int k, bool positive) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = ar, ldb = br, ldc = m;
    std::vector<float> A((size_t)lda * ac), B((size_t)ldb * bc);
    for (size_t i = 0; i < A.size(); ++i) A[i] = hrand(i, 1, positive);
    for (size_t i = 0; i < B.size(); ++i) B[i] = hrand(i, 2, positive);
    std::vector<double> R((size_t)m * n);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double s = 0;
            for (int kk = 0; kk < k; ++kk) s += (double)geta(A, ta, lda, i, kk) * (double)getb(B, tb, ldb, kk, j);
            R[(size_t)j * m + i] = s;
        }
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, (size_t)m * n * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    float alpha = 1.f, beta = 0.f;
    std::vector<float> O((size_t)m * n);
    double cmax = 0; for (double r : R) cmax = std::max(cmax, std::fabs(r));
    auto report = [&](const char* who) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(O.data(), dC, O.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, e2 = 0;
        for (size_t i = 0; i < O.size(); ++i) { double d = std::fabs((double)O[i] - R[i]); emax = std::max(emax, d); e2 += d * d; }
        printf("  %-22s max|err|=%.3e  rms=%.3e  max/(k)=%.2e  max/max|C|=%.2e\n", who, emax, std::sqrt(e2 / O.size()), emax / k, emax / cmax);
    };
    const char* w = getenv("TMM_TC_WINDOW");
    printf("precision %c%c m=%d n=%d k=%d %s window=%s max|C|=%.1f\n", ta, tb, m, n, k, positive ? "uniform(0,1)" : "uniform(-1,1)", w ? w : "default", cmax);
    cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
    cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc); report("cuBLAS fp32 pedantic");
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
    cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc); report("cuBLAS fp32 default");
    tmm_set_f32_math(TMM_MATH_FP32);
    if (tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr)) { printf("rc!=0 %s\n", tmm_last_error()); exit(2); }
    report("tmm fp32 (3xTF32)");
    tmm_set_f32_math(TMM_MATH_SIMT);
    tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr); report("tmm simt ffma");
    tmm_set_f32_math(TMM_MATH_TF32);
    tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr); report("tmm tf32");
    tmm_set_f32_math(TMM_MATH_FP32);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

static void bench(cublasHandle_t h, char ta, char tb, int m, int n, int k, float beta) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = (ar + 31) & ~31, ldb = (br + 31) & ~31, ldc = m;
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, (size_t)lda * ac * 4)); CK(cudaMalloc(&dB, (size_t)ldb * bc * 4)); CK(cudaMalloc(&dC, (size_t)ldc * n * 4));
    std::vector<float> hb((size_t)1 << 22); fill(hb, 7, false);
    for (float* p : {dA, dB, dC}) {
        size_t tot = (p == dA ? (size_t)lda * ac : p == dB ? (size_t)ldb * bc : (size_t)ldc * n) * 4;
        for (size_t off = 0; off < tot; off += hb.size() * 4) CK(cudaMemcpy((char*)p + off, hb.data(), std::min(hb.size() * 4, tot - off), cudaMemcpyHostToDevice));
    }
    float alpha = 1.f;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float t[4] = {1e30f, 1e30f, 1e30f, 1e30f};  // cublas fp32, cublas tf32, tmm fp32(3x), tmm tf32
    for (int r = 0; r < 4; ++r) {
        float ms;
        cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
        CK(cudaEventRecord(e0)); cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) t[0] = std::min(t[0], ms);
        cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
        CK(cudaEventRecord(e0)); cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) t[1] = std::min(t[1], ms);
        cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
        for (int mode = 0; mode < 2; ++mode) {
            tmm_set_f32_math(mode ? TMM_MATH_TF32 : TMM_MATH_FP32);
            CK(cudaEventRecord(e0)); int rc = tmm_device_gemm(TMM_F32, ta, tb, m, n, k, &alpha, dA, lda, dB, ldb, &beta, dC, ldc, nullptr); CK(cudaEventRecord(e1));
            cudaError_t se = cudaEventSynchronize(e1);
            if (rc || se != cudaSuccess) { printf("bench rc=%d sync=%s\n", rc, cudaGetErrorString(se)); fflush(stdout); exit(2); }
            CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) t[2 + mode] = std::min(t[2 + mode], ms);
        }
        tmm_set_f32_math(TMM_MATH_FP32);
    }
    double fl = 2.0 * m * (double)n * k * 1e-9;
    printf("bench %c%c %6d %6d %6d beta=%.0f: cublas fp32 %.3f ms %.1f TF | cublas tf32 %.3f ms %.1f TF | tmm fp32(3xTF32) %.3f ms %.1f TF | tmm tf32 %.3f ms %.1f TF\n", ta, tb, m,
           n, k, beta, t[0], fl / t[0], t[1], fl / t[1], t[2], fl / t[2], t[3], fl / t[3]);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

static void host_gemm(cublasHandle_t h, char ta, char tb, int m, int n, int k, float beta, int reps) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    size_t na = (size_t)ar * ac, nb = (size_t)br * bc, nc = (size_t)m * n;
    float *A, *B, *C, *C0;
    tmm_malloc_pinned(na * 4, (void**)&A); tmm_malloc_pinned(nb * 4, (void**)&B); tmm_malloc_pinned(nc * 4, (void**)&C); tmm_malloc_pinned(nc * 4, (void**)&C0);
    { std::vector<float> t(1 << 20); fill(t, 11, false); for (size_t i = 0; i < na; ++i) A[i] = t[i & (t.size() - 1)]; for (size_t i = 0; i < nb; ++i) B[i] = t[(i * 7 + 3) & (t.size() - 1)]; for (size_t i = 0; i < nc; ++i) C0[i] = t[(i * 13 + 5) & (t.size() - 1)]; }
    tmm_context* ctx; int rc = tmm_context_create(TMM_F32, 2, 5000, 5000, 5000, &ctx);
    if (rc) { printf("ctx create failed %s\n", tmm_last_error()); return; }
    float alpha = 1.f; double best = 1e30;
    tmm_call_stats st;
    for (int r = 0; r < reps + 1; ++r) {
        memcpy(C, C0, nc * 4);
        rc = tmm_gemm(ctx, ta, tb, m, n, k, &alpha, A, ar, B, br, &beta, C, m, 0, 1);
        if (rc) { printf("tmm_gemm rc=%d %s\n", rc, tmm_last_error()); return; }
        tmm_context_last_stats(ctx, &st);
        if (r) best = std::min(best, st.wall_ms);
        printf("  run %d: %.2f ms (h2d %.1f MB, d2h %.1f MB, %llu launches, regime %d)\n", r, st.wall_ms, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6, (unsigned long long)st.kernel_launches, st.regime);
    }
    printf("HOST sgemm %c%c %d %d %d beta=%.0f: best %.2f ms = %.2f TFLOP/s\n", ta, tb, m, n, k, beta, best, 2.0 * m * (double)n * k / best * 1e-9);
    float *dA, *dB, *dC;
    CK(cudaMalloc(&dA, na * 4)); CK(cudaMalloc(&dB, nb * 4)); CK(cudaMalloc(&dC, nc * 4));
    CK(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, nb * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dC, C0, nc * 4, cudaMemcpyHostToDevice));
    cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
    cublasSgemm(h, op(ta), op(tb), m, n, k, &alpha, dA, ar, dB, br, &beta, dC, m);
    CK(cudaMemcpy(C0, dC, nc * 4, cudaMemcpyDeviceToHost));
    double err = 0; for (size_t i = 0; i < nc; ++i) err = std::max(err, (double)std::fabs(C[i] - C0[i]));
    printf("  vs cuBLAS fp32: max|diff| = %.3e (/k = %.2e) %s\n", err, err / k, err / k < 2e-6 ? "OK" : "FAIL");
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    tmm_context_destroy(ctx);
}

// ---- complex<float>: the tcgen05 embedding (TMM_CMATH_TC) against cuBLAS CGEMM and the SIMT kernel -------------------------
static void cfill(std::vector<float>& v, unsigned seed, bool ints) { fill(v, seed, ints); }

static int ccheck_bench(cublasHandle_t h, char ta, char tb, int m, int n, int k, float ar_, float ai_, float br_, float bi_, int pad, bool ints, bool timing) {
    int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    int lda = ar + pad, ldb = br + pad, ldc = m + pad;
    std::vector<float> A((size_t)2 * lda * ac), B((size_t)2 * ldb * bc), C((size_t)2 * ldc * n);
    cfill(A, 5, ints); cfill(B, 6, ints); cfill(C, 7, ints);
    float *dA, *dB, *dC, *dR;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&dR, C.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const float alpha[2] = {ar_, ai_}, beta[2] = {br_, bi_};
    const cuComplex ca = make_cuComplex(ar_, ai_), cb = make_cuComplex(br_, bi_);
    CK(cudaMemcpy(dR, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
    cublasCgemm(h, op(ta), op(tb), m, n, k, &ca, (const cuComplex*)dA, lda, (const cuComplex*)dB, ldb, &cb, (cuComplex*)dR, ldc);
    std::vector<float> R(C.size()), O(C.size());
    CK(cudaMemcpy(R.data(), dR, C.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        tmm_set_c32_math(mode ? TMM_CMATH_TC : TMM_CMATH_SIMT);
        CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
        int rc = tmm_device_gemm(TMM_C32, ta, tb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, nullptr);
        cudaError_t se = cudaDeviceSynchronize();
        if (rc || se != cudaSuccess) { printf("cgemm %s %c%c: rc=%d (%s) sync=%s FAIL\n", mode ? "tc" : "simt", ta, tb, rc, tmm_last_error(), cudaGetErrorString(se)); fflush(stdout); exit(2); }
        CK(cudaMemcpy(O.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0; size_t clobber = 0;
        for (int j = 0; j < n; ++j) for (int i = 0; i < 2 * ldc; ++i) {
            size_t idx = (size_t)j * 2 * ldc + i;
            if (i < 2 * m) { double d = std::fabs((double)O[idx] - (double)R[idx]); if (!(d == d)) d = 1e30; err = std::max(err, d); }
            else if (O[idx] != C[idx]) ++clobber;
        }
        double rel = err / (std::max(1, k) * (ints ? 81.0 : 1.0) * 2);
        bool ok = rel < 4e-6 && !clobber && (!ints || err == 0);
        if (!ok) ++bad;
        double ms = 0;
        if (timing) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int r = 0; r < 4; ++r) {
                CK(cudaEventRecord(e0)); tmm_device_gemm(TMM_C32, ta, tb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, nullptr); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t, e0, e1)); if (r) best = std::min(best, t);
            }
            ms = best;
        }
        printf("cgemm %-4s %c%c m=%d n=%d k=%d alpha=(%g,%g) beta=(%g,%g) pad=%d %s vs cuBLAS: max|diff|=%.3e rel=%.2e pad-clobber=%zu %s", mode ? "tc" : "simt", ta, tb, m, n, k,
               ar_, ai_, br_, bi_, pad, ints ? "ints" : "rand", err, rel, clobber, ok ? "OK" : "FAIL");
        if (timing) printf("  %.3f ms = %.1f TF (8mnk)", ms, 8.0 * m * (double)n * k / ms * 1e-9);
        printf("\n"); fflush(stdout);
    }
    if (timing) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0)); cublasCgemm(h, op(ta), op(tb), m, n, k, &ca, (const cuComplex*)dA, lda, (const cuComplex*)dB, ldb, &cb, (cuComplex*)dR, ldc); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t, e0, e1)); if (r) best = std::min(best, t);
        }
        printf("cgemm cuBLAS %c%c %d %d %d: %.3f ms = %.1f TF (8mnk)\n", ta, tb, m, n, k, best, 8.0 * m * (double)n * k / best * 1e-9);
    }
    tmm_set_c32_math(TMM_CMATH_SIMT);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
    return bad;
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "check";
    cublasHandle_t h; cublasCreate(&h);
    cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
    if (mode == "probe") {  // probe ta tb
        char ta = argv[2][0], tb = argv[3][0];
        int w = probe(ta, tb, true) + probe(ta, tb, false);
        return w ? 3 : 0;
    }
    if (mode == "one") {  // one ta tb m n k alpha beta pad ints cublasref
        double r = check(h, argv[2][0], argv[3][0], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), (float)atof(argv[7]), (float)atof(argv[8]), atoi(argv[9]), atoi(argv[10]) != 0,
                         atoi(argv[11]) != 0, "one");
        return r < 2e-6 ? 0 : 3;
    }
    if (mode == "check") {  // check ta tb : the full shape list for one transpose pair
        char ta = argv[2][0], tb = argv[3][0];
        int bad = 0;
        auto run = [&](int m, int n, int k, float al, float be, int pad, bool ints, bool cub) { if (!(check(h, ta, tb, m, n, k, al, be, pad, ints, cub, "chk") < 2e-6)) ++bad; };
        run(128, 128, 32, 1.f, 0.f, 0, true, false);
        run(128, 128, 32, 1.f, 0.f, 0, false, false);
        run(128, 128, 256, 1.f, 0.f, 0, false, false);
        run(256, 384, 96, 1.f, 0.f, 0, false, false);
        run(257, 131, 77, 1.5f, 0.f, 3, false, false);
        run(5, 2, 2, 1.f, -0.5f, 1, false, false);
        run(50, 200, 21, 2.f, 0.f, 7, true, false);
        run(1000, 1000, 1000, 1.f, 1.f, 0, true, true);
        run(1000, 1000, 1000, 1.f, 1.f, 0, false, true);
        run(3001, 2003, 1099, -1.f, 0.5f, 9, false, true);
        run(4096, 4096, 4096, 1.f, 0.f, 0, false, true);
        tmm_set_f32_math(TMM_MATH_TF32);
        double r = check(h, ta, tb, 512, 512, 512, 1.f, 0.f, 0, false, false, "tf32-mode(expect ~1e-4)");
        if (!(r < 2e-3)) ++bad;
        tmm_set_f32_math(TMM_MATH_FP32);
        printf("check %c%c: %d failing\n", ta, tb, bad);
        return bad ? 3 : 0;
    }
    if (mode == "precision") {  // precision [k...]
        precision(h, 'N', 'N', 512, 512, 256, false);
        precision(h, 'N', 'N', 512, 512, 4096, false);
        precision(h, 'N', 'T', 512, 512, 4096, false);
        precision(h, 'N', 'N', 512, 512, 4096, true);
        precision(h, 'T', 'N', 384, 384, 32768, false);
        return 0;
    }
    if (mode == "bench") {
        bench(h, 'N', 'N', 8192, 8192, 8192, 0.f);
        bench(h, 'T', 'N', 8192, 8192, 8192, 0.f);
        bench(h, 'N', 'T', 8192, 8192, 8192, 0.f);
        bench(h, 'T', 'T', 8192, 8192, 8192, 0.f);
        bench(h, 'N', 'N', 10000, 10000, 10000, 0.f);
        bench(h, 'N', 'N', 10000, 4800, 512, 1.f);
        bench(h, 'N', 'N', 10000, 2048, 10000, 0.f);
        return 0;
    }
    if (mode == "cgemm") {  // all nine op pairs, integer (exact) and random data, padded lds, complex alpha / beta; then timing
        int bad = 0;
        const char ops[3] = {'N', 'T', 'C'};
        cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
        for (char ta : ops) for (char tb : ops) {
            bad += ccheck_bench(h, ta, tb, 130, 67, 95, 1.f, -2.f, 2.f, 1.f, 3, true, false);
            bad += ccheck_bench(h, ta, tb, 777, 530, 1111, 1.5f, -0.5f, 0.25f, 0.f, 1, false, false);
        }
        bad += ccheck_bench(h, 'N', 'N', 1000, 1000, 1000, 1.f, 0.f, 0.f, 0.f, 0, false, false);
        printf("cgemm check: %d failing\n", bad);
        ccheck_bench(h, 'N', 'N', 8192, 8192, 8192, 1.f, 0.f, 0.f, 0.f, 0, false, true);
        ccheck_bench(h, 'C', 'N', 8192, 8192, 8192, 1.f, 0.f, 0.f, 0.f, 0, false, true);
        ccheck_bench(h, 'N', 'T', 8192, 8192, 8192, 1.f, 0.f, 1.f, 0.f, 0, false, true);
        ccheck_bench(h, 'N', 'N', 10000, 2048, 10000, 1.f, 0.f, 0.f, 0.f, 0, false, true);
        return bad ? 3 : 0;
    }
    if (mode == "benchone") { bench(h, argv[2][0], argv[3][0], atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), (float)atof(argv[7])); return 0; }
    if (mode == "host") {
        host_gemm(h, 'N', 'N', 1000, 1000, 1000, 1.f, 2);
        host_gemm(h, 'N', 'N', 10000, 10000, 10000, 0.f, 3);
        host_gemm(h, 'T', 'N', 10000, 10000, 10000, 1.f, 2);
        return 0;
    }
    printf("unknown mode\n");
    return 1;
}
