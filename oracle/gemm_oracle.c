/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the arithmetic the reference hot path performs:
 *   gpu::gemm<Scalar>  (reference src/Tiled-MM/tiled_mm.cpp:492-624)
 *     -> round_robin / round_robin_without_copy_c      (tiled_mm.cpp:270-365, 367-454)
 *     -> cublas_gemm_wrapper -> blas_api::{s,d,c,z}gemm (tiled_mm.cpp:181-268, gpu_blas_api.hpp:194-252)
 *
 * The arithmetic itself lives in NVIDIA cuBLAS (closed source, no pinned version; 12.9.1.4 in this
 * image), which is absent from /root/reference.  What is restated here is its published contract,
 * BLAS xGEMM:   C = alpha * op(A) * op(B) + beta * C,   column-major, op in {N, T, C},
 * as the reference drives it:
 *   - stored shapes: A is (N ? m x k : k x m), B is (N ? k x n : n x k)   tiled_mm.cpp:507-514
 *   - op mapping 'T' -> transpose, 'C' -> conjugate transpose, else none    tiled_mm.cpp:168-179
 *   - k-tile accumulation uses beta' = (k_tile == 0 ? beta : 1)             tiled_mm.cpp:309
 *   - host C is read only when |beta| > 0                                   tiled_mm.cpp:325,423
 *     (so NaN/Inf in C must not propagate when beta == 0)
 *
 * Parity pinning: see oracle/README.md — checked against (a) the unmodified reference sources run
 * on the CPU over emulated CUDA/cuBLAS entry points (oracle/_ref/libtiledmm_ref_cpu.so, built by
 * oracle/Makefile), (b) the reference itself + cuBLAS on the GPU box (oracle/_ref/libtiledmm_ref.so),
 * and (c) the reference test's exact integer fixture (tests/test-multiply.cpp:58-66).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <complex.h>
#include <math.h>

#define ORACLE_F32 0
#define ORACLE_F64 1
#define ORACLE_C32 2
#define ORACLE_C64 3

static int is_zero_d(double re, double im) { return re == 0.0 && im == 0.0; }

/* op(X) packed into a dense column-major rows x cols buffer (conjugated when op == 'C'). */
#define DEFINE_PACK(NAME, T, CONJ)                                                              \
    static T* NAME(char op, int64_t rows, int64_t cols, const T* x, int64_t ld) {               \
        T* p = (T*)malloc(sizeof(T) * (size_t)(rows > 0 ? rows : 1) * (size_t)(cols > 0 ? cols : 1)); \
        if (!p) return NULL;                                                                    \
        if (op == 'N') {                                                                        \
            for (int64_t j = 0; j < cols; ++j)                                                  \
                memcpy(p + j * rows, x + j * ld, sizeof(T) * (size_t)rows);                     \
        } else {                                                                                \
            /* stored matrix is cols x rows: op(X)[i,j] = X[j,i] */                             \
            _Pragma("omp parallel for schedule(static)")                                        \
            for (int64_t j = 0; j < cols; ++j)                                                  \
                for (int64_t i = 0; i < rows; ++i) {                                            \
                    T v = x[i * ld + j];                                                        \
                    p[j * rows + i] = (op == 'C') ? CONJ(v) : v;                                \
                }                                                                               \
        }                                                                                       \
        return p;                                                                               \
    }

#define ID(x) (x)
DEFINE_PACK(pack_f32, float, ID)
DEFINE_PACK(pack_f64, double, ID)
DEFINE_PACK(pack_c32, float complex, conjf)
DEFINE_PACK(pack_c64, double complex, conj)

/* One column of C at a time, k in increasing order, accumulate in the scalar's own precision
 * (ACC == T) or a wider type (the *_wide entry point, used for error bars). */
#define DEFINE_GEMM(NAME, T, ACC, PACK)                                                         \
    static int NAME(char ta, char tb, int64_t m, int64_t n, int64_t k, T alpha, const T* a,     \
                    int64_t lda, const T* b, int64_t ldb, T beta, T* c, int64_t ldc,            \
                    int beta_is_zero, int alpha_is_zero) {                                      \
        if (m <= 0 || n <= 0) return 0;                                                         \
        T* pa = NULL; T* pb = NULL;                                                             \
        if (k > 0 && !alpha_is_zero) {                                                          \
            pa = PACK(ta, m, k, a, lda);                                                        \
            pb = PACK(tb, k, n, b, ldb);                                                        \
            if (!pa || !pb) { free(pa); free(pb); return -1; }                                  \
        }                                                                                       \
        _Pragma("omp parallel")                                                                 \
        {                                                                                       \
            ACC* acc = (ACC*)malloc(sizeof(ACC) * (size_t)m);                                   \
            _Pragma("omp for schedule(static)")                                                 \
            for (int64_t j = 0; j < n; ++j) {                                                   \
                for (int64_t i = 0; i < m; ++i) acc[i] = 0;                                     \
                if (pa) {                                                                       \
                    for (int64_t p = 0; p < k; ++p) {                                           \
                        ACC bv = (ACC)pb[j * k + p];                                            \
                        const T* ap = pa + p * m;                                               \
                        for (int64_t i = 0; i < m; ++i) acc[i] += (ACC)ap[i] * bv;              \
                    }                                                                           \
                }                                                                               \
                T* cj = c + j * ldc;                                                            \
                if (beta_is_zero) {                                                             \
                    for (int64_t i = 0; i < m; ++i) cj[i] = (T)((ACC)alpha * acc[i]);           \
                } else {                                                                        \
                    for (int64_t i = 0; i < m; ++i)                                             \
                        cj[i] = (T)((ACC)alpha * acc[i] + (ACC)beta * (ACC)cj[i]);              \
                }                                                                               \
            }                                                                                   \
            free(acc);                                                                          \
        }                                                                                       \
        free(pa); free(pb);                                                                     \
        return 0;                                                                               \
    }

DEFINE_GEMM(gemm_f32, float, float, pack_f32)
DEFINE_GEMM(gemm_f64, double, double, pack_f64)
DEFINE_GEMM(gemm_c32, float complex, float complex, pack_c32)
DEFINE_GEMM(gemm_c64, double complex, double complex, pack_c64)
DEFINE_GEMM(gemm_f32_wide, float, double, pack_f32)
DEFINE_GEMM(gemm_f64_wide, double, long double, pack_f64)
DEFINE_GEMM(gemm_c32_wide, float complex, double complex, pack_c32)
DEFINE_GEMM(gemm_c64_wide, double complex, long double complex, pack_c64)

static int norm_op(char t, char* out) {
    char u = (char)toupper((unsigned char)t);          /* tiled_mm.cpp:503-504 */
    if (u != 'N' && u != 'T' && u != 'C') return -1;   /* accept only N/T/C */
    *out = u;
    return 0;
}

/* Returns 0 on success, -1 out of memory, -2 bad argument. */
int oracle_gemm_ex(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k,
                   const void* alpha, const void* a, int64_t lda, const void* b, int64_t ldb,
                   const void* beta, void* c, int64_t ldc, int wide) {
    char ta, tb;
    if (norm_op(trans_a, &ta) || norm_op(trans_b, &tb)) return -2;
    if (m < 0 || n < 0 || k < 0) return -2;
    int64_t a_rows = ta == 'N' ? m : k, b_rows = tb == 'N' ? k : n;
    if (lda < (a_rows > 1 ? a_rows : 1) || ldb < (b_rows > 1 ? b_rows : 1) || ldc < (m > 1 ? m : 1)) return -2;
    switch (dtype) {
    case ORACLE_F32: {
        float al = *(const float*)alpha, be = *(const float*)beta;
        return (wide ? gemm_f32_wide : gemm_f32)(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, be == 0.0f, al == 0.0f);
    }
    case ORACLE_F64: {
        double al = *(const double*)alpha, be = *(const double*)beta;
        return (wide ? gemm_f64_wide : gemm_f64)(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, be == 0.0, al == 0.0);
    }
    case ORACLE_C32: {
        float complex al = *(const float complex*)alpha, be = *(const float complex*)beta;
        return (wide ? gemm_c32_wide : gemm_c32)(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc,
                                                 is_zero_d(crealf(be), cimagf(be)), is_zero_d(crealf(al), cimagf(al)));
    }
    case ORACLE_C64: {
        double complex al = *(const double complex*)alpha, be = *(const double complex*)beta;
        return (wide ? gemm_c64_wide : gemm_c64)(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc,
                                                 is_zero_d(creal(be), cimag(be)), is_zero_d(creal(al), cimag(al)));
    }
    default: return -2;
    }
}

int oracle_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k,
                const void* alpha, const void* a, int64_t lda, const void* b, int64_t ldb,
                const void* beta, void* c, int64_t ldc) {
    return oracle_gemm_ex(dtype, trans_a, trans_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 0);
}

/* ---- tiling math of the reference (pure integer) ------------------------------------------ */

/* optimal_tile_size, reference mm_handle.cpp:89-110.  The curr_tile_size argument only matters
 * through "curr < max && dim <= max -> dim"; when dim <= max the divisor search below also
 * returns dim, so the result is a function of (dim, max) alone. */
int oracle_optimal_tile_size(int dim, int max_tile) {
    if (dim <= max_tile) return dim;
    int limit = max_tile < dim ? max_tile : dim;
    int tile = 1;
    for (int i = 1; i <= limit; ++i)
        if (dim % i == 0) tile = i;
    if (abs(max_tile - tile) <= max_tile / 2) return tile;
    return limit;
}

/* number of tiles along one dimension, tiled_matrix.cpp:13-17 (clamp, then ceil). */
int oracle_num_tiles(int dim, int tile) {
    if (tile > dim) tile = dim;
    return (dim + tile - 1) / tile;
}

/* H2D bytes the reference schedule moves (round_robin, tiled_mm.cpp:292-330):
 * A and B tiles are re-sent for every (m_tile, n_tile): n_tiles_n*|A| + n_tiles_m*|B| (+|C| if beta != 0). */
int64_t oracle_reference_h2d_bytes(int64_t m, int64_t n, int64_t k, int tile_m, int tile_n, int64_t elem, int beta_nonzero) {
    int64_t ntm = oracle_num_tiles((int)m, tile_m), ntn = oracle_num_tiles((int)n, tile_n);
    return elem * (ntn * m * k + ntm * k * n + (beta_nonzero ? m * n : 0));
}
