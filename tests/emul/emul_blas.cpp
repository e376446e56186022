// TEST INFRASTRUCTURE ONLY - never part of the product.
// Stand-in for the device GEMM layer (csrc/tmm_blas.h: the hand-written sm_100a kernels) when the scheduler runs over the CPU
// emulation of the CUDA runtime: every launch the scheduler issues is executed by the oracle's GEMM on the "device" operands,
// after checking what the real kernels rely on - operands inside their allocations, and A/B meeting the TMA contract
// (16-byte aligned base, pitch a multiple of 16 bytes), which the scheduler promises for every panel it builds.
#include "../../tiled-mm_b200/csrc/tmm_blas.h"

#include <atomic>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <complex>
#include <vector>

extern "C" int oracle_gemm_ex(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t lda, const void* b,
                              int64_t ldb, const void* beta, void* c, int64_t ldc, int wide);
extern "C" void emul_check_device_range(const void* p, size_t bytes);
extern "C" int emul_dry_run();  // address-only run: check and count, do not compute
extern "C" int emul_op_begin(void* stream);  // race detector: a launch enters the stream ...
extern "C" void emul_op_access(int sid, const void* p, size_t pitch, size_t width, size_t height, int write, const char* what);  // ... and touches these regions

namespace {
std::atomic<uint64_t> g_launches{0}, g_contract_violations{0}, g_prepared{0};
std::atomic<int> g_f32_mode{3}, g_c32_mode{0};
}

extern "C" __attribute__((visibility("default"))) uint64_t emul_tma_contract_violations() { return g_contract_violations; }
extern "C" __attribute__((visibility("default"))) uint64_t emul_prepared_cgemm_launches() { return g_prepared; }
// the counter is about panels the SCHEDULER builds; user-supplied device operands may have any ld (the real dispatcher re-pitches them)
extern "C" __attribute__((visibility("default"))) void emul_reset_tma_contract_violations() { g_contract_violations = 0; }

namespace tmm {

void count_launch() { g_launches.fetch_add(1); }
uint64_t launch_count() { return g_launches.load(); }
int sm_count() { return 148; }
int f32_math_mode() { return g_f32_mode; }
void set_f32_math_mode(int m) { g_f32_mode = m; }
int c32_math_mode() { return g_c32_mode; }
void set_c32_math_mode(int m) { g_c32_mode = m; }

cudaError_t device_scale(int dtype, int64_t m, int64_t n, const void* beta, void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    const size_t es = dtype_size(dtype);
    emul_check_device_range(c, ((size_t)(n - 1) * ldc + m) * es);
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, c, (size_t)ldc * es, (size_t)m * es, (size_t)n, 1, "scale C");
    }
    // C = beta * C is the GEMM with k = 0
    count_launch();
    const double one[2] = {0, 0};
    if (emul_dry_run()) return cudaSuccess;
    return oracle_gemm_ex(dtype, 'N', 'N', m, n, 0, one, c, m > 1 ? m : 1, c, 1, beta, c, ldc, 0) == 0 ? cudaSuccess : cudaErrorInvalidValue;
}

// C += beta * S: element-wise on the host, in the arithmetic of the element type (one multiply, one add - what the device kernel does)
template <typename T>
static void add_scaled_host(int64_t m, int64_t n, const void* beta, const void* s, int64_t lds, void* c, int64_t ldc) {
    const T b = *static_cast<const T*>(beta);
    const T* sp = static_cast<const T*>(s);
    T* cp = static_cast<T*>(c);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < m; ++i) cp[j * ldc + i] = cp[j * ldc + i] + b * sp[j * lds + i];
}
cudaError_t device_add_scaled(int dtype, int64_t m, int64_t n, const void* beta, const void* s, int64_t lds, void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    const size_t es = dtype_size(dtype);
    emul_check_device_range(c, ((size_t)(n - 1) * ldc + m) * es);
    emul_check_device_range(s, ((size_t)(n - 1) * lds + m) * es);
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, s, (size_t)lds * es, (size_t)m * es, (size_t)n, 0, "add_scaled S");
        emul_op_access(sid, c, (size_t)ldc * es, (size_t)m * es, (size_t)n, 1, "add_scaled C");
    }
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    switch (dtype) {
    case F32: add_scaled_host<float>(m, n, beta, s, lds, c, ldc); break;
    case F64: add_scaled_host<double>(m, n, beta, s, lds, c, ldc); break;
    case C32: add_scaled_host<std::complex<float>>(m, n, beta, s, lds, c, ldc); break;
    case C64: add_scaled_host<std::complex<double>>(m, n, beta, s, lds, c, ldc); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

cudaError_t device_gemm(int dtype, char trans_a, char trans_b, int64_t m, int64_t n, int64_t k, const void* alpha, const void* a, int64_t lda, const void* b,
                        int64_t ldb, const void* beta, void* c, int64_t ldc, cudaStream_t st) {
    const char ta = (char)std::toupper((unsigned char)trans_a), tb = (char)std::toupper((unsigned char)trans_b);
    if ((ta != 'N' && ta != 'T' && ta != 'C') || (tb != 'N' && tb != 'T' && tb != 'C') || m < 0 || n < 0 || k < 0) return cudaErrorInvalidValue;
    if (m == 0 || n == 0) return cudaSuccess;
    if (k == 0) return device_scale(dtype, m, n, beta, c, ldc, st);
    const size_t es = dtype_size(dtype);
    const int64_t ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    emul_check_device_range(a, ((size_t)(ac - 1) * lda + ar) * es);
    emul_check_device_range(b, ((size_t)(bc - 1) * ldb + br) * es);
    emul_check_device_range(c, ((size_t)(n - 1) * ldc + m) * es);
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (((size_t)lda * es) & 15) || (((size_t)ldb * es) & 15)) {
        g_contract_violations.fetch_add(1);
        fprintf(stderr, "[emul] operand outside the TMA contract: a %p lda %lld b %p ldb %lld es %zu\n", a, (long long)lda, b, (long long)ldb, es);
    }
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, a, (size_t)lda * es, (size_t)ar * es, (size_t)ac, 0, "gemm A");
        emul_op_access(sid, b, (size_t)ldb * es, (size_t)br * es, (size_t)bc, 0, "gemm B");
        emul_op_access(sid, c, (size_t)ldc * es, (size_t)m * es, (size_t)n, 1, "gemm C");
    }
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    return oracle_gemm_ex(dtype, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 0) == 0 ? cudaSuccess : cudaErrorInvalidValue;
}

// ---- complex<float> on prepared operands (csrc/gemm_c32_tc.cu): the embedding written out on the host, the product by the oracle's float GEMM --------
// Same layouts and the same declared accesses as the device passes, so that the scheduler's offsets into A' / B' and its event wiring (a chunk of A
// embedded once on one stream, multiplied on several) are checked by the range checker and the race detector.
cudaError_t cgemm_tc_embed_a(char ta, int m, int k, const float* al, const void* a, int64_t lda, float* a2, int64_t pitch_a2, cudaStream_t st) {
    if (m <= 0 || k <= 0) return cudaSuccess;
    const bool a_n = ta == 'N';
    const int64_t ar = a_n ? m : k, ac = a_n ? k : m;           // stored extent (complex elements)
    const int64_t r2 = 2 * ar, c2 = 2 * ac;                     // extent of A' (a_n) or A'^T in floats
    emul_check_device_range(a, ((size_t)(ac - 1) * lda + ar) * 8);
    emul_check_device_range(a2, ((size_t)(c2 - 1) * pitch_a2 + r2) * 4);
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, a, (size_t)lda * 8, (size_t)ar * 8, (size_t)ac, 0, "embed A");
        emul_op_access(sid, a2, (size_t)pitch_a2 * 4, (size_t)r2 * 4, (size_t)c2, 1, "embed A'");
    }
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    const std::complex<float> alpha(al[0], al[1]);
    const std::complex<float>* ap = static_cast<const std::complex<float>*>(a);
    for (int64_t l = 0; l < k; ++l)
        for (int64_t i = 0; i < m; ++i) {
            std::complex<float> v = a_n ? ap[l * lda + i] : ap[i * lda + l];
            if (ta == 'C') v = std::conj(v);
            // the device kernel's cmul: (ar*br - ai*bi, ar*bi + ai*br) in float, no fused operations assumed on integer test data
            const float wr = alpha.real() * v.real() - alpha.imag() * v.imag(), wi = alpha.real() * v.imag() + alpha.imag() * v.real();
            if (a_n) {
                a2[(2 * l) * pitch_a2 + 2 * i] = wr;      a2[(2 * l) * pitch_a2 + 2 * i + 1] = wi;
                a2[(2 * l + 1) * pitch_a2 + 2 * i] = -wi; a2[(2 * l + 1) * pitch_a2 + 2 * i + 1] = wr;
            } else {
                a2[(2 * i) * pitch_a2 + 2 * l] = wr;      a2[(2 * i) * pitch_a2 + 2 * l + 1] = -wi;
                a2[(2 * i + 1) * pitch_a2 + 2 * l] = wi;  a2[(2 * i + 1) * pitch_a2 + 2 * l + 1] = wr;
            }
        }
    return cudaSuccess;
}

cudaError_t cgemm_tc_split_b(char tb, int n, int k, const void* b, int64_t ldb, float* b2, int64_t pitch_b2, cudaStream_t st) {
    if (n <= 0 || k <= 0) return cudaSuccess;
    emul_check_device_range(b, ((size_t)(k - 1) * ldb + n) * 8);
    emul_check_device_range(b2, ((size_t)(2 * k - 1) * pitch_b2 + n) * 4);
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, b, (size_t)ldb * 8, (size_t)n * 8, (size_t)k, 0, "split B");
        emul_op_access(sid, b2, (size_t)pitch_b2 * 4, (size_t)n * 4, (size_t)(2 * k), 1, "split B'");
    }
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    const std::complex<float>* bp = static_cast<const std::complex<float>*>(b);
    for (int64_t l = 0; l < k; ++l)
        for (int64_t j = 0; j < n; ++j) {
            const std::complex<float> v = bp[l * ldb + j];
            b2[(2 * l) * pitch_b2 + j] = v.real();
            b2[(2 * l + 1) * pitch_b2 + j] = tb == 'C' ? -v.imag() : v.imag();
        }
    return cudaSuccess;
}

cudaError_t cgemm_tc_prepared(char ta, char tb, int m, int n, int k, const float* a2, int64_t pitch_a2, const float* b2, int64_t pitch_b2, const float* be,
                              void* c, int64_t ldc, cudaStream_t st) {
    if (m <= 0 || n <= 0 || k <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(a2) & 15) || (reinterpret_cast<uintptr_t>(b2) & 15) || (pitch_a2 & 3) || (pitch_b2 & 3)) return cudaErrorInvalidValue;  // sgemm_tc_eligible
    const bool a_n = ta == 'N', b_n = tb == 'N';
    const int64_t a_r = a_n ? 2 * m : 2 * k, a_c = a_n ? 2 * k : 2 * m, b_r = b_n ? 2 * k : n, b_c = b_n ? n : 2 * k;
    emul_check_device_range(a2, ((size_t)(a_c - 1) * pitch_a2 + a_r) * 4);
    emul_check_device_range(b2, ((size_t)(b_c - 1) * pitch_b2 + b_r) * 4);
    emul_check_device_range(c, ((size_t)(n - 1) * ldc + m) * 8);
    float beta_r = be[0];
    if (be[1] != 0.f) {
        cudaError_t e = device_scale(C32, m, n, be, c, ldc, st);
        if (e != cudaSuccess) return e;
        beta_r = 1.f;
    }
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, a2, (size_t)pitch_a2 * 4, (size_t)a_r * 4, (size_t)a_c, 0, "cgemm A'");
        emul_op_access(sid, b2, (size_t)pitch_b2 * 4, (size_t)b_r * 4, (size_t)b_c, 0, "cgemm B'");
        emul_op_access(sid, c, (size_t)ldc * 8, (size_t)m * 8, (size_t)n, 1, "cgemm C");
    }
    g_prepared.fetch_add(1);
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    const float one = 1.f;
    return oracle_gemm_ex(F32, a_n ? 'N' : 'T', b_n ? 'N' : 'T', 2 * (int64_t)m, n, 2 * (int64_t)k, &one, a2, pitch_a2, b2, pitch_b2, &beta_r, c, 2 * ldc, 1) == 0 ? cudaSuccess
                                                                                                                                                          : cudaErrorInvalidValue;
}

// bf16-input GEMM: widen on the host, then the oracle's float GEMM (exact products, like the tensor-core path)
cudaError_t bgemm_tc_launch(char ta, char tb, int m, int n, int k, float alpha, const void* a, int64_t lda, const void* b, int64_t ldb, float beta, float* c,
                            int64_t ldc, cudaStream_t st) {
    const int ar = ta == 'N' ? m : k, ac = ta == 'N' ? k : m, br = tb == 'N' ? k : n, bc = tb == 'N' ? n : k;
    emul_check_device_range(a, ((size_t)(ac - 1) * lda + ar) * 2);
    emul_check_device_range(b, ((size_t)(bc - 1) * ldb + br) * 2);
    emul_check_device_range(c, ((size_t)(n - 1) * ldc + m) * 4);
    {
        const int sid = emul_op_begin(st);
        emul_op_access(sid, a, (size_t)lda * 2, (size_t)ar * 2, (size_t)ac, 0, "bgemm A");
        emul_op_access(sid, b, (size_t)ldb * 2, (size_t)br * 2, (size_t)bc, 0, "bgemm B");
        emul_op_access(sid, c, (size_t)ldc * 4, (size_t)m * 4, (size_t)n, 1, "bgemm C");
    }
    count_launch();
    if (emul_dry_run()) return cudaSuccess;
    auto widen = [](const void* p, int64_t ld, int rows, int cols) {
        std::vector<float> out((size_t)rows * cols);
        const uint16_t* q = static_cast<const uint16_t*>(p);
        for (int j = 0; j < cols; ++j)
            for (int i = 0; i < rows; ++i) { const uint32_t u = (uint32_t)q[(size_t)j * ld + i] << 16; memcpy(&out[(size_t)j * rows + i], &u, 4); }
        return out;
    };
    const std::vector<float> a32 = widen(a, lda, ar, ac), b32 = widen(b, ldb, br, bc);
    return oracle_gemm_ex(F32, ta, tb, m, n, k, &alpha, a32.data(), ar > 1 ? ar : 1, b32.data(), br > 1 ? br : 1, &beta, c, ldc, 1) == 0 ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace tmm
