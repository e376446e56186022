// test-multiply - the correctness app of the reference (tests/test-multiply.cpp) on tiled_mm_b200: same flags, same verdict
// line ("The result is CORRECT" / "NOT CORRECT") and exit code (0 / 1), same inputs (mt19937(42), ints 0..9, A -> B -> C).
// The reference checks the tiled result against ONE-SHOT cuBLAS on the same GPU (test-multiply.cpp:16-53).  Here the one-shot
// comparator is blas_api::?gemm of this library (the same device kernels without the scheduler), so on its own that check would
// only prove the scheduler; every run therefore ALSO verifies sampled entries of C against a long-double dot product computed
// on the host from the original inputs, which is independent of every kernel in the library.
// Additions: --type s|d|c|z, -t accepts C, --gpus N, leading dimensions are honoured when sizing buffers.
#include <Tiled-MM/device_vector.hpp>
#include <Tiled-MM/gpu_blas_api.hpp>
#include <Tiled-MM/gpu_blas_handle.hpp>
#include <Tiled-MM/tiled_mm.hpp>
#include <Tiled-MM/util.hpp>

#include "cli.hpp"

#include <chrono>
#include <cstdlib>
#include <complex>
#include <random>

namespace {

template <typename T> struct scalar_traits { using real = T; static constexpr bool is_complex = false; };
template <typename R> struct scalar_traits<std::complex<R>> { using real = R; static constexpr bool is_complex = true; };

std::mt19937& generator() {
    static std::mt19937 rng(42);  // the reference's fixed seed (tests/test-multiply.cpp:60)
    return rng;
}
template <typename T>
void fill_matrix(T* ptr, size_t count) {
    static std::uniform_int_distribution<int> digit(0, 9);
    for (size_t i = 0; i < count; ++i) ptr[i] = static_cast<T>(static_cast<typename scalar_traits<T>::real>(digit(generator())));
}

template <typename T>
void print_matrix(const T* mat, long long rows, long long cols, long long ld) {
    for (long long i = 0; i < rows; ++i) {
        for (long long j = 0; j < cols; ++j) std::cout << mat[j * ld + i] << "\t";
        std::cout << "\n";
    }
    std::cout << std::endl;
}

inline gpu::blas_api::StatusType one_shot(void* st, gpu::blas_api::OperationType ta, gpu::blas_api::OperationType tb, int m, int n, int k, const float* al,
                                          const float* a, int lda, const float* b, int ldb, const float* be, float* c, int ldc) {
    return gpu::blas_api::sgemm(st, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
inline gpu::blas_api::StatusType one_shot(void* st, gpu::blas_api::OperationType ta, gpu::blas_api::OperationType tb, int m, int n, int k, const double* al,
                                          const double* a, int lda, const double* b, int ldb, const double* be, double* c, int ldc) {
    return gpu::blas_api::dgemm(st, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
inline gpu::blas_api::StatusType one_shot(void* st, gpu::blas_api::OperationType ta, gpu::blas_api::OperationType tb, int m, int n, int k,
                                          const std::complex<float>* al, const std::complex<float>* a, int lda, const std::complex<float>* b, int ldb,
                                          const std::complex<float>* be, std::complex<float>* c, int ldc) {
    return gpu::blas_api::cgemm(st, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}
inline gpu::blas_api::StatusType one_shot(void* st, gpu::blas_api::OperationType ta, gpu::blas_api::OperationType tb, int m, int n, int k,
                                          const std::complex<double>* al, const std::complex<double>* a, int lda, const std::complex<double>* b, int ldb,
                                          const std::complex<double>* be, std::complex<double>* c, int ldc) {
    return gpu::blas_api::zgemm(st, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
}

// C_ref = alpha op(A) op(B) + beta C by one device GEMM over whole matrices (what compute_reference does with cuBLAS)
template <typename T>
void one_shot_reference(const cli::Problem& p, const T* a, const T* b, T* c, T alpha, T beta) {
    const size_t na = (size_t)p.ld_a * p.a_cols, nb = (size_t)p.ld_b * p.b_cols, nc = (size_t)p.ld_c * p.n;
    gpu::device_vector<T> a_device(na), b_device(nb), c_device(nc);
    gpu::copy_to_device(a, a_device.data(), na);
    gpu::copy_to_device(b, b_device.data(), nb);
    gpu::copy_to_device(c, c_device.data(), nc);
    gpu::gpu_blas_handle handle;
    gpu::check_blas_status(one_shot(handle.handle(), gpu::get_blas_operation(p.trans_a), gpu::get_blas_operation(p.trans_b), (int)p.m, (int)p.n, (int)p.k, &alpha,
                                    a_device.data(), (int)p.ld_a, b_device.data(), (int)p.ld_b, &beta, c_device.data(), (int)p.ld_c));
    gpu::check_runtime_status(gpu::runtime_api::device_synchronize());
    gpu::copy_to_host(c_device.data(), c, nc);
}

template <typename T>
std::complex<long double> widen(const T& v) { return std::complex<long double>((long double)std::real(v), (long double)std::imag(v)); }

// entry (i, j) of alpha op(A) op(B) + beta C0 in long double, straight from the definition
template <typename T>
std::complex<long double> host_entry(const cli::Problem& p, const T* a, const T* b, const T* c0, T alpha, T beta, long long i, long long j) {
    std::complex<long double> sum = 0;
    for (long long l = 0; l < p.k; ++l) {
        std::complex<long double> av = widen(p.trans_a == 'N' ? a[l * p.ld_a + i] : a[i * p.ld_a + l]);
        std::complex<long double> bv = widen(p.trans_b == 'N' ? b[j * p.ld_b + l] : b[l * p.ld_b + j]);
        if (p.trans_a == 'C') av = std::conj(av);
        if (p.trans_b == 'C') bv = std::conj(bv);
        sum += av * bv;
    }
    std::complex<long double> out = widen(alpha) * sum;
    if (std::abs(beta) > 0) out += widen(beta) * widen(c0[j * p.ld_c + i]);
    return out;
}

template <typename T>
bool same_block(const T* v1, const T* v2, const cli::Problem& p, double eps = 1e-6) {
    for (long long j = 0; j < p.n; ++j)
        for (long long i = 0; i < p.m; ++i)
            if (std::abs(v1[j * p.ld_c + i] - v2[j * p.ld_c + i]) > eps) return false;
    return true;
}

template <typename T>
int run(const cli::Problem& p) {
    const bool small_sizes = std::max(p.m, std::max(p.n, p.k)) < 20;
    const size_t na = (size_t)p.ld_a * p.a_cols, nb = (size_t)p.ld_b * p.b_cols, nc = (size_t)p.ld_c * p.n;
    T* a_host = gpu::malloc_pinned<T>(na, T(1));
    T* b_host = gpu::malloc_pinned<T>(nb, T(1));
    T* c_host = gpu::malloc_pinned<T>(nc, T(0));
    T* c_host2 = gpu::malloc_pinned<T>(nc, T(0));
    T* c_initial = gpu::malloc_pinned<T>(nc, T(0));
    T* c_reference = gpu::malloc_pinned<T>(nc, T(0));
    fill_matrix(a_host, na);
    fill_matrix(b_host, nb);
    fill_matrix(c_host, nc);
    std::copy(c_host, c_host + nc, c_host2);
    std::copy(c_host, c_host + nc, c_initial);
    std::copy(c_host, c_host + nc, c_reference);
    const T alpha = T(p.alpha), beta = T(p.beta);

    if (small_sizes) {
        std::cout << "Initial values in matrix A: " << std::endl; print_matrix(a_host, p.a_rows, p.a_cols, p.ld_a);
        std::cout << "Initial values in matrix B: " << std::endl; print_matrix(b_host, p.b_rows, p.b_cols, p.ld_b);
        std::cout << "Initial values in matrix C: " << std::endl; print_matrix(c_host, p.m, p.n, p.ld_c);
    }
    one_shot_reference(p, a_host, b_host, c_reference, alpha, beta);
    if (small_sizes) { std::cout << "Correct result C = beta*C + alpha*A*B: " << std::endl; print_matrix(c_reference, p.m, p.n, p.ld_c); }

    auto ctx = gpu::make_context<T>((int)p.n_streams, (int)p.tile_m, (int)p.tile_n, (int)p.tile_k);
    if (p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), (int)p.gpus, nullptr));

    // VERSION WITH COPYING C BACK
    auto start = std::chrono::steady_clock::now();
    gpu::gemm64<T>(*ctx, p.trans_a, p.trans_b, p.m, p.n, p.k, alpha, a_host, p.ld_a, b_host, p.ld_b, beta, c_host, p.ld_c, false, true);
    std::cout << "Time [ms] with copying C back: " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count() << std::endl;
    if (small_sizes) { std::cout << "Computed result by Tiled-MM with copying C back : " << std::endl; print_matrix(c_host, p.m, p.n, p.ld_c); }
    bool correct = same_block(c_host, c_reference, p);

    // VERSION WITHOUT COPYING C BACK: the result stays in the context's device C, column-major m x n with ld = m
    if (p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), 1, nullptr));
    start = std::chrono::steady_clock::now();
    gpu::gemm64<T>(*ctx, p.trans_a, p.trans_b, p.m, p.n, p.k, alpha, a_host, p.ld_a, b_host, p.ld_b, beta, c_host2, p.ld_c, false, false);
    std::cout << "Time [ms] without copying C back: " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count() << std::endl;
    {
        const size_t compact = (size_t)p.m * p.n;
        T* dense = gpu::malloc_pinned<T>(compact, T(0));
        gpu::copy_to_host(ctx->get_full_device_buffer_c().data(), dense, compact);
        for (long long j = 0; j < p.n; ++j) std::copy(dense + j * p.m, dense + (j + 1) * p.m, c_host2 + j * p.ld_c);
        tmm_free_pinned(dense);
    }
    if (small_sizes) { std::cout << "Computed result by Tiled-MM without copying C back : " << std::endl; print_matrix(c_host2, p.m, p.n, p.ld_c); }
    correct = correct && same_block(c_host2, c_reference, p);

    // independent of every kernel in the library: sampled entries against the definition, long double on the host
    {
        std::mt19937_64 pick(7);
        const int samples = (int)std::min<long long>(256, p.m * p.n);
        long double worst = 0;
        for (int s = 0; s < samples; ++s) {
            const long long i = (long long)(pick() % (unsigned long long)p.m), j = (long long)(pick() % (unsigned long long)p.n);
            const std::complex<long double> want = host_entry(p, a_host, b_host, c_initial, alpha, beta, i, j);
            worst = std::max(worst, std::abs(want - widen(c_host[j * p.ld_c + i])));
            worst = std::max(worst, std::abs(want - widen(c_host2[j * p.ld_c + i])));
        }
        // integer inputs: exact in FP64; in FP32 exact while |C| < 2^24, else within FP32 rounding of the largest term sum
        const long double tol = sizeof(typename scalar_traits<T>::real) == 8 ? 1e-6L : 1e-6L + 1.2e-7L * 81.0L * (long double)p.k * (long double)(std::abs(p.alpha) + 1) * 4;
        std::cout << "Host check of " << samples << " sampled entries (long double): max abs error = " << (double)worst << std::endl;
        correct = correct && worst <= tol;
    }
    std::cout << "The result is " << (correct ? "CORRECT" : "NOT CORRECT") << std::endl;
    for (T* q : {a_host, b_host, c_host, c_host2, c_initial, c_reference}) tmm_free_pinned(q);
    return correct ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
    cli::Args args(cli::gemm_options(false));
    if (!args.read(argc, argv)) return 2;
    if (args.help_requested) { args.usage("test-multiply", "Testing Tiled-MM: checks the result of the tiled out-of-core GEMM."); return 0; }
    cli::Problem p;
    if (!cli::problem_from(args, &p)) return 0;
    if (p.gpus > 1) setenv("TMM_PINNED_NUMA", "interleave", 0);  // one copy of A, B, C read by the GPUs of both sockets: spread its pages
    cli::print_banner(p, 1);
    try {
        switch (p.type) {
        case 's': return run<float>(p);
        case 'c': return run<std::complex<float>>(p);
        case 'z': return run<std::complex<double>>(p);
        default: return run<double>(p);
        }
    } catch (const std::exception& e) {
        std::cerr << "test-multiply: " << e.what() << std::endl;
        return 1;
    }
}
