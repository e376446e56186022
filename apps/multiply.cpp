// multiply - the miniapp of the reference (examples/multiply.cpp) on tiled_mm_b200: same flags, same printed report
// ("-> Avg Time [ms]", "-> Throughput [Gflops]" for the copy-back and the device-resident variant), so scripts such as the
// reference's examples/compare.sh keep working.  Additions: --type s|d|c|z, -t accepts C, --gpus N spreads the C blocks
// over N GPUs of the box (copy-back variant), times are fractional milliseconds (the reference's integer division
// reports 0 ms / inf Gflops for small sizes), flop counts are 64-bit doubles.
#include <Tiled-MM/tiled_mm.hpp>

#include "cli.hpp"

#include <chrono>
#include <cstdlib>
#include <complex>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

template <typename T> T make_scalar(double v) { return T(v); }
template <typename T> struct flops_per_fma { static constexpr double value = 2.0; };
template <typename R> struct flops_per_fma<std::complex<R>> { static constexpr double value = 8.0; };

// The reference fills A and B with ones (examples/multiply.cpp:156-160).  For the out-of-core sizes (hundreds of GB) a single-threaded
// std::fill takes minutes, and all-ones operands draw less power than real data; --random 1 fills with uniform(-1, 1) values from
// one xorshift generator per thread instead.  Either way the fill runs on all host cores.
template <typename T>
void parallel_fill(T* ptr, size_t count, bool random, unsigned seed) {
    const unsigned n_threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_threads; ++t)
        pool.emplace_back([=] {
            const size_t lo = count * t / n_threads, hi = count * (t + 1) / n_threads;
            if (!random) { std::fill(ptr + lo, ptr + hi, T(1)); return; }
            unsigned long long x = 0x9E3779B97F4A7C15ull * (seed * 64ull + t + 1);
            for (size_t i = lo; i < hi; ++i) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                ptr[i] = T((double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0);
            }
        });
    for (auto& th : pool) th.join();
}

// --scaling "8,1": the copy-back call at several GPU counts on ONE set of pinned host buffers (allocating and filling hundreds of GB
// dominates an out-of-core run), each checked by the linearity property C x = A (B x) on sampled rows, with the speed-up over the last
// count (strong scaling: the problem is fixed).  This is how BASELINE configs[3] / [4] are measured (tools/r2_c5.sh).
template <typename T>
double linearity_defect(const cli::Problem& p, const T* a, const T* b, const T* c, int rows_sampled) {
    // x = ones: (B x)[l] = sum_j op(B)[l, j];  then for sampled rows i: |sum_j C[i, j] - sum_l op(A)[i, l] (B x)[l]| / (k * n)
    const bool ta = p.trans_a != 'N' && p.trans_a != 'n', tb = p.trans_b != 'N' && p.trans_b != 'n';
    const bool ca = p.trans_a == 'C' || p.trans_a == 'c', cb = p.trans_b == 'C' || p.trans_b == 'c';
    // B x with x = ones: sums over the columns of op(B).  Walk the STORED matrix column by column (contiguous), whatever the op:
    //   op(B) = N: stored k x n, (B x)[l] = row sums  -> threads split the stored columns and keep a private accumulator vector
    //   op(B) = T/C: stored n x k, (B x)[l] = sum of stored column l -> threads split the stored columns, one scalar each
    std::vector<std::complex<double>> bx((size_t)p.k);
    const unsigned n_threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::vector<std::thread> pool;
    if (!tb) {
        std::vector<std::vector<std::complex<double>>> part(n_threads);
        for (unsigned t = 0; t < n_threads; ++t)
            pool.emplace_back([&, t] {
                part[t].assign((size_t)p.k, 0.0);
                for (long long j = p.n * t / n_threads; j < p.n * (t + 1) / n_threads; ++j) {
                    const T* col = b + (size_t)j * p.ld_b;
                    for (long long l = 0; l < p.k; ++l) part[t][(size_t)l] += std::complex<double>(col[l]);
                }
            });
        for (auto& th : pool) th.join();
        for (long long l = 0; l < p.k; ++l) {
            std::complex<double> sum = 0;
            for (unsigned t = 0; t < n_threads; ++t) sum += part[t][(size_t)l];
            bx[(size_t)l] = sum;
        }
    } else {
        for (unsigned t = 0; t < n_threads; ++t)
            pool.emplace_back([&, t] {
                for (long long l = p.k * t / n_threads; l < p.k * (t + 1) / n_threads; ++l) {
                    const T* col = b + (size_t)l * p.ld_b;
                    std::complex<double> sum = 0;
                    for (long long j = 0; j < p.n; ++j) sum += std::complex<double>(col[j]);
                    bx[(size_t)l] = cb ? std::conj(sum) : sum;
                }
            });
        for (auto& th : pool) th.join();
    }
    double worst = 0;
    for (int r = 0; r < rows_sampled; ++r) {
        const long long i = (long long)((double)r / rows_sampled * (double)p.m) + (r * 7) % std::max<long long>(1, p.m / rows_sampled);
        if (i >= p.m) continue;
        std::complex<double> lhs = 0, rhs = 0;
        for (long long j = 0; j < p.n; ++j) lhs += std::complex<double>(c[(size_t)j * p.ld_c + i]);
        for (long long l = 0; l < p.k; ++l) {
            std::complex<double> v = ta ? std::complex<double>(a[(size_t)i * p.ld_a + l]) : std::complex<double>(a[(size_t)l * p.ld_a + i]);
            rhs += (ca ? std::conj(v) : v) * bx[(size_t)l];
        }
        worst = std::max(worst, std::abs(lhs - std::complex<double>(p.alpha) * rhs) / ((double)p.k * (double)p.n));
    }
    return worst;
}

template <typename T>
int run_scaling(const cli::Problem& p, bool random, const std::string& counts, long long repetitions) {
    const double flops_per_mul = flops_per_fma<T>::value * (double)p.m * (double)p.n * (double)p.k;
    const size_t na = (size_t)p.ld_a * p.a_cols, nb = (size_t)p.ld_b * p.b_cols, nc = (size_t)p.ld_c * p.n;
    void *pa = nullptr, *pb = nullptr, *pc = nullptr;
    auto t_alloc = std::chrono::steady_clock::now();
    // cudaHostAlloc page-locks at ~2 GB/s whatever the thread count (240 GB: two minutes); tmm_malloc_pinned_large reaches ~26 GB/s
    gpu::check_tmm_status(tmm_malloc_pinned_large(na * sizeof(T), &pa));
    gpu::check_tmm_status(tmm_malloc_pinned_large(nb * sizeof(T), &pb));
    gpu::check_tmm_status(tmm_malloc_pinned_large(nc * sizeof(T), &pc));
    T *a_host = static_cast<T*>(pa), *b_host = static_cast<T*>(pb), *c_host = static_cast<T*>(pc);
    const double alloc_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_alloc).count();
    auto t_fill = std::chrono::steady_clock::now();
    parallel_fill(a_host, na, random, 1);
    parallel_fill(b_host, nb, random, 2);
    const double fill_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_fill).count();
    std::printf("host buffers: %.1f GB pinned in %.1f s, filled in %.1f s\n", (double)(na + nb + nc) * sizeof(T) / 1e9, alloc_s, fill_s);
    auto ctx = gpu::make_context<T>((int)p.n_streams, (int)p.tile_m, (int)p.tile_n, (int)p.tile_k);
    const T alpha = make_scalar<T>(p.alpha), beta = make_scalar<T>(0.0);
    std::vector<std::pair<int, double>> results;
    size_t pos = 0;
    while (pos < counts.size()) {
        const size_t comma = counts.find(',', pos);
        const int g = std::atoi(counts.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos).c_str());
        pos = comma == std::string::npos ? counts.size() : comma + 1;
        if (g < 1) continue;
        parallel_fill(c_host, nc, false, 0);  // (ones: beta = 0 must overwrite them)
        gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), g, nullptr));
        if (p.warm_size > 0) {  // first-use costs (stream / event creation, NCCL and peer mappings, the first cudaMalloc) on a small product
            const long long w = std::min<long long>(p.warm_size, std::min(p.m, std::min(p.n, p.k)));
            gpu::gemm64<T>(*ctx, p.trans_a, p.trans_b, w, w, w, alpha, a_host, p.ld_a, b_host, p.ld_b, beta, c_host, p.ld_c, false, true);
            parallel_fill(c_host, (size_t)p.ld_c * w, false, 0);
        }
        double ms = 1e300;
        for (long long rep = 0; rep < repetitions; ++rep) {  // the first call at a size also grows the device buffers; the best call is reported
            const auto t0 = std::chrono::steady_clock::now();
            gpu::gemm64<T>(*ctx, p.trans_a, p.trans_b, p.m, p.n, p.k, alpha, a_host, p.ld_a, b_host, p.ld_b, beta, c_host, p.ld_c, false, true);
            const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (repetitions > 1) std::printf("  gpus %d call %lld: %.1f ms\n", g, rep, t);
            ms = std::min(ms, t);
        }
        tmm_call_stats st{};
        tmm_context_last_stats(ctx->native(), &st);
        const double defect = linearity_defect(p, a_host, b_host, c_host, 16);
        std::printf("SCALING gpus %d: %.1f ms = %.2f TFLOP/s | H2D %.1f GB, D2H %.1f GB, NVLink %.1f GB, %llu launches, %s | linearity defect %.2e %s\n", g, ms,
                    flops_per_mul / ms * 1e-9, st.h2d_bytes / 1e9, st.d2h_bytes / 1e9, st.peer_bytes / 1e9, (unsigned long long)st.kernel_launches,
                    st.regime == 0 ? "resident" : "streaming", defect, defect <= 1e-14 * std::max(1.0, std::abs(p.alpha)) ? "OK" : "FAIL");
        std::fflush(stdout);
        results.push_back({g, ms});
    }
    if (results.size() > 1)
        for (size_t i = 0; i + 1 < results.size(); ++i)
            std::printf("SPEEDUP %d GPUs over %d: %.2fx\n", results[i].first, results.back().first, results.back().second / results[i].second);
    tmm_free_pinned(a_host); tmm_free_pinned(b_host); tmm_free_pinned(c_host);
    return 0;
}

template <typename T>
int run(const cli::Problem& p, long long repetitions, bool random, const std::string& variants, long long warmup) {
    const double flops_per_mul = flops_per_fma<T>::value * (double)p.m * (double)p.n * (double)p.k;
    const size_t na = (size_t)p.ld_a * p.a_cols, nb = (size_t)p.ld_b * p.b_cols, nc = (size_t)p.ld_c * p.n;
    void *pa = nullptr, *pb = nullptr, *pc = nullptr;
    gpu::check_tmm_status(tmm_malloc_pinned(na * sizeof(T), &pa));
    gpu::check_tmm_status(tmm_malloc_pinned(nb * sizeof(T), &pb));
    gpu::check_tmm_status(tmm_malloc_pinned(nc * sizeof(T), &pc));
    T *a_host = static_cast<T*>(pa), *b_host = static_cast<T*>(pb), *c_host = static_cast<T*>(pc);
    parallel_fill(a_host, na, random, 1);
    parallel_fill(b_host, nb, random, 2);
    std::memset(pc, 0, nc * sizeof(T));
    auto ctx = gpu::make_context<T>((int)p.n_streams, (int)p.tile_m, (int)p.tile_n, (int)p.tile_k);
    const T alpha = make_scalar<T>(p.alpha), beta = make_scalar<T>(p.beta);

    std::cout << "\n==================================================\n"
              << "         Results of benchmarking Tiled-MM    \n"
              << "==================================================" << std::endl;
    for (int variant = 0; variant < 2; ++variant) {
        const bool copy_c_back = variant == 0;
        if ((copy_c_back && variants == "device") || (!copy_c_back && variants == "back")) continue;
        std::cout << (copy_c_back ? " 1) The version with copying C to back to host: " : " 2) The version without copying C to back to host: ") << std::endl;
        if (copy_c_back && p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), (int)p.gpus, nullptr));
        if (!copy_c_back && p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), 1, nullptr));  // device C lives on one GPU
        auto start = std::chrono::steady_clock::now();
        for (long long i = 0; i < repetitions + warmup; ++i) {
            if (i == warmup) start = std::chrono::steady_clock::now();  // the first run warms the context up (examples/multiply.cpp:170-176)
            gpu::gemm64<T>(*ctx, p.trans_a, p.trans_b, p.m, p.n, p.k, alpha, a_host, p.ld_a, b_host, p.ld_b, beta, c_host, p.ld_c,
                           /*pin_host_buffers=*/false, copy_c_back);
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count() / (double)repetitions;
        std::cout << "    -> Avg Time [ms] = " << ms << std::endl;
        std::cout << "    -> Throughput [Gflops] = " << flops_per_mul / (1e-3 * ms) / 1e9 << std::endl;
        tmm_call_stats st;
        if (tmm_context_last_stats(ctx->native(), &st) == TMM_OK)
            std::printf("    -> last call: H2D %.1f MB, D2H %.1f MB, NVLink %.1f MB, %llu kernel launches, %s regime\n", st.h2d_bytes / 1e6, st.d2h_bytes / 1e6,
                        st.peer_bytes / 1e6, (unsigned long long)st.kernel_launches, st.regime == 0 ? "resident" : "streaming");
        std::cout << "==================================================" << std::endl;
    }
    // the reference's apps never release their pinned buffers (util.hpp has no free helper); these do
    tmm_free_pinned(a_host); tmm_free_pinned(b_host); tmm_free_pinned(c_host);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    auto table = cli::gemm_options(true);
    table.push_back({"", "variants", "both", "both | back | device: run the copy-C-back variant, the device-resident-C variant, or both (the reference runs both)."});
    table.push_back({"", "warmup", "1", "Untimed runs before the timed ones (the reference does one); 0 for runs that take minutes."});
    table.push_back({"", "random", "0", "1: fill A and B with uniform(-1,1) values instead of ones (realistic power draw)."});
    table.push_back({"", "scaling", "", "Comma-separated GPU counts, e.g. 8,1: one timed copy-back call per count on the same host buffers, with a result check and the speed-up."});
    table.push_back({"", "warm_size", "2048", "--scaling: size of the small warm-up product run before each timed call (0: none)."});
    cli::Args args(table);
    if (!args.read(argc, argv)) return 2;
    if (args.help_requested) { args.usage("multiply", "Benchmarking Tiled-MM: measures the runtime of the tiled out-of-core GEMM."); return 0; }
    cli::Problem p;
    if (!cli::problem_from(args, &p)) return 0;
    if (p.gpus > 1) setenv("TMM_PINNED_NUMA", "interleave", 0);  // one copy of A, B, C read by the GPUs of both sockets: spread its pages  // the reference also exits 0 after its [ERROR] message (examples/multiply.cpp:83-91)
    const long long repetitions = std::max<long long>(1, args.integer("n_rep"));
    const bool random = args.integer("random") != 0;
    const std::string variants = args.text("variants");
    const long long warmup = std::max<long long>(0, args.integer("warmup"));
    const std::string scaling = args.text("scaling");
    p.warm_size = args.integer("warm_size");
    cli::print_banner(p, repetitions);
    try {
        if (!scaling.empty()) {
            switch (p.type) {
            case 's': return run_scaling<float>(p, random, scaling, repetitions);
            case 'c': return run_scaling<std::complex<float>>(p, random, scaling, repetitions);
            case 'z': return run_scaling<std::complex<double>>(p, random, scaling, repetitions);
            default: return run_scaling<double>(p, random, scaling, repetitions);
            }
        }
        switch (p.type) {
        case 's': return run<float>(p, repetitions, random, variants, warmup);
        case 'c': return run<std::complex<float>>(p, repetitions, random, variants, warmup);
        case 'z': return run<std::complex<double>>(p, repetitions, random, variants, warmup);
        default: return run<double>(p, repetitions, random, variants, warmup);
        }
    } catch (const std::exception& e) {
        std::cerr << "multiply: " << e.what() << std::endl;
        return 1;
    }
}
