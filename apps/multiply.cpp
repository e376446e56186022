// multiply - the miniapp of the reference (examples/multiply.cpp) on tiled_mm_b200: same flags, same printed report
// ("-> Avg Time [ms]", "-> Throughput [Gflops]" for the copy-back and the device-resident variant), so scripts such as the
// reference's examples/compare.sh keep working.  Additions: --type s|d|c|z, -t accepts C, --gpus N spreads the C blocks
// over N GPUs of the box (copy-back variant), times are fractional milliseconds (the reference's integer division
// reports 0 ms / inf Gflops for small sizes), flop counts are 64-bit doubles.
#include <Tiled-MM/tiled_mm.hpp>

#include "cli.hpp"

#include <chrono>
#include <complex>
#include <cstdio>

namespace {

template <typename T> T make_scalar(double v) { return T(v); }
template <typename T> struct flops_per_fma { static constexpr double value = 2.0; };
template <typename R> struct flops_per_fma<std::complex<R>> { static constexpr double value = 8.0; };

template <typename T>
int run(const cli::Problem& p, long long repetitions) {
    const double flops_per_mul = flops_per_fma<T>::value * (double)p.m * (double)p.n * (double)p.k;
    T* a_host = gpu::malloc_pinned<T>((size_t)p.ld_a * p.a_cols, T(1));
    T* b_host = gpu::malloc_pinned<T>((size_t)p.ld_b * p.b_cols, T(1));
    T* c_host = gpu::malloc_pinned<T>((size_t)p.ld_c * p.n, T(0));
    auto ctx = gpu::make_context<T>((int)p.n_streams, (int)p.tile_m, (int)p.tile_n, (int)p.tile_k);
    const T alpha = make_scalar<T>(p.alpha), beta = make_scalar<T>(p.beta);

    std::cout << "\n==================================================\n"
              << "         Results of benchmarking Tiled-MM    \n"
              << "==================================================" << std::endl;
    for (int variant = 0; variant < 2; ++variant) {
        const bool copy_c_back = variant == 0;
        std::cout << (copy_c_back ? " 1) The version with copying C to back to host: " : " 2) The version without copying C to back to host: ") << std::endl;
        if (copy_c_back && p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), (int)p.gpus, nullptr));
        if (!copy_c_back && p.gpus > 1) gpu::check_tmm_status(tmm_context_set_devices(ctx->native(), 1, nullptr));  // device C lives on one GPU
        auto start = std::chrono::steady_clock::now();
        for (long long i = 0; i < repetitions + 1; ++i) {
            if (i == 1) start = std::chrono::steady_clock::now();  // run 0 warms the context up (examples/multiply.cpp:170-176)
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
    cli::Args args(cli::gemm_options(true));
    if (!args.read(argc, argv)) return 2;
    if (args.help_requested) { args.usage("multiply", "Benchmarking Tiled-MM: measures the runtime of the tiled out-of-core GEMM."); return 0; }
    cli::Problem p;
    if (!cli::problem_from(args, &p)) return 0;  // the reference also exits 0 after its [ERROR] message (examples/multiply.cpp:83-91)
    const long long repetitions = std::max<long long>(1, args.integer("n_rep"));
    cli::print_banner(p, repetitions);
    try {
        switch (p.type) {
        case 's': return run<float>(p, repetitions);
        case 'c': return run<std::complex<float>>(p, repetitions);
        case 'z': return run<std::complex<double>>(p, repetitions);
        default: return run<double>(p, repetitions);
        }
    } catch (const std::exception& e) {
        std::cerr << "multiply: " << e.what() << std::endl;
        return 1;
    }
}
