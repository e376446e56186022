// cublasXt comparator (GPU box; development tool, NOT part of the product - it exists because the reference's headline figure compares
// Tiled-MM with cublasXt, README.md:20-25 and examples/cublasXt-multiply.cpp).  Same flags as bin/multiply; "-> Avg Time [ms]" /
// "-> Throughput [Gflops]" report lines; --block sets cublasXtSetBlockDim (the reference's "tuned" variant uses 4000 + pinned buffers).
#include "../apps/cli.hpp"

#include <cublasXt.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>

int main(int argc, char** argv) {
    auto table = cli::gemm_options(true);
    table.push_back({"", "block", "4000", "cublasXt block dimension (cublasXtSetBlockDim)."});
    cli::Args args(table);
    if (!args.read(argc, argv)) return 2;
    if (args.help_requested) { args.usage("cublasxt-multiply", "Benchmarking cublasXt dgemm on host pointers."); return 0; }
    cli::Problem p;
    if (!cli::problem_from(args, &p)) return 0;
    if (p.type != 'd') { std::cout << "[ERROR]: the comparator runs double only" << std::endl; return 0; }
    const long long reps = std::max<long long>(1, args.integer("n_rep"));
    cli::print_banner(p, reps);
    double *a, *b, *c;
    if (cudaHostAlloc((void**)&a, sizeof(double) * p.ld_a * p.a_cols, 0) != cudaSuccess || cudaHostAlloc((void**)&b, sizeof(double) * p.ld_b * p.b_cols, 0) != cudaSuccess ||
        cudaHostAlloc((void**)&c, sizeof(double) * p.ld_c * p.n, 0) != cudaSuccess) { std::cerr << "cudaHostAlloc failed" << std::endl; return 1; }
    for (long long i = 0; i < p.ld_a * p.a_cols; ++i) a[i] = 1.0;
    for (long long i = 0; i < p.ld_b * p.b_cols; ++i) b[i] = 1.0;
    for (long long i = 0; i < p.ld_c * p.n; ++i) c[i] = 0.0;
    cublasXtHandle_t h;
    if (cublasXtCreate(&h) != CUBLAS_STATUS_SUCCESS) { std::cerr << "cublasXtCreate failed" << std::endl; return 1; }
    std::vector<int> devices((size_t)p.gpus);
    for (int i = 0; i < (int)p.gpus; ++i) devices[i] = i;
    cublasXtDeviceSelect(h, (int)p.gpus, devices.data());
    cublasXtSetBlockDim(h, (int)args.integer("block"));
    auto op = [](char t) { return t == 'N' ? CUBLAS_OP_N : (t == 'T' ? CUBLAS_OP_T : CUBLAS_OP_C); };
    auto start = std::chrono::steady_clock::now();
    for (long long i = 0; i < reps + 1; ++i) {
        if (i == 1) start = std::chrono::steady_clock::now();
        cublasStatus_t st = cublasXtDgemm(h, op(p.trans_a), op(p.trans_b), (size_t)p.m, (size_t)p.n, (size_t)p.k, &p.alpha, a, (size_t)p.ld_a, b, (size_t)p.ld_b, &p.beta, c,
                                          (size_t)p.ld_c);
        if (st != CUBLAS_STATUS_SUCCESS) { std::cerr << "cublasXtDgemm failed: " << (int)st << std::endl; return 1; }
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count() / (double)reps;
    std::cout << "==================================================\n         Results of benchmarking cublasXt    \n==================================================\n"
              << "    -> Avg Time [ms] = " << ms << "\n    -> Throughput [Gflops] = " << 2.0 * p.m * p.n * p.k / (1e-3 * ms) / 1e9
              << "\n==================================================" << std::endl;
    cublasXtDestroy(h);
    return 0;
}
