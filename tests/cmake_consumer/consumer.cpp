// Reference-style caller (README.md:60-103 usage): pinned buffers, make_context, gemm with and without copying C back.
#include <Tiled-MM/tiled_mm.hpp>
#include <Tiled-MM/util.hpp>

#include <cstdio>

#ifndef TILED_MM_CUDA
#error "the package must define TILED_MM_CUDA like the reference's target does"
#endif

int main(int argc, char**) {
    if (argc > 1) {  // only runs on a GPU box; the CPU test just builds and links this
        const int m = 300, n = 200, k = 100;
        double* a = gpu::malloc_pinned<double>(size_t(m) * k, 1.0);
        double* b = gpu::malloc_pinned<double>(size_t(k) * n, 2.0);
        double* c = gpu::malloc_pinned<double>(size_t(m) * n, 0.0);
        auto ctx = gpu::make_context<double>();
        gpu::gemm(*ctx, 'N', 'N', m, n, k, 1.0, a, m, b, k, 0.0, c, m, false, true);
        std::printf("%g\n", c[0]);
        return c[0] == 2.0 * k ? 0 : 1;
    }
    return 0;
}
