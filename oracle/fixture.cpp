// ORACLE — TEST INFRASTRUCTURE ONLY.
// Restates the input generator of the reference's only test (tests/test-multiply.cpp:58-66,269-274):
// one process-wide std::mt19937 seeded with 42 feeding std::uniform_int_distribution<int>(0, 9),
// shared by consecutive fills in the order A, B, C.  Every value is a small integer, so any correct
// FP64 GEMM is bit-identical on it regardless of summation order (|partial sums| <= 81 k + 9 << 2^53).
// The sequence depends on libstdc++'s uniform_int_distribution algorithm (g++ 13 here), which is why
// this is C++ over <random> and not a hand-rolled generator; oracle_fixture_lemire() below is the
// explicit restatement of that algorithm, cross-checked against <random> in tests/test_oracle.py.
#include <cstddef>
#include <cstdint>
#include <random>

namespace {
struct Gen {
    std::mt19937 rng{42};
    std::uniform_int_distribution<int> dist{0, 9};
};
Gen g_gen;
}

extern "C" {

void oracle_fixture_reset(unsigned seed) { g_gen = Gen{}; g_gen.rng.seed(seed); }

// dtype: 0 f32, 1 f64, 2 c32, 3 c64.  For complex types the imaginary part is 0, exactly what
// static_cast<std::complex<T>>(int) yields.
void oracle_fixture_fill(int dtype, void* ptr, size_t count) {
    for (size_t i = 0; i < count; ++i) {
        int v = g_gen.dist(g_gen.rng);
        switch (dtype) {
        case 0: static_cast<float*>(ptr)[i] = static_cast<float>(v); break;
        case 1: static_cast<double*>(ptr)[i] = static_cast<double>(v); break;
        case 2: static_cast<float*>(ptr)[2 * i] = static_cast<float>(v); static_cast<float*>(ptr)[2 * i + 1] = 0.f; break;
        default: static_cast<double*>(ptr)[2 * i] = static_cast<double>(v); static_cast<double*>(ptr)[2 * i + 1] = 0.0; break;
        }
    }
}

// Explicit restatement: mt19937 output -> Lemire's nearly-divisionless bounded integer in [0, range),
// the path libstdc++ takes for a 32-bit URBG (bits/uniform_int_dist.h, _S_nd).
void oracle_fixture_lemire(unsigned seed, uint32_t range, int32_t* out, size_t count) {
    std::mt19937 rng(seed);
    for (size_t i = 0; i < count; ++i) {
        uint64_t product = uint64_t(rng()) * uint64_t(range);
        uint32_t low = uint32_t(product);
        if (low < range) {
            uint32_t threshold = -range % range;
            while (low < threshold) {
                product = uint64_t(rng()) * uint64_t(range);
                low = uint32_t(product);
            }
        }
        out[i] = int32_t(product >> 32);
    }
}
}
