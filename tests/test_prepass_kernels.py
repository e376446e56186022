"""The elementwise operand-preparation kernels of the tcgen05 CGEMM embedding and of the BF16 entry point (csrc/tmm_prepass.cuh) were
written after the round's GPU budget was spent.  Their source is device code only, so here the VERY SAME header is compiled for the
CPU - blockIdx / threadIdx / float2 come from a small shim, every "thread" of every "block" runs in a loop, with the launch geometry
of the real launchers - and the output is compared with the numpy restatement that tests/test_c32_embedding.py proves correct."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from test_c32_embedding import embed_a, embed_b

ROOT = Path(__file__).resolve().parent.parent

HARNESS = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static dim3 blockIdx, threadIdx, blockDim, gridDim;
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#include "tmm_prepass.cuh"
template <typename K, typename... A>
static void launch(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (unsigned tx = 0; tx < block.x; ++tx) { blockIdx = dim3(bx, by); threadIdx = dim3(tx); kernel(args...); }
}
// the launch geometry of gemm_c32_tc.cu / gemm_bf16_tc.cu, with the column cap lowered so that the grid-stride loop is exercised
static dim3 pass_grid(int contiguous, int columns, int cap) { return dim3((unsigned)((contiguous + 255) / 256), (unsigned)(columns < 1 ? 1 : (columns > cap ? cap : columns))); }
using namespace tmm::c32tc;
extern "C" {
void run_embed_a_n(const float* a, long lda, int m, int k, float ar, float ai, float* out, long pitch, int cap) {
    launch(embed_a_n, pass_grid(m, k, cap), dim3(256), (const float2*)a, (int64_t)lda, m, k, make_float2(ar, ai), (float2*)out, (int64_t)(pitch / 2));
}
void run_embed_a_t(const float* a, long lda, int k, int m, float ar, float ai, int conj, float* out, long pitch, int cap) {
    launch(embed_a_t, pass_grid(k, m, cap), dim3(256), (const float2*)a, (int64_t)lda, k, m, make_float2(ar, ai), conj, (float2*)out, (int64_t)(pitch / 2));
}
void run_split_b_t(const float* b, long ldb, int n, int k, int conj, float* out, long pitch, int cap) {
    launch(split_b_t, pass_grid(n, k, cap), dim3(256), (const float2*)b, (int64_t)ldb, n, k, conj, out, (int64_t)pitch);
}
void run_split(const float* x, long count, float* hi, float* lo, float* lo_trunc) {
    for (long i = 0; i < count; ++i) { tmm::f32tc::split_tf32(x[i], hi[i], lo[i]); lo_trunc[i] = tmm::f32tc::lo_of_truncated(x[i]); }
}
// the split as the kernel's split warps apply it: 16-byte chunks of four elements (count is a multiple of 4)
void run_split_x4(const float* x, long count, float* hi, float* lo) {
    for (long i = 0; i + 4 <= count; i += 4) {
        const float v[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]};
        float h[4], l[4];
        tmm::f32tc::split_tf32_x4(v, h, l);
        for (int j = 0; j < 4; ++j) { hi[i + j] = h[j]; lo[i + j] = l[j]; }
    }
}
// the A-split reads of sgemm_tc_ts_kernel, thread by thread: lane-quarter q, lane l -> row 32q + l, two halves of 16 k-values
void run_widen(const uint16_t* in, long ld, int rows, int cols, float* out, long pitch, int cap) {
    launch(widen, pass_grid(rows, cols, cap), dim3(256), in, (int64_t)ld, rows, cols, out, (int64_t)pitch);
}
}
'''


@pytest.fixture(scope="module")
def kernels(tmp_path_factory):
    d = tmp_path_factory.mktemp("prepass")
    (d / "harness.cpp").write_text(HARNESS)
    so = d / "libprepass_cpu.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", str(ROOT / "tiled-mm_b200" / "csrc"), str(d / "harness.cpp"), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def _fp(x):
    return x.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("ta", ["N", "T", "C"])
@pytest.mark.parametrize("cap", [32768, 3])
def test_embed_a_kernels_match_the_restatement(kernels, ta, cap):
    rng = np.random.default_rng(3)
    m, k = 300, 77                      # more rows than one 256-thread block
    ar, ac = (m, k) if ta == "N" else (k, m)
    lda = ar + 5
    a = (rng.integers(-9, 10, lda * ac) + 1j * rng.integers(-9, 10, lda * ac)).astype(np.complex64)
    alpha = np.complex64(2 - 1j)
    want, _, pitch = embed_a(ta, a, lda, m, k, alpha)
    out = np.zeros_like(want)
    if ta == "N":
        kernels.run_embed_a_n(_fp(a), ctypes.c_long(lda), m, k, ctypes.c_float(alpha.real), ctypes.c_float(alpha.imag), _fp(out), ctypes.c_long(pitch), cap)
    else:
        kernels.run_embed_a_t(_fp(a), ctypes.c_long(lda), k, m, ctypes.c_float(alpha.real), ctypes.c_float(alpha.imag), 1 if ta == "C" else 0, _fp(out),
                              ctypes.c_long(pitch), cap)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("tb", ["T", "C"])
def test_split_b_kernel_matches_the_restatement(kernels, tb):
    rng = np.random.default_rng(4)
    n, k = 290, 61
    ldb = n + 3
    b = (rng.integers(-9, 10, ldb * k) + 1j * rng.integers(-9, 10, ldb * k)).astype(np.complex64)
    want, _ = embed_b(tb, b, ldb, k, n)
    out = np.zeros_like(want)
    kernels.run_split_b_t(_fp(b), ctypes.c_long(ldb), n, k, 1 if tb == "C" else 0, _fp(out), ctypes.c_long(want.shape[1]), 7)
    assert np.array_equal(out, want)


def test_widen_kernel(kernels):
    rng = np.random.default_rng(5)
    rows, cols, ld = 513, 40, 520
    x = (rng.random(ld * cols).astype(np.float32) - 0.5) * np.float32(1e3)
    x[::97] = np.float32(np.inf); x[5] = np.float32(-0.0)
    bf = (x.view(np.uint32) >> 16).astype(np.uint16)
    pitch = -(-rows // 32) * 32
    out = np.full((cols, pitch), np.float32(7), np.float32)
    kernels.run_widen(_fp(bf), ctypes.c_long(ld), rows, cols, _fp(out), ctypes.c_long(pitch), 9)
    want = (bf.astype(np.uint32) << 16).view(np.float32).reshape(cols, ld)[:, :rows]
    assert np.array_equal(out[:, :rows].view(np.uint32), want.view(np.uint32))
    assert np.all(out[:, rows:] == 7), "padding of the widened panel must stay untouched"


def test_tf32_operand_splits(kernels):
    """split_tf32 (the hardware-validated round-to-nearest split of the FP32-accurate SGEMM) and lo_of_truncated (the experimental
    raw-bits variant): hi and lo are TF32 numbers, hi + lo reproduces x to 2^-21 |x| or better, special values travel in hi alone."""
    rng = np.random.default_rng(6)
    x = np.concatenate([(rng.standard_normal(20000) * 10.0 ** rng.integers(-30, 30, 20000)).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 3.0, 1e-45, -1e-40, 3.4028235e38, np.inf, -np.inf, np.nan, 9.0, 1 + 2 ** -11, 1 + 2 ** -12], np.float32)])
    hi, lo, lot = (np.zeros_like(x) for _ in range(3))
    kernels.run_split(_fp(x), ctypes.c_long(x.size), _fp(hi), _fp(lo), _fp(lot))
    low13 = np.uint32(0x1FFF)
    assert np.all((hi.view(np.uint32) & low13) == 0) and np.all((lo.view(np.uint32) & low13) == 0) and np.all((lot.view(np.uint32) & low13) == 0)
    fin = np.isfinite(x) & (np.abs(x) < 1e38) & (np.abs(x) > 1e-30)
    xd = x[fin].astype(np.float64)
    assert np.max(np.abs(hi[fin].astype(np.float64) + lo[fin] - xd) / np.abs(xd)) <= 2.0 ** -21      # round-to-nearest split: ~2^-23
    trunc = (x.view(np.uint32) & ~low13).view(np.float32)
    assert np.max(np.abs(trunc[fin].astype(np.float64) + lot[fin] - xd) / np.abs(xd)) <= 2.0 ** -20  # raw-bits split: one bit less
    ints = np.arange(-2048, 2049, dtype=np.float32)                                                   # small integers are exact in TF32: lo = 0
    h2, l2, lt2 = (np.zeros_like(ints) for _ in range(3))
    kernels.run_split(_fp(ints), ctypes.c_long(ints.size), _fp(h2), _fp(l2), _fp(lt2))
    assert np.array_equal(h2, ints) and not l2.any() and not lt2.any()
    special = ~np.isfinite(x)
    assert np.all(lo[special] == 0) and np.all(lot[special] == 0) and np.array_equal(np.isnan(hi[special]), np.isnan(x[special]))
    assert np.array_equal(hi[np.isinf(x)], x[np.isinf(x)])
    # the chunked form the kernel uses (short path for chunks of plain values, split_tf32 for a chunk with an Inf / NaN / huge value): identical
    n4 = x.size // 4 * 4
    hi4, lo4 = np.zeros(n4, np.float32), np.zeros(n4, np.float32)
    kernels.run_split_x4(_fp(x), ctypes.c_long(n4), _fp(hi4), _fp(lo4))
    assert np.array_equal(hi4.view(np.uint32), hi[:n4].view(np.uint32)) and np.array_equal(lo4.view(np.uint32), lo[:n4].view(np.uint32))
    near_max = np.isfinite(x) & (np.abs(x) > 3.4e38)                                                  # finite inputs never turn into Inf: hi is clamped
    assert near_max.any() and np.all(np.isfinite(hi[near_max])) and np.all(np.abs(hi[near_max].astype(np.float64) + lo[near_max] - x[near_max]) <= 2.0 ** -21 * np.abs(x[near_max].astype(np.float64)))




def test_chunked_split_equals_the_per_element_split_on_random_bit_patterns(kernels):
    """split_tf32_x4 (what the split warps run: the short path for chunks of plain values, split_tf32 for a chunk holding an Inf / NaN / value of
    2^127 or more) against split_tf32 element by element, over two million uniformly random 32-bit patterns - every exponent, denormals, both
    zeros, NaN payloads - and over chunks that mix one special value with plain ones: bit-identical hi and lo."""
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 1 << 32, 1 << 21, dtype=np.uint64).astype(np.uint32)
    bits[::64] = rng.choice(np.array([0x7F800000, 0xFF800000, 0x7FC00000, 0x7F7FFFFF, 0xFF7FFFFF, 0x7F000000, 0x7EFFFFFF, 0x00000001, 0x80000000], np.uint32), bits[::64].size)
    x = bits.view(np.float32)
    hi, lo, lot, hi4, lo4 = (np.zeros_like(x) for _ in range(5))
    kernels.run_split(_fp(x), ctypes.c_long(x.size), _fp(hi), _fp(lo), _fp(lot))
    kernels.run_split_x4(_fp(x), ctypes.c_long(x.size), _fp(hi4), _fp(lo4))
    assert np.array_equal(hi4.view(np.uint32), hi.view(np.uint32)) and np.array_equal(lo4.view(np.uint32), lo.view(np.uint32))
    fin = np.isfinite(x)
    assert np.all(np.isfinite(hi[fin])) and np.all(np.isfinite(lo[fin])), "a finite operand must never split into a non-finite part"
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0) and np.all((lo.view(np.uint32) & np.uint32(0x1FFF)) == 0)
