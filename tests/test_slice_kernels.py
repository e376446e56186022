"""The operand-slicing kernels of the experimental FP64-emulating DGEMM (csrc/tmm_slice.cuh, TMM_F64_MATH=i8) were written after the round's
GPU budget was spent.  Their source is device code only, so the VERY SAME header is compiled here for the CPU (blockIdx / threadIdx shim, every
"thread" of every "block" in a loop, the launch geometry of gemm_f64_i8.cu's prepare()) and checked against the numpy restatement of
tools/fp64_emulation_study.py; the slices are then pushed through the arithmetic of the GEMM kernel's epilogue (exact integer slice products,
group scales, row / column exponents) to show the bound of the FP64 parity tests is met."""
import ctypes
import importlib.util
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

HARNESS = r'''
#include <cstdint>
#include <cstring>
#include <cmath>
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static dim3 blockIdx, threadIdx, blockDim, gridDim;
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline int atomicMax(int* p, int v) { int old = *p; if (v > old) *p = v; return old; }
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#include "tmm_slice.cuh"
template <typename K, typename... A>
static void launch(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (unsigned tx = 0; tx < block.x; ++tx) { blockIdx = dim3(bx, by); threadIdx = dim3(tx); kernel(args...); }
}
using namespace tmm::f64i8;
extern "C" {
int slice_bits() { return SLICE_BITS; }
int no_data() { return NO_DATA; }
// the launches of gemm_f64_i8.cu prepare(), with the row cap of the slicing grid lowered to exercise its grid-stride loop
void run_prepare(const double* x, long stride_row, long stride_k, int rows, int k, int slices, int* e, int8_t* out, long pitch, long slice_stride, int row_cap) {
    std::memset(e, 0x88, (size_t)rows * sizeof(int));
    if (stride_k == 1) {
        const int gx = (k + 1023) / 1024 < 8 ? (k + 1023) / 1024 : 8;
        launch(row_exponents_kmajor, dim3((unsigned)gx, (unsigned)(rows < row_cap ? rows : row_cap)), dim3(256), x, (int64_t)stride_row, rows, k, e);
    } else {
        const int k_per_block = 128;
        launch(row_exponents, dim3((unsigned)((rows + 255) / 256), (unsigned)((k + k_per_block - 1) / k_per_block)), dim3(256), x, (int64_t)stride_row, (int64_t)stride_k, rows, k,
               k_per_block, e);
    }
    if (stride_row == 1)   // as in prepare(): the coalesced variant for row-contiguous operands (k-group cap lowered with row_cap to exercise its grid-stride loop)
        launch(slice_rows_contiguous, dim3((unsigned)((rows + 255) / 256), (unsigned)(((k + K_PER_THREAD - 1) / K_PER_THREAD) < row_cap ? ((k + K_PER_THREAD - 1) / K_PER_THREAD) : row_cap)), dim3(256), x, (int64_t)stride_k,
               rows, k, (const int*)e, out, (int64_t)pitch, (int64_t)slice_stride, slices);
    else
        launch(slice_rows, dim3((unsigned)((k + 1023) / 1024), (unsigned)(rows < row_cap ? rows : row_cap)), dim3(256), x, (int64_t)stride_row, (int64_t)stride_k, rows, k,
               (const int*)e, out, (int64_t)pitch, (int64_t)slice_stride, slices);
}
}
'''


@pytest.fixture(scope="module")
def kernels(tmp_path_factory):
    d = tmp_path_factory.mktemp("slice")
    (d / "harness.cpp").write_text(HARNESS)
    so = d / "libslice_cpu.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", str(ROOT / "tiled-mm_b200" / "csrc"), str(d / "harness.cpp"), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


@pytest.fixture(scope="module")
def study():
    spec = importlib.util.spec_from_file_location("fp64_emulation_study", ROOT / "tools" / "fp64_emulation_study.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _fp(x):
    return x.ctypes.data_as(ctypes.c_void_p)


def prepare(kernels, stored, rows_contiguous, rows, k, ld, slices, row_cap=32768):
    """stored: flat column-major array; the operand's (row i, k index l) element sits at i + l * ld (rows contiguous) or l + i * ld."""
    pitch = -(-k // 128) * 128
    rows_pad = -(-rows // 128) * 128
    e = np.zeros(rows, np.int32)
    out = np.full(slices * rows_pad * pitch, 99, np.int8)
    sr, sk = (1, ld) if rows_contiguous else (ld, 1)
    kernels.run_prepare(_fp(stored), ctypes.c_long(sr), ctypes.c_long(sk), rows, k, slices, _fp(e), _fp(out), ctypes.c_long(pitch), ctypes.c_long(rows_pad * pitch), row_cap)
    return e, out.reshape(slices, rows_pad, pitch)


@pytest.mark.parametrize("rows_contiguous", [True, False])
@pytest.mark.parametrize("row_cap", [32768, 5])
def test_slices_match_the_restatement(kernels, study, rows_contiguous, row_cap):
    rng = np.random.default_rng(11)
    rows, k, S = 37, 1031, 8
    x = rng.standard_normal((rows, k)) * 10.0 ** rng.uniform(-8, 8, (rows, 1))
    x[3, :] = 0.0                      # an all-zero row
    x[5, 7] = 0.0
    x[9, :] = rng.integers(0, 10, k)   # the reference's test data
    ld = (rows if rows_contiguous else k) + 3
    stored = np.zeros(ld * (k if rows_contiguous else rows))
    if rows_contiguous:
        stored.reshape(k, ld)[:, :rows] = x.T
    else:
        stored.reshape(rows, ld)[:, :k] = x
    e, q = prepare(kernels, stored, rows_contiguous, rows, k, ld, S, row_cap)
    bits = kernels.slice_bits()
    e_np, q_np, p0 = study.slice_rows(x, bits, S)
    zero_rows = ~np.any(x != 0, axis=1)
    assert np.all(e[zero_rows] <= kernels.no_data())
    assert np.array_equal(e[~zero_rows], e_np[~zero_rows].astype(np.int32))
    for s in range(S):
        assert np.array_equal(q[s, :rows, :k].astype(np.int64), q_np[s]), s
        assert np.abs(q[s, :rows, :k].astype(np.int64)).max() <= 2 ** (bits - 1)
        tail = q[s, :rows, k:(k + 15) // 16 * 16 if rows_contiguous else (k + 3) // 4 * 4]
        assert not tail.any(), "the k tail of the last group must be written as zeros"
    # reconstruction: the slices reproduce x to the stated remainder
    ee = np.where(zero_rows, 0, e).astype(np.int64)
    recon = sum(np.ldexp(q[s, :rows, :k].astype(np.float64), -(p0 + bits * s)) for s in range(S))
    rem = np.abs(np.ldexp(x, -ee[:, None]) - recon)
    assert rem.max() <= 2.0 ** -(p0 + bits * (S - 1) + 1)
    assert not np.abs(q[1:, 9, :k]).any(), "small integers live entirely in slice 0"


@pytest.mark.parametrize("rows_contiguous", [True, False])
def test_non_finite_rows_are_marked_and_sliced_as_zero(kernels, rows_contiguous):
    rng = np.random.default_rng(13)
    rows, k, S = 9, 300, 4
    x = rng.standard_normal((rows, k))
    x[2, 17] = np.inf; x[4, 0] = np.nan; x[6, k - 1] = -np.inf
    ld = rows if rows_contiguous else k
    stored = np.ascontiguousarray(x.T if rows_contiguous else x).reshape(-1)
    e, q = prepare(kernels, stored, rows_contiguous, rows, k, ld, S)
    bad = np.array([2, 4, 6])
    assert np.all(e[bad] >= 5000) and np.all(e[np.setdiff1d(np.arange(rows), bad)] < 100)
    assert not q[:, bad, :k].any()
    assert q[0, 0, :k].any()


@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")])
def test_emulated_product_meets_the_fp64_parity_bound(kernels, ta, tb):
    """Slices from the kernels (all four stored orientations) -> exact integer slice products -> the epilogue's arithmetic, in its order."""
    rng = np.random.default_rng(12)
    m, n, k, S = 70, 45, 900, 8
    a = rng.random((m, k)) * 2 - 1
    b = rng.random((k, n)) * 2 - 1
    for ints in (False, True):
        if ints:
            a, b = rng.integers(0, 10, (m, k)).astype(np.float64), rng.integers(0, 10, (k, n)).astype(np.float64)
        lda = (m if ta == "N" else k) + 1
        ldb = (k if tb == "N" else n) + 2
        sa = np.zeros(lda * (k if ta == "N" else m)); sb = np.zeros(ldb * (n if tb == "N" else k))
        if ta == "N": sa.reshape(k, lda)[:, :m] = a.T
        else: sa.reshape(m, lda)[:, :k] = a
        if tb == "N": sb.reshape(n, ldb)[:, :k] = b.T
        else: sb.reshape(k, ldb)[:, :n] = b
        ea, qa = prepare(kernels, sa, ta == "N", m, k, lda, S)           # op(A) row i: N -> rows contiguous
        eb, qb = prepare(kernels, sb, tb != "N", n, k, ldb, S)           # op(B) column j: T -> "rows" (columns) contiguous
        bits, p0 = kernels.slice_bits(), kernels.slice_bits() - 1
        total = np.zeros((m, n))
        for g in range(S - 1, -1, -1):
            acc = np.zeros((m, n), np.int64)
            for s in range(g + 1):
                acc += qa[s, :m, :k].astype(np.int64) @ qb[g - s, :n, :k].astype(np.int64).T
            assert np.abs(acc).max() < 2 ** 31, "one int32 window per group must not overflow at this k"
            total += acc.astype(np.float64) * 2.0 ** -(2 * p0 + bits * g)
        c = np.ldexp(total, ea.astype(np.int64)[:, None] + eb.astype(np.int64)[None, :])
        ref = (a.astype(np.longdouble) @ b.astype(np.longdouble))
        err = float(np.max(np.abs(c.astype(np.longdouble) - ref))) / (k * np.abs(a).max() * np.abs(b).max())
        if ints:
            assert np.array_equal(c, a @ b)
        else:
            assert err <= 1e-16, err   # the parity tests allow 1e-15
        # the CTA-pair version (igemm_group_kernel) sums over g in GLOBAL memory: one launch per group, C = beta C + alpha 2^(..) acc first,
        # then S - 1 read-modify-writes, each rounded to FP64
        alpha, beta = 1.5, 0.25
        c0 = rng.random((m, n)) - 0.5 if not ints else rng.integers(0, 10, (m, n)).astype(np.float64)
        cg = c0.copy()
        exps = ea.astype(np.int64)[:, None] + eb.astype(np.int64)[None, :]
        for g in range(S - 1, -1, -1):
            acc = np.zeros((m, n), np.int64)
            for s in range(g + 1):
                acc += qa[s, :m, :k].astype(np.int64) @ qb[g - s, :n, :k].astype(np.int64).T
            add = alpha * np.ldexp(acc.astype(np.float64), exps - (2 * p0 + bits * g))
            cg = add + beta * cg if g == S - 1 else cg + add
        ref2 = alpha * ref + beta * c0.astype(np.longdouble)
        err2 = float(np.max(np.abs(cg.astype(np.longdouble) - ref2))) / (k * np.abs(a).max() * np.abs(b).max())
        if ints:
            assert np.array_equal(cg, alpha * (a @ b) + beta * c0)
        else:
            assert err2 <= 2e-16, err2
