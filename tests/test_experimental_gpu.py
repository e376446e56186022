"""GPU tests of the paths that were written after round 1's GPU budget was spent.  Round 2 ran them first thing (profiles/r2_experimental_first_run.txt,
r2_f64_i8_first_run.txt): all passed, so they are regular GPU tests now.  What became of the paths: the tcgen05 CGEMM embedding is the default
complex<float> kernel; BF16 inputs and device-pointer operands are additive entry points; the FP64 emulation on the integer tensor cores stays an
opt-in math mode (TMM_F64_MATH=i8[:S]; north_star names DMMA for FP64).  The SGEMM variants with A through tensor memory / CTA pairs and the
CTA-pair int8 kernel were slower than the kernels they were meant to replace and were removed (profiles/r2_tc_variants.txt)."""
import itertools
import os

import numpy as np
import pytest

from test_gemm_gpu import run_case

pytestmark = [pytest.mark.gpu]
ALL_TT = ["".join(p) for p in itertools.product("NTC", "NTC")]


@pytest.fixture(params=["tc", "simt"])
def c32_tc(gpu_tmm, request):
    """both complex<float> kernels: the tcgen05 embedding (default) and the complex-FMA kernel"""
    gpu_tmm.set_c32_math(gpu_tmm.CMATH_TC if request.param == "tc" else gpu_tmm.CMATH_SIMT)
    yield gpu_tmm
    gpu_tmm.set_c32_math(gpu_tmm.CMATH_TC)


@pytest.mark.parametrize("tt", ALL_TT)
def test_cgemm_tensor_core_embedding_exact_on_integers(c32_tc, oracle, tt):
    run_case(c32_tc, oracle, np.complex64, tt, 130, 67, 95, 1 - 2j, 2 + 1j, pad=(1, 2, 3), ints=True)


@pytest.mark.parametrize("tt", ["NN", "TN", "NC", "CT"])
def test_cgemm_tensor_core_embedding_random(c32_tc, oracle, tt):
    run_case(c32_tc, oracle, np.complex64, tt, 777, 530, 1111, 1.5 - 0.5j, 0.25 + 0.75j, pad=(3, 0, 9), tiles=(256, 300, 500))
    run_case(c32_tc, oracle, np.complex64, tt, 777, 530, 1111, 1.0, 0.0, pad=(3, 0, 9))            # beta = 0: C never read


def test_cgemm_tensor_core_device_boundary_odd_ld(c32_tc, oracle):
    """blas_api::cgemm with an odd ldb (no zero-copy view of B) and a base at 8 mod 16."""
    tmm = c32_tc
    m, n, k = 200, 150, 90
    rng = np.random.default_rng(3)
    def gen(count):
        return (rng.integers(0, 10, count) + 1j * rng.integers(0, 10, count)).astype(np.complex64)
    a0, b0, c0 = gen(203 * k + 1), gen(91 * n + 1), gen(m * n)
    da, db, dc = (tmm.malloc_device(x.nbytes) for x in (a0, b0, c0))
    tmm.copy_to_device(a0, da); tmm.copy_to_device(b0, db)
    for oa, ob in ((0, 0), (1, 1)):
        expect = oracle.gemm("N", "N", m, n, k, 1 + 1j, a0[oa:], 203, b0[ob:], 91, 2.0, c0.copy(), m)
        tmm.copy_to_device(c0, dc)
        tmm.device_gemm(np.complex64, "N", "N", m, n, k, 1 + 1j, da + 8 * oa, 203, db + 8 * ob, 91, 2.0, dc, m)
        out = np.empty_like(c0)
        tmm.copy_to_host(dc, out)
        assert np.array_equal(out, expect), (oa, ob)
    for p in (da, db, dc):
        tmm.free_device(p)


@pytest.mark.parametrize("tt", ["NN", "TN", "NT", "TT"])
def test_bf16_input_gemm(gpu_tmm, oracle, tt):
    """tmm_device_gemm_bf16: bf16 operands widened on the device and multiplied on the tcgen05 TF32 path (exact products, FP32 accumulation)."""
    tmm = gpu_tmm
    ta, tb = tt
    m, n, k = 515, 260, 777
    rng = np.random.default_rng(21)
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb = ar + 3, br + 5
    to_bf16 = lambda x: (x.view(np.uint32) >> 16).astype(np.uint16)
    for ints in (True, False):
        gen = (lambda c: rng.integers(-8, 9, c).astype(np.float32)) if ints else (lambda c: (rng.random(c).astype(np.float32) - 0.5))
        a16, b16 = to_bf16(gen(lda * ac)), to_bf16(gen(ldb * bc))
        af, bf = (a16.astype(np.uint32) << 16).view(np.float32), (b16.astype(np.uint32) << 16).view(np.float32)   # the values the device sees
        c0 = gen(m * n)
        want = oracle.gemm(ta, tb, m, n, k, np.float32(1.5), af, lda, bf, ldb, np.float32(0.5), c0.copy(), m, wide=True)
        da, db, dc = tmm.malloc_device(a16.nbytes), tmm.malloc_device(b16.nbytes), tmm.malloc_device(c0.nbytes)
        tmm.copy_to_device(a16, da); tmm.copy_to_device(b16, db); tmm.copy_to_device(c0, dc)
        tmm.device_gemm_bf16(ta, tb, m, n, k, 1.5, da, lda, db, ldb, 0.5, dc, m)
        got = np.empty_like(c0); tmm.copy_to_host(dc, got)
        for p in (da, db, dc):
            tmm.free_device(p)
        if ints:
            assert np.array_equal(got, want), tt
        else:
            assert float(np.max(np.abs(got - want))) / (k * 0.25) <= 2e-6, tt


def test_bf16_native_kind_f16_tn(gpu_tmm, oracle, monkeypatch):
    """TMM_BF16_NATIVE=1 (read per call): k-contiguous bf16 operands straight through kind::f16 MMAs, no widening pass"""
    monkeypatch.setenv("TMM_BF16_NATIVE", "1")
    tmm = gpu_tmm
    to_bf16 = lambda x: (x.view(np.uint32) >> 16).astype(np.uint16)
    rng = np.random.default_rng(22)
    for (m, n, k, ints) in [(128, 128, 64, True), (515, 260, 777, True), (1000, 900, 4100, False)]:
        lda = ldb = -(-k // 8) * 8 + 8                    # 16-byte pitch: the TMA contract of the native path
        gen = (lambda c: rng.integers(-8, 9, c).astype(np.float32)) if ints else (lambda c: (rng.random(c).astype(np.float32) - 0.5))
        a16, b16 = to_bf16(gen(lda * m)), to_bf16(gen(ldb * n))
        af, bf = (a16.astype(np.uint32) << 16).view(np.float32), (b16.astype(np.uint32) << 16).view(np.float32)
        c0 = gen(m * n)
        want = oracle.gemm("T", "N", m, n, k, np.float32(1.0), af, lda, bf, ldb, np.float32(1.0), c0.copy(), m, wide=True)
        da, db, dc = tmm.malloc_device(a16.nbytes), tmm.malloc_device(b16.nbytes), tmm.malloc_device(c0.nbytes)
        tmm.copy_to_device(a16, da); tmm.copy_to_device(b16, db); tmm.copy_to_device(c0, dc)
        tmm.device_gemm_bf16("T", "N", m, n, k, 1.0, da, lda, db, ldb, 1.0, dc, m)
        got = np.empty_like(c0); tmm.copy_to_host(dc, got)
        for p in (da, db, dc):
            tmm.free_device(p)
        if ints:
            assert np.array_equal(got, want), (m, n, k)
        else:
            assert float(np.max(np.abs(got - want))) / (k * 0.25) <= 2e-6, (m, n, k)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_device_pointer_operands(gpu_tmm, oracle, dtype):
    """tmm_gemm with DEVICE pointers (additive): all operands on the device -> one launch on them (any ld), result in place or in the
    context's device C; a mix of host and device operands -> the scheduler with direction-inferring copies."""
    tmm = gpu_tmm
    cplx = np.dtype(dtype).kind == "c"
    alpha, beta = (1 - 2j, 1j) if cplx else (2.0, -1.0)
    rng = np.random.default_rng(5)
    m, n, k, lda, ldb, ldc = 1030, 995, 1170, 1033, 1171, 1031
    def gen(count):
        v = rng.integers(0, 10, count).astype(np.float64)
        return (v + 1j * rng.integers(0, 10, count)).astype(dtype) if cplx else v.astype(dtype)
    a0, b0, c0 = gen(lda * k), gen(ldb * n), gen(ldc * n)
    expect = oracle.gemm("N", "N", m, n, k, alpha, a0, lda, b0, ldb, beta, c0.copy(), ldc)
    da, db, dc = tmm.malloc_device(a0.nbytes), tmm.malloc_device(b0.nbytes), tmm.malloc_device(c0.nbytes)
    tmm.copy_to_device(a0, da); tmm.copy_to_device(b0, db); tmm.copy_to_device(c0, dc)
    with tmm.make_context(dtype) as ctx:
        tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc, pin_host_buffers=True, copy_c_back=True)
        assert ctx.last_stats().h2d_bytes == 0
        out = np.empty_like(c0); tmm.copy_to_host(dc, out)
        assert np.array_equal(out, expect)
        tmm.copy_to_device(c0, dc)
        tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc, pin_host_buffers=False, copy_c_back=False)
        out2 = np.empty(m * n, dtype=dtype); tmm.copy_to_host(ctx.get_full_device_buffer_c().data(), out2, m * n)
        assert np.array_equal(out2.reshape(n, m), expect.reshape(n, ldc)[:, :m])
        ch = tmm.malloc_pinned(dtype, c0.size); ch[:] = c0
        bh = tmm.malloc_pinned(dtype, b0.size); bh[:] = b0
        tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, bh, ldb, beta, ch, ldc, pin_host_buffers=False, copy_c_back=True)
        assert np.array_equal(np.asarray(ch), expect)
    for p in (da, db, dc):
        tmm.free_device(p)


@pytest.fixture(params=["i8", "i8:7"], ids=["8-slices", "7-slices"])
def f64_on_int8(request):
    """TMM_F64_MATH=i8[:S]: DGEMM as S (S + 1) / 2 exact int8 slice GEMMs on tcgen05 (gemm_f64_i8.cu).  Read per launch."""
    os.environ["TMM_F64_MATH"] = request.param
    yield
    os.environ.pop("TMM_F64_MATH", None)


@pytest.mark.parametrize("tt", ALL_TT)
def test_dgemm_on_int8_tensor_cores_exact_on_integers(gpu_tmm, oracle, f64_on_int8, tt):
    run_case(gpu_tmm, oracle, np.float64, tt, 130, 67, 95, 2.0, -1.0, pad=(1, 2, 3), ints=True)
    run_case(gpu_tmm, oracle, np.float64, tt, 1000, 520, 1100, 1.0, 0.0, pad=(4, 8, 0), ints=True)


@pytest.mark.parametrize("tt", ["NN", "TN", "NT", "TT"])
def test_dgemm_on_int8_tensor_cores_random_within_the_fp64_bound(gpu_tmm, oracle, f64_on_int8, tt):
    """uniform(-1, 1) data against the tests' FP64 tolerance (1e-15 relative to k max|A| max|B|, tests/_util.py)."""
    run_case(gpu_tmm, oracle, np.float64, tt, 777, 530, 4100, 1.5, 0.25, pad=(3, 2, 9), tiles=(256, 300, 500))


def test_dgemm_on_int8_non_finite_rows_and_columns(gpu_tmm, f64_on_int8):
    """An Inf / NaN in row i of A or column j of B makes row i / column j of C non-finite (NaN here); everything else stays exact."""
    tmm = gpu_tmm
    m, n, k = 200, 150, 300
    rng = np.random.default_rng(4)
    a = rng.integers(0, 10, (k, m)).astype(np.float64).T.copy(order="F")   # column-major m x k
    b = rng.integers(0, 10, (n, k)).astype(np.float64).T.copy(order="F")   # column-major k x n
    a[7, 11] = np.inf; b[5, 140] = np.nan
    c = np.zeros((m, n), order="F")
    da, db, dc = (tmm.malloc_device(x.nbytes) for x in (a, b, c))
    tmm.copy_to_device(a.reshape(-1, order="F"), da); tmm.copy_to_device(b.reshape(-1, order="F"), db); tmm.copy_to_device(c.reshape(-1, order="F"), dc)
    tmm.device_gemm(np.float64, "N", "N", m, n, k, 1.0, da, m, db, k, 0.0, dc, m)
    out = np.empty(m * n); tmm.copy_to_host(dc, out)
    out = out.reshape(n, m).T
    for p in (da, db, dc):
        tmm.free_device(p)
    bad = np.zeros((m, n), bool); bad[7, :] = True; bad[:, 140] = True
    assert np.all(~np.isfinite(out[bad])) and np.all(np.isfinite(out[~bad]))
    a0, b0 = a.copy(), b.copy(); a0[7, 11] = 0; b0[5, 140] = 0
    assert np.array_equal(out[~bad], (a0 @ b0)[~bad])


@pytest.mark.parametrize("tt,beta", [("NN", 0.0), ("TN", -1.0), ("NT", 2.0)])
def test_dgemm_on_int8_through_the_scheduler_with_cached_slices(gpu_tmm, oracle, f64_on_int8, tt, beta):
    """Host-to-host calls in the opt-in mode: several k-chunks, stripes and phase-2 column blocks, so that the slices cached per call (A chunk once
    for all stripes, A over the full k once for all blocks) and the row sub-ranges of a slice stack are exercised; exact on integer data."""
    run_case(gpu_tmm, oracle, np.float64, tt, 1100, 1700, 900, 1.0, beta, pad=(3, 5, 7), ints=True, tiles=(256, 256, 128))
    run_case(gpu_tmm, oracle, np.float64, tt, 777, 1530, 4100, 1.5, beta, pad=(0, 1, 2), tiles=(256, 300, 500))
