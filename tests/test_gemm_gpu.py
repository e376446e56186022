"""GPU suite (-m gpu): parity of the CUDA path, called through the C ABI, against the oracle, the golden vectors and
(when its .so travelled with the repo) the UNMODIFIED reference + cuBLAS on the same device.

Bars: bit-exact on integer-valued inputs (the reference's own known-answer mechanism, tests/test-multiply.cpp:58-66);
max|C - C_ref| / (k max|A| max|B|) <= 1e-15 for FP64 (north_star), 2e-15 complex FP64, 2e-6 / 4e-6 for FP32 / complex FP32.
"""
import itertools
from pathlib import Path

import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
ALL_TT = ["".join(p) for p in itertools.product("NTC", "NTC")]


def run_case(tmm, oracle, dtype, tt, m, n, k, alpha, beta, pad=(0, 0, 0), tiles=(5000, 5000, 5000), streams=2, ints=False, ctx=None,
             copy_modes=(True, False), seed=0, budget=None, pin=False):
    ta, tb = tt
    ar, ac = _util.stored_shape(ta, m, k)
    br, bc = _util.stored_shape(tb, k, n)
    lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
    rng = np.random.default_rng(seed + 17)
    if ints:
        def gen(count):
            v = rng.integers(0, 10, count).astype(np.float64)
            return (v + 1j * rng.integers(0, 10, count)).astype(dtype) if np.dtype(dtype).kind == "c" else v.astype(dtype)
    else:
        def gen(count):
            return _util.random_matrix(rng, dtype, count)
    src_a, src_b, src_c = gen(max(1, lda * ac)), gen(max(1, ldb * bc)), gen(max(1, ldc * n))
    if pin:
        a, b = src_a, src_b  # pageable numpy memory, registered by the call
    else:
        a = tmm.malloc_pinned(dtype, src_a.size); a[:] = src_a
        b = tmm.malloc_pinned(dtype, src_b.size); b[:] = src_b
    expect = oracle.gemm(ta, tb, m, n, k, alpha, src_a, lda, src_b, ldb, beta, src_c.copy(), ldc)
    own = ctx is None
    if own:
        ctx = tmm.make_context(dtype, streams, *tiles)
        if budget:
            ctx.set_device_budget(budget)
    tol = 0.0 if ints else _util.TOL[np.dtype(dtype)]
    amax, bmax = float(np.abs(src_a).max()), float(np.abs(src_b).max())
    for copy_c_back in copy_modes:
        c = src_c.copy() if pin else tmm.malloc_pinned(dtype, src_c.size)
        c[:] = src_c
        tmm.gemm(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers=pin, copy_c_back=copy_c_back)
        if m * n == 0:
            assert np.array_equal(np.asarray(c), src_c), "empty product must leave host C untouched"
            continue
        if copy_c_back:
            got = np.asarray(c)
            assert np.array_equal(got.reshape(n, ldc)[:, m:], src_c.reshape(n, ldc)[:, m:]), "ld padding of host C clobbered"
            got_w, exp_w = got.reshape(n, ldc)[:, :m], expect.reshape(n, ldc)[:, :m]
        else:
            assert np.array_equal(np.asarray(c), src_c), "host C must not be written when copy_c_back=false"
            dv = ctx.get_full_device_buffer_c()
            assert dv.size() == m * n and (dv.data() != 0 or m * n == 0)
            dev = np.empty(m * n, dtype=dtype)
            tmm.copy_to_host(dv.data(), dev, m * n)      # column-major m x n, ld = m (README.md:102-103)
            got_w, exp_w = dev.reshape(n, m), expect.reshape(n, ldc)[:, :m]
        if ints:
            assert np.array_equal(got_w, exp_w), f"{tt} {m}x{n}x{k} copy_c_back={copy_c_back}: not bit-exact"
        else:
            err = float(np.max(np.abs(got_w - exp_w))) / (max(k, 1) * max(amax, 1e-300) * max(bmax, 1e-300)) if m * n else 0.0
            assert err <= tol, f"{tt} {m}x{n}x{k} {np.dtype(dtype)} copy_c_back={copy_c_back}: err {err:.3e} > {tol:.1e}"
    st = ctx.last_stats()
    if own:
        ctx.close()
    return st


# ---- the reference's own registered tests / CI command lines ------------------------------------------------------
def test_c1_square_small_fixture_exact(gpu_tmm, oracle):
    """BASELINE config 1 / ctest `square-small`: dgemm 1000^3 NN alpha=beta=1 on the mt19937(42) integer fixture, both copy
    modes on ONE context (tests/test-multiply.cpp:296-346)."""
    tmm = gpu_tmm
    m = n = k = 1000
    a0, b0, c0 = oracle.fixture_abc(np.float64, m * k, k * n, m * n)
    for beta in (1.0, 0.0):
        expect = oracle.gemm("N", "N", m, n, k, 1.0, a0, m, b0, k, beta, c0.copy(), m)
        a = tmm.malloc_pinned(np.float64, a0.size); a[:] = a0
        b = tmm.malloc_pinned(np.float64, b0.size); b[:] = b0
        c = tmm.malloc_pinned(np.float64, c0.size); c[:] = c0
        c2 = tmm.malloc_pinned(np.float64, c0.size); c2[:] = c0
        ctx = tmm.make_context(np.float64, 2, 5000, 5000, 5000)
        tmm.gemm(ctx, "N", "N", m, n, k, 1.0, a, m, b, k, beta, c, m, False, True)
        assert np.array_equal(np.asarray(c), expect)
        tmm.gemm(ctx, "N", "N", m, n, k, 1.0, a, m, b, k, beta, c2, m, False, False)
        tmm.copy_to_host(ctx.get_full_device_buffer_c().data(), c2, m * n)
        assert np.array_equal(np.asarray(c2), expect)
        assert ctx.last_stats().kernel_launches >= 1
        ctx.close()


@pytest.mark.parametrize("case", [(50, 200, 21, (4, 4, 4)), (5, 2, 2, (4, 4, 4)), (1234, 4567, 1357, (5000, 5000, 5000))])
def test_reference_ci_and_ctest_shapes(gpu_tmm, oracle, case):
    m, n, k, tiles = case  # ci/daint-alps.yml:51,60 ; tests/CMakeLists.txt:13
    st = run_case(gpu_tmm, oracle, np.float64, "NN", m, n, k, 1.0, 0.0, tiles=tiles, ints=True)
    # tile hints never turn into thousands of launches (the reference issues 8750 tile gemms for the first case)
    assert st.kernel_launches <= 32


# ---- transposes, leading dimensions, tiles, scalars ---------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("tt", ALL_TT)
def test_transpose_ld_sweep_exact(gpu_tmm, oracle, dtype, tt):
    alpha, beta = (2.0, -1.0) if np.dtype(dtype).kind == "f" else (1 - 2j, 2 + 1j)
    run_case(gpu_tmm, oracle, dtype, tt, 301, 203, 409, alpha, beta, pad=(7, 13, 5), tiles=(70, 110, 190), ints=True)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
@pytest.mark.parametrize("tt", ["NN", "TN", "NC", "CT"])
def test_random_values_within_tolerance(gpu_tmm, oracle, dtype, tt):
    alpha, beta = (1.5, 0.75) if np.dtype(dtype).kind == "f" else (1.5 - 0.5j, 0.25 + 0.75j)
    run_case(gpu_tmm, oracle, dtype, tt, 777, 530, 1111, alpha, beta, pad=(3, 0, 9), tiles=(256, 300, 500))


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
@pytest.mark.parametrize("tt", ALL_TT)
def test_float_types_exact_on_integers(gpu_tmm, oracle, dtype, tt):
    alpha, beta = (1.0, 1.0) if np.dtype(dtype).kind == "f" else (1 + 0j, 1j)
    run_case(gpu_tmm, oracle, dtype, tt, 130, 67, 95, alpha, beta, pad=(1, 2, 3), ints=True)


def test_c3_sweep_larger_unequal_tiles(gpu_tmm, oracle):
    """BASELINE config 3 shape class: odd sizes, ld > dim, tile_m != tile_n != tile_k."""
    for tt, dtype in (("TN", np.float64), ("NC", np.complex128), ("CT", np.complex128)):
        alpha, beta = (1.0, 1.0) if dtype is np.float64 else (1 - 1j, 0.5 + 0j)
        run_case(gpu_tmm, oracle, dtype, tt, 3001, 2003, 1099, alpha, beta, pad=(7, 129, 33), tiles=(700, 1100, 1900), ints=True)


def test_beta_zero_does_not_read_c(gpu_tmm, oracle):
    """|beta| == 0 => host C is never uploaded (tiled_mm.cpp:325,423): NaNs in C must not propagate."""
    tmm = gpu_tmm
    m, n, k = 257, 190, 300
    rng = np.random.default_rng(3)
    a = tmm.malloc_pinned(np.float64, m * k); a[:] = rng.uniform(-1, 1, m * k)
    b = tmm.malloc_pinned(np.float64, k * n); b[:] = rng.uniform(-1, 1, k * n)
    c = tmm.malloc_pinned(np.float64, m * n, np.nan)
    with tmm.make_context(np.float64) as ctx:
        tmm.gemm(ctx, "N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m, False, True)
        st = ctx.last_stats()
    assert np.isfinite(np.asarray(c)).all()
    assert st.h2d_bytes == 8 * (m * k + k * n)      # every A/B element crosses once; C not uploaded
    assert st.d2h_bytes == 8 * m * n


def test_transfer_volume_once(gpu_tmm, oracle):
    """A and B cross PCIe exactly once in the resident regime (the reference re-sends them per C tile)."""
    st = run_case(gpu_tmm, oracle, np.float64, "NN", 2048, 3000, 1500, 1.0, 1.0, tiles=(500, 500, 500), ints=True, copy_modes=(True,))
    assert st.regime == 0
    assert st.h2d_bytes == 8 * (2048 * 1500 + 1500 * 3000 + 2048 * 3000)
    assert st.d2h_bytes == 8 * 2048 * 3000
    assert st.h2d_bytes < oracle.lib.oracle_reference_h2d_bytes(2048, 3000, 1500, 500, 500, 8, 1)


# ---- context reuse, regimes, edge cases ---------------------------------------------------------------------------
def test_context_reuse_20_calls_grow_and_shrink(gpu_tmm, oracle):
    """BASELINE config 3: copy_c_back=false with one context reused over 20 calls of changing shape."""
    tmm = gpu_tmm
    rng = np.random.default_rng(5)
    for dtype in (np.float64, np.complex128):
        ctx = tmm.make_context(dtype, 3, 300, 200, 250)
        for it in range(20):
            m, n, k = (int(x) for x in rng.integers(1, 400, 3))
            tt = ALL_TT[int(rng.integers(0, 9))]
            beta = [0.0, 1.0, -0.5][it % 3]
            run_case(tmm, oracle, dtype, tt, m, n, k, 1.0, beta, pad=(int(rng.integers(0, 4)),) * 3, ints=True, ctx=ctx,
                     copy_modes=(False,) if it % 4 else (True, False), seed=it)
        ctx.close()


@pytest.mark.parametrize("dtype,tt", [(np.float64, "NN"), (np.float64, "TT"), (np.complex128, "CN"), (np.float32, "NT")])
def test_streaming_regime_out_of_core(gpu_tmm, oracle, dtype, tt):
    """Force the out-of-core path (k-chunk ring, C super-blocks, two C buffers) with a tiny device budget."""
    alpha, beta = (1.0, 1.0) if np.dtype(dtype).kind == "f" else (1 + 1j, 1 - 1j)
    unit = np.dtype(dtype).itemsize << 20  # budgets scale with the element size so every dtype lands in the streaming regime
    st = run_case(gpu_tmm, oracle, dtype, tt, 1500, 1300, 2100, alpha, beta, pad=(1, 2, 3), ints=True, budget=2 * unit)
    assert st.regime == 1
    st = run_case(gpu_tmm, oracle, dtype, tt, 700, 900, 5000, alpha, 0.0, ints=True, budget=2 * unit, copy_modes=(True,))
    assert st.regime == 1 and st.k_chunks >= 2


@pytest.mark.parametrize("shape", [(0, 5, 5), (5, 0, 5), (5, 5, 0), (1, 1, 1), (1, 7, 3), (129, 1, 65)])
def test_degenerate_shapes(gpu_tmm, oracle, shape):
    """m==0 or n==0: no-op; k==0: C = beta*C (BLAS convention; the reference divides by zero here)."""
    m, n, k = shape
    run_case(gpu_tmm, oracle, np.float64, "NN", m, n, k, 1.0, 2.0, pad=(1, 1, 1), ints=True)
    run_case(gpu_tmm, oracle, np.complex128, "CT", m, n, k, 1j, 0.0, pad=(1, 1, 1), ints=True)


def test_alpha_zero_scales_c(gpu_tmm, oracle):
    run_case(gpu_tmm, oracle, np.float64, "NN", 200, 100, 50, 0.0, 3.0, ints=True)
    run_case(gpu_tmm, oracle, np.float64, "NN", 200, 100, 50, 0.0, 0.0, ints=True)


def test_pin_host_buffers_on_pageable_memory(gpu_tmm, oracle):
    """pin_host_buffers=true: pageable pointers are registered for the call and left unregistered (tiled_mm.cpp:529-554,606-618)."""
    run_case(gpu_tmm, oracle, np.float64, "NN", 600, 500, 400, 1.0, 1.0, ints=True, pin=True)
    run_case(gpu_tmm, oracle, np.float64, "NN", 600, 500, 400, 1.0, 1.0, ints=True, pin=True)  # and again: must not be "already registered"


def test_stream_and_tile_hints_never_change_results(gpu_tmm, oracle):
    for streams, tiles in [(1, (4, 4, 4)), (2, (64, 64, 64)), (4, (5000, 5000, 5000)), (7, (128, 4096, 333))]:
        run_case(gpu_tmm, oracle, np.float64, "TN", 513, 300, 777, 1.0, -1.0, tiles=tiles, streams=streams, ints=True, copy_modes=(True,))


def test_error_behaviour(gpu_tmm):
    tmm = gpu_tmm
    a = tmm.malloc_pinned(np.float64, 16)
    with tmm.make_context(np.float64) as ctx:
        with pytest.raises(ValueError):
            tmm.gemm(ctx, "X", "N", 2, 2, 2, 1.0, a, 2, a, 2, 0.0, a, 2, False, True)     # only N/T/C
        with pytest.raises(ValueError):
            tmm.gemm(ctx, "N", "N", 4, 2, 2, 1.0, a, 2, a, 2, 0.0, a, 4, False, True)     # ld_a < rows
        with pytest.raises(ValueError):
            tmm.gemm(ctx, "N", "N", -1, 2, 2, 1.0, a, 2, a, 2, 0.0, a, 2, False, True)
        tmm.gemm(ctx, "n", "t", 2, 2, 2, 1.0, a, 2, a, 2, 0.0, a, 2, False, True)          # lower case accepted (tiled_mm.cpp:503-504)
        assert ctx.optimal_tile_sizes(12345, 23456, 67891) == (4115, 2932, 5000)
        assert ctx.get_max_tile_sizes() == (5000, 5000, 5000) and ctx.get_num_streams() == 2
        # the rest of the handle (mm_handle.cpp:36-80): stream / tile hints never move the construction-time maxima
        ctx.set_streams_and_tiles(3, 100, 200, 300)
        assert ctx.get_num_streams() == 3 and ctx.get_max_tile_sizes() == (5000, 5000, 5000)
        ctx.set_tile_sizes(64)
        ctx.set_num_streams(2)
        assert ctx.get_num_streams() == 2 and ctx.optimal_tile_sizes(12345, 23456, 67891) == (4115, 2932, 5000)
        ctx.set_full_sizes(10, 20, 30)                                                       # full device C sized up front
        assert ctx.get_full_device_buffer_c().size() == 200 and ctx.get_full_device_buffer_c().data() != 0
        assert ctx.stream(tmm.STREAM_COMPUTE, 0) != 0 and ctx.stream(tmm.STREAM_D2H) != 0 and ctx.stream(tmm.STREAM_COMPUTE, 99) == 0
        with pytest.raises(ValueError):
            ctx.set_full_sizes(0, 5)


def test_device_gemm_boundary(gpu_tmm, oracle):
    """blas_api::dgemm replacement on device pointers (the a9 row of SURVEY 8a)."""
    tmm = gpu_tmm
    m, n, k = 300, 200, 100
    a0, b0, c0 = oracle.fixture_abc(np.float64, 304 * k, 104 * n, m * n)
    expect = oracle.gemm("N", "N", m, n, k, 1.0, a0, 304, b0, 104, 1.0, c0.copy(), m)
    da, db, dc = (tmm.malloc_device(x.nbytes) for x in (a0, b0, c0))
    tmm.copy_to_device(a0, da); tmm.copy_to_device(b0, db); tmm.copy_to_device(c0, dc)
    tmm.device_gemm(np.float64, "N", "N", m, n, k, 1.0, da, 304, db, 104, 1.0, dc, m)
    out = np.empty_like(c0)
    tmm.copy_to_host(dc, out)
    assert np.array_equal(out, expect)
    # operands outside the TMA contract (odd ld, base at 8 mod 16) are accepted like cuBLAS accepts them: the copy engine re-pitches
    # them into stream-ordered scratch and the same DMMA kernel runs (reference tests/test-multiply.cpp:44 passes ld_b = k = 1357)
    for tt, oa, lda, ob, ldb in [("NN", 1, 304, 0, 104), ("NN", 0, 301, 0, 103), ("TT", 1, 101, 1, 201), ("TN", 0, 101, 1, 101)]:
        expect = oracle.gemm(tt[0], tt[1], m, n, k, 2.0, a0[oa:], lda, b0[ob:], ldb, -1.0, c0.copy(), m)
        tmm.copy_to_device(c0, dc)
        tmm.device_gemm(np.float64, tt[0], tt[1], m, n, k, 2.0, da + 8 * oa, lda, db + 8 * ob, ldb, -1.0, dc, m)
        tmm.copy_to_host(dc, out)
        assert np.array_equal(out, expect), (tt, oa, lda, ob, ldb)
    with pytest.raises(ValueError):
        tmm.device_gemm(np.float64, "Q", "N", m, n, k, 1.0, da, 304, db, 104, 1.0, dc, m)
    for p in (da, db, dc):
        tmm.free_device(p)


def test_device_sgemm_outside_the_tma_contract(gpu_tmm, oracle):
    """blas_api::sgemm on device pointers with an odd ld or a base at 4 mod 16: re-pitched by the copy engine and run on the tcgen05 kernel
    (round 1 dropped these to the SIMT kernel); exact on integer data like every other path, and the launch count shows which kernel ran."""
    tmm = gpu_tmm
    m, n, k = 300, 200, 100
    a0, b0, c0 = oracle.fixture_abc(np.float32, 305 * max(m, k) + 8, 203 * max(n, k) + 8, m * n)   # room for either stored orientation
    da, db, dc = (tmm.malloc_device(x.nbytes) for x in (a0, b0, c0))
    tmm.copy_to_device(a0, da); tmm.copy_to_device(b0, db)
    out = np.empty_like(c0)
    for tt, oa, lda, ob, ldb in [("NN", 0, 301, 0, 103), ("NN", 1, 304, 0, 104), ("TT", 1, 101, 3, 201), ("TN", 0, 103, 1, 101), ("NT", 2, 303, 0, 202)]:
        expect = oracle.gemm(tt[0], tt[1], m, n, k, 2.0, a0[oa:], lda, b0[ob:], ldb, -1.0, c0.copy(), m)
        tmm.copy_to_device(c0, dc)
        tmm.device_gemm(np.float32, tt[0], tt[1], m, n, k, 2.0, da + 4 * oa, lda, db + 4 * ob, ldb, -1.0, dc, m)
        tmm.copy_to_host(dc, out)
        bad = np.flatnonzero(out != expect)
        assert bad.size == 0, (tt, oa, lda, ob, ldb, int(bad.size), [(int(i % m), int(i // m), float(out[i]), float(expect[i])) for i in bad[:6]])
    for p in (da, db, dc):
        tmm.free_device(p)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_float_non_finite_and_huge_operands(gpu_tmm, dtype):
    """FP32-accurate tensor-core path (3xTF32 split, csrc/tmm_prepass.cuh split_tf32_x4): an Inf / NaN in row i of A or column j of B makes
    row i / column j of C non-finite (the split sends Inf through hi alone, so Inf times an exactly representable value comes out as NaN
    where FP32 arithmetic gives Inf - the documented deviation); a finite operand next to FLT_MAX stays finite; every other
    element is exact on integer data."""
    tmm = gpu_tmm
    m, n, k = 200, 150, 300
    rng = np.random.default_rng(9)
    a = rng.integers(0, 10, (k, m)).astype(dtype).T.copy(order="F")   # column-major m x k
    b = rng.integers(0, 10, (n, k)).astype(dtype).T.copy(order="F")   # column-major k x n
    a[7, 11] = np.inf; b[5, 140] = np.nan
    huge = np.float32(3.4028235e38)                                    # FLT_MAX: rounding it to TF32 must not produce Inf
    a[20, :] = 0; a[20, 3] = huge; b[3, :] = 0; b[3, 9] = np.float32(0.25); b[5, 140] = np.nan
    c = np.zeros((m, n), dtype, order="F")
    da, db, dc = (tmm.malloc_device(x.nbytes) for x in (a, b, c))
    tmm.copy_to_device(a.reshape(-1, order="F"), da); tmm.copy_to_device(b.reshape(-1, order="F"), db); tmm.copy_to_device(c.reshape(-1, order="F"), dc)
    tmm.device_gemm(dtype, "N", "N", m, n, k, 1.0, da, m, db, k, 0.0, dc, m)
    out = np.empty(m * n, dtype); tmm.copy_to_host(dc, out)
    out = out.reshape(n, m).T
    for p in (da, db, dc):
        tmm.free_device(p)
    bad = np.zeros((m, n), bool); bad[7, :] = True; bad[:, 140] = True
    assert np.all(~np.isfinite(out[bad])) and np.all(np.isfinite(out[~bad]))
    a0, b0 = a.astype(np.complex128 if dtype == np.complex64 else np.float64), b.astype(np.complex128 if dtype == np.complex64 else np.float64)
    a0[7, 11] = 0; b0[5, 140] = 0
    want = a0 @ b0
    ok = ~bad
    ok[20, 9] = False
    assert np.array_equal(out[ok], want[ok].astype(dtype))
    assert np.isfinite(out[20, 9]) and abs(out[20, 9] - want[20, 9]) <= 2.0 ** -21 * abs(want[20, 9])   # FLT_MAX / 4, from hi + lo


def test_large_pinned_allocation_and_box_probes(gpu_tmm, oracle):
    """The additive entry points of round 2: tmm_malloc_pinned_large (2 MiB pages + one cudaHostRegister; zero-filled; DMA-able like
    cudaHostAlloc memory - the gemm below runs with pin_host_buffers = false) and the probes behind the bench's roofline denominators."""
    tmm = gpu_tmm
    m, n, k = 700, 500, 300
    a0, b0, c0 = oracle.fixture_abc(np.float64, m * k, k * n, m * n)
    expect = oracle.gemm("N", "N", m, n, k, 1.0, a0, m, b0, k, 1.0, c0.copy(), m)
    a, b, c = (tmm.malloc_pinned_large(np.float64, x.size) for x in (a0, b0, c0))
    assert not np.asarray(a).any() and not np.asarray(c).any(), "fresh large pinned memory must read as zeros"
    a[:] = a0; b[:] = b0; c[:] = c0
    with tmm.make_context(np.float64) as ctx:
        tmm.gemm(ctx, "N", "N", m, n, k, 1.0, a, m, b, k, 1.0, c, m, pin_host_buffers=False, copy_c_back=True)
        st = ctx.last_stats()
    assert np.array_equal(np.asarray(c), expect)
    assert st.h2d_bytes == 8 * (m * k + k * n + m * n) and st.d2h_bytes == 8 * m * n
    del a, b, c                                            # tmm_free_pinned: cudaHostUnregister + munmap for this kind
    peak = tmm.probe_fp64_peak()
    assert 25.0 < peak < 60.0, peak                        # B200: 148 SMs x 64 FMA/clk x 2 x ~1.9 GHz = 37 TF
    (up, down), = tmm.probe_host_links([0], nbytes=32 << 20)
    assert 5.0 < up < 80.0 and 5.0 < down < 80.0, (up, down)   # PCIe Gen5 x16: ~55 GB/s each way alone


# ---- golden vectors and the real reference ------------------------------------------------------------------------
def test_golden_vectors(gpu_tmm, oracle):
    tmm = gpu_tmm
    files = sorted(GOLDEN.glob("*.npz"))
    assert files
    for f in files:
        g = np.load(f, allow_pickle=False)
        ta, tb = str(g["trans"])[0], str(g["trans"])[1]
        m, n, k = (int(x) for x in g["mnk"])
        lda, ldb, ldc = (int(x) for x in g["lds"])
        dtype = g["c_out"].dtype
        if bool(g["seed42"]):
            ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
            a0, b0, c0 = oracle.fixture_abc(dtype, lda * ac, ldb * bc, ldc * n)
        else:
            a0, b0, c0 = g["a"], g["b"], g["c_in"]
        a = tmm.malloc_pinned(dtype, a0.size); a[:] = a0
        b = tmm.malloc_pinned(dtype, b0.size); b[:] = b0
        c = tmm.malloc_pinned(dtype, c0.size); c[:] = c0
        with tmm.make_context(dtype, 2, *(int(x) for x in g["tiles"])) as ctx:
            tmm.gemm(ctx, ta, tb, m, n, k, g["alpha"][()], a, lda, b, ldb, g["beta"][()], c, ldc, False, True)
        if bool(g["exact"]):
            assert np.array_equal(np.asarray(c), g["c_out"]), f.name
        else:
            err = _util.rel_err(np.asarray(c), g["c_out"], k, np.abs(a0).max(), np.abs(b0).max(), m, n, ldc)
            assert err <= _util.TOL[np.dtype(dtype)], (f.name, err)


def test_against_unmodified_reference_with_cublas(gpu_tmm, oracle):
    """Same inputs through the reference library + cuBLAS (oracle/_ref/libtiledmm_ref.so) and through this library."""
    tmm = gpu_tmm
    try:
        ref = _util.Reference(cpu=False)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"reference .so not available on this box: {e}")
    rng = np.random.default_rng(9)
    cases = [(np.float64, "NN", 1000, 1000, 1000, (5000, 5000, 5000)), (np.float64, "NT", 1234, 457, 1357, (500, 300, 400)),
             (np.float64, "TN", 900, 800, 1000, (300, 400, 500)),   # k multiple of tile_k: inside the reference's valid domain (Q8)
             (np.complex128, "CN", 400, 300, 600, (200, 150, 300)), (np.float32, "NN", 512, 384, 640, (256, 128, 320)),
             (np.complex64, "NC", 300, 200, 256, (128, 128, 128))]
    for dtype, tt, m, n, k, tiles in cases:
        ta, tb = tt
        ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
        a0 = _util.random_matrix(rng, dtype, ar * ac); b0 = _util.random_matrix(rng, dtype, br * bc); c0 = _util.random_matrix(rng, dtype, m * n)
        alpha, beta = (1.25, -0.5) if np.dtype(dtype).kind == "f" else (1.25 - 0.5j, 0.5 + 0.25j)
        a = tmm.malloc_pinned(dtype, a0.size); a[:] = a0
        b = tmm.malloc_pinned(dtype, b0.size); b[:] = b0
        c_ref = tmm.malloc_pinned(dtype, c0.size); c_ref[:] = c0
        c_new = tmm.malloc_pinned(dtype, c0.size); c_new[:] = c0
        rctx = ref.context(dtype, 2, *tiles)
        rctx.gemm(ta, tb, m, n, k, alpha, a, ar, b, br, beta, c_ref, m, pin=False, copy_c_back=True)
        rctx.close()
        with tmm.make_context(dtype, 2, *tiles) as ctx:
            tmm.gemm(ctx, ta, tb, m, n, k, alpha, a, ar, b, br, beta, c_new, m, False, True)
        err = _util.rel_err(np.asarray(c_new), np.asarray(c_ref), k, np.abs(a0).max(), np.abs(b0).max(), m, n, m)
        assert err <= _util.TOL[np.dtype(dtype)], (np.dtype(dtype), tt, err)


# ---- full-size property check -------------------------------------------------------------------------------------
def test_full_size_10000_cubed_linearity(gpu_tmm):
    """BASELINE config 2 at full size: C x = A (B x) for a random x (size-independent property; a 10000^3 oracle run would
    take minutes).  Also pins the transfer volume and the launch count of the headline configuration."""
    tmm = gpu_tmm
    n = 10000
    rng = np.random.default_rng(1)
    a = tmm.malloc_pinned(np.float64, n * n); b = tmm.malloc_pinned(np.float64, n * n); c = tmm.malloc_pinned(np.float64, n * n, np.nan)
    a[:] = rng.uniform(-1, 1, n * n); b[:] = rng.uniform(-1, 1, n * n)
    with tmm.make_context(np.float64, 2, 5000, 5000, 5000) as ctx:
        tmm.gemm(ctx, "N", "N", n, n, n, 1.0, a, n, b, n, 0.0, c, n, False, True)
        st = ctx.last_stats()
    A = np.asarray(a).reshape(n, n).T; B = np.asarray(b).reshape(n, n).T; C = np.asarray(c).reshape(n, n).T
    x = rng.uniform(-1, 1, n)
    lhs, rhs = C @ x, A @ (B @ x)
    # |C x - A B x| <= k * eps-level * |A||B||x| ; allow 1e-15 * k per product entry as in north_star, times sum |x|
    assert np.max(np.abs(lhs - rhs)) <= 1e-15 * n * np.abs(x).sum() + 1e-9
    assert st.h2d_bytes == 2 * 8 * n * n and st.d2h_bytes == 8 * n * n and st.kernel_launches <= 64
