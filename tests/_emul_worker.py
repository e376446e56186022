"""Worker of tests/test_scheduler_emulated.py: drives the product's REAL scheduler and multi-GPU layer (compiled as plain C++ over the
CPU emulation in tests/emul/) through the product's own Python binding, and checks every result against the oracle.

  python tests/_emul_worker.py single|grid <n_devices> <direct|nccl>

Test infrastructure: the binding is pointed at tests/emul/_build/libtiledmm_emul.so here, in this process only."""
import ctypes
import itertools
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _util  # noqa: E402
import tiled_mm_b200 as tmm  # noqa: E402

EMUL = Path(os.environ.get("TMM_EMUL_LIB", str(ROOT / "tests" / "emul" / "_build" / "libtiledmm_emul.so")))
tmm.LIB_PATH = EMUL          # test-only redirection of the ctypes binding; the product never does this
lib = tmm.load_library()
lib.emul_violations.restype = ctypes.c_uint64
lib.emul_tma_contract_violations.restype = ctypes.c_uint64
lib.emul_unpinned_async_copies.restype = ctypes.c_uint64
lib.emul_first_violation.restype = ctypes.c_char_p
lib.emul_live_device_bytes.restype = ctypes.c_uint64
lib.emul_live_device_bytes.argtypes = [ctypes.c_int]
lib.emul_races.restype = ctypes.c_uint64
lib.emul_first_race.restype = ctypes.c_char_p
if os.environ.get("TMM_NCCL_LIB"):  # let the race detector see that a blocking collective orders the host threads of its ranks
    _stub = ctypes.CDLL(os.environ["TMM_NCCL_LIB"], mode=ctypes.RTLD_GLOBAL)
    _stub.nccl_stub_set_sync_hook(ctypes.cast(lib.emul_collective, ctypes.c_void_p))
    _stub.nccl_stub_set_stream_hook(ctypes.cast(lib.emul_collective_stream, ctypes.c_void_p))
if os.environ.get("TMM_EMUL_C32_TC") == "1":   # complex<float> takes the scheduler's prepared-operand path (stand-ins in tests/emul/emul_blas.cpp)
    tmm.set_c32_math(3)
oracle = _util.Oracle()
ALL_TT = ["".join(p) for p in itertools.product("NTC", "NTC")]


def gen(rng, dtype, count):
    v = rng.integers(0, 10, count).astype(np.float64)
    return (v + 1j * rng.integers(0, 10, count)).astype(dtype) if np.dtype(dtype).kind == "c" else v.astype(dtype)


def case(ctx, dtype, tt, m, n, k, alpha, beta, pad, copy_modes=(True, False), seed=0, pageable=False, nan_c=False):
    ta, tb = tt
    ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
    lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
    rng = np.random.default_rng(seed)
    a0, b0, c0 = gen(rng, dtype, lda * ac), gen(rng, dtype, ldb * bc), gen(rng, dtype, ldc * n)
    if nan_c:  # beta == 0: C must never be read (reference tiled_mm.cpp:325)
        c0[:] = np.nan
    expect = oracle.gemm(ta.upper(), tb.upper(), m, n, k, alpha, a0, lda, b0, ldb, beta, c0.copy(), ldc)
    if pageable:
        a, b = a0, b0
    else:
        a = tmm.malloc_pinned(dtype, a0.size); a[:] = a0
        b = tmm.malloc_pinned(dtype, b0.size); b[:] = b0
    for copy_c_back in copy_modes:
        c = c0.copy() if pageable else tmm.malloc_pinned(dtype, c0.size)
        c[:] = c0
        tmm.gemm(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers=pageable, copy_c_back=copy_c_back)
        if copy_c_back:
            assert np.array_equal(np.asarray(c), expect, equal_nan=True), f"{np.dtype(dtype)} {tt} {m}x{n}x{k} copy-back result differs from the oracle"
        else:
            assert np.array_equal(np.asarray(c), c0, equal_nan=True), "host C written although copy_c_back = false"
            dev = np.empty(m * n, dtype=dtype)
            tmm.copy_to_host(ctx.get_full_device_buffer_c().data(), dev, m * n)
            assert np.array_equal(dev.reshape(n, m), expect.reshape(n, ldc)[:, :m]), f"{tt} device-resident C differs from the oracle"
    return ctx.last_stats()


def check_clean(where):
    assert lib.emul_violations() == 0, f"{where}: {lib.emul_first_violation().decode()}"
    assert lib.emul_tma_contract_violations() == 0, f"{where}: the scheduler built a panel outside the TMA contract"
    assert lib.emul_unpinned_async_copies() == 0, f"{where}: an async copy touched pageable host memory"
    assert lib.emul_races() == 0, f"{where}: unordered conflicting accesses (missing dependency in the schedule): {lib.emul_first_race().decode()}"


def run_single():
    assert tmm.device_count() >= 1
    es = 8
    with tmm.make_context(np.float64, 2, 64, 64, 64) as ctx:
        for tt in ALL_TT:                                                  # resident regime, every op pair, padded lds
            st = case(ctx, np.float64, tt, 301, 203, 409, 2.0, -1.0, (7, 13, 5), seed=1)
            assert st.regime == 0 and st.h2d_bytes == es * (301 * 409 + 409 * 203 + 301 * 203), (tt, st.h2d_bytes)
        check_clean("resident f64")
        ctx.set_device_budget(4 << 20)                                     # 4 MiB: streaming ring, C super-blocks, two C buffers
        for tt in ("NN", "TN", "NT", "CC"):
            st = case(ctx, np.float64, tt, 700, 900, 500, 1.0, 1.0, (3, 5, 7), seed=2)
            assert st.regime == 1, (tt, st.regime, st.c_blocks)
        check_clean("streaming f64")
        ctx.set_device_budget(0)
        for shape in [(1, 1, 1), (5, 2, 2), (50, 200, 21), (129, 1, 65), (1, 300, 7), (64, 64, 2049)]:
            case(ctx, np.float64, "TN", *shape, 1.0, 0.0, (0, 0, 0), seed=3)
        case(ctx, np.float64, "NN", 600, 500, 400, 1.0, 1.0, (1, 2, 3), pageable=True, seed=4)   # cudaHostRegister path, unregistered again after the call
        case(ctx, np.float64, "NN", 600, 500, 400, 1.0, 1.0, (1, 2, 3), pageable=True, seed=4)
        # a pageable operand of 288 MB: registered in 2 MiB-aligned pieces from several host threads, all of them released after the call
        lib.emul_registered_host_blocks.restype = ctypes.c_uint64
        for tt in ("NN", "TN"):
            case(ctx, np.float64, tt, 36000, 8, 1000, 1.0, 1.0, (0, 0, 0), copy_modes=(True,), pageable=True, seed=12)
            assert lib.emul_registered_host_blocks() == 0, "host registrations left behind"
        check_clean("shapes + pageable")
    for dtype, alpha, beta in [(np.complex128, 1 - 2j, 2 + 1j), (np.float32, 1.0, 1.0), (np.complex64, 1 + 1j, 1j)]:
        with tmm.make_context(dtype, 3, 100, 70, 50) as ctx:
            for tt in ("NN", "CT", "TC"):
                case(ctx, dtype, tt, 257, 129, 300, alpha, beta, (1, 2, 3), seed=5)
            ctx.set_device_budget(4 << 20)
            case(ctx, dtype, "CN", 513, 300, 777, alpha, beta, (2, 0, 1), seed=6)
    check_clean("other dtypes")
    # BLAS quick returns and argument errors - the same front end the GPU build runs
    with tmm.make_context(np.float64) as ctx:
        for shape in [(0, 5, 5), (5, 0, 5), (5, 5, 0), (7, 3, 0)]:
            if shape[0] * shape[1]:
                case(ctx, np.float64, "NN", *shape, 1.0, 2.0, (1, 1, 1), seed=10)        # k = 0: C = beta C
            else:                                                                        # empty C: nothing to do, nothing dereferenced
                tmm.gemm(ctx, "N", "N", *shape, 1.0, None, max(1, shape[0]), None, max(1, shape[2]), 0.0, None, max(1, shape[0]), False, True)
        case(ctx, np.float64, "TN", 40, 30, 20, 0.0, -1.5, (1, 2, 3), seed=11)          # alpha = 0: C = beta C, A and B never touched
        buf = tmm.malloc_pinned(np.float64, 64)
        for bad in [lambda: tmm.gemm(ctx, "X", "N", 2, 2, 2, 1.0, buf, 2, buf, 2, 0.0, buf, 2, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 4, 2, 2, 1.0, buf, 2, buf, 2, 0.0, buf, 4, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 2, 2, 4, 1.0, buf, 2, buf, 2, 0.0, buf, 2, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 2, 2, 2, 1.0, buf, 2, buf, 2, 0.0, buf, 1, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", -1, 2, 2, 1.0, buf, 2, buf, 2, 0.0, buf, 2, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 2**31 - 1, 2**31 - 1, 2**31 - 1, 1.0, buf, 2**31 - 1, buf, 2**31 - 1, 0.0, buf, 2**31 - 1, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 2**31, 2, 2, 1.0, buf, 2**31, buf, 2, 0.0, buf, 2**31, False, True),
                    lambda: tmm.gemm(ctx, "N", "N", 2, 2, 2, 1.0, None, 2, buf, 2, 0.0, buf, 2, False, True)]:
            try:
                bad()
                raise AssertionError("invalid call was accepted")
            except ValueError:
                pass
        tmm.gemm(ctx, "n", "t", 2, 2, 2, 1.0, buf, 2, buf, 2, 0.0, buf, 2, False, True)   # lower case accepted (tiled_mm.cpp:503-504)
    check_clean("quick returns and argument errors")
    check_clean("before device-pointer operands")
    # device-pointer operands (additive): all on the device -> one launch, no staging; mixed -> scheduler with inferred copies
    for dtype, alpha, beta in [(np.float64, 2.0, -1.0), (np.complex128, 1 - 2j, 1j)]:
        rng = np.random.default_rng(77)
        m, n, k, lda, ldb, ldc = 130, 95, 170, 133, 171, 131                 # odd leading dimensions on purpose
        a0, b0, c0 = gen(rng, dtype, lda * k), gen(rng, dtype, ldb * n), gen(rng, dtype, ldc * n)
        expect = oracle.gemm("N", "N", m, n, k, alpha, a0, lda, b0, ldb, beta, c0.copy(), ldc)
        da, db, dc = tmm.malloc_device(a0.nbytes), tmm.malloc_device(b0.nbytes), tmm.malloc_device(c0.nbytes)
        tmm.copy_to_device(a0, da); tmm.copy_to_device(b0, db)
        with tmm.make_context(dtype, 2, 64, 64, 64) as ctx:
            tmm.copy_to_device(c0, dc)
            tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc, pin_host_buffers=True, copy_c_back=True)   # all three on the device
            assert ctx.last_stats().h2d_bytes == 0 and ctx.last_stats().kernel_launches == 1
            out = np.empty_like(c0); tmm.copy_to_host(dc, out)
            assert np.array_equal(out, expect), "device operands, result in place"
            tmm.copy_to_device(c0, dc)
            tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc, pin_host_buffers=False, copy_c_back=False)  # result in the context's C
            out2 = np.empty(m * n, dtype=dtype); tmm.copy_to_host(ctx.get_full_device_buffer_c().data(), out2, m * n)
            assert np.array_equal(out2.reshape(n, m), expect.reshape(n, ldc)[:, :m]), "device operands, device-resident C"
            tmm.copy_to_host(dc, out); assert np.array_equal(out, c0), "c must not be written when copy_c_back = false"
            ch = tmm.malloc_pinned(dtype, c0.size); ch[:] = c0                                                                  # mixed: A on the device, B and C on the host
            bh = tmm.malloc_pinned(dtype, b0.size); bh[:] = b0
            tmm.gemm(ctx, "N", "N", m, n, k, alpha, da, lda, bh, ldb, beta, ch, ldc, pin_host_buffers=False, copy_c_back=True)
            assert np.array_equal(np.asarray(ch), expect), "mixed host / device operands"
        for p in (da, db, dc):
            tmm.free_device(p)
    lib.emul_reset_tma_contract_violations()   # odd-ld USER operands went straight to the GEMM layer here (the real one re-pitches them)
    check_clean("device-pointer operands")
    # bf16 entry point: argument checking and plumbing of tmm_device_gemm_bf16 (the arithmetic here is the GEMM double's)
    rng = np.random.default_rng(9)
    m, n, k = 70, 50, 90
    af, bf = (rng.integers(-8, 9, 95 * m).astype(np.float32), rng.integers(-8, 9, 55 * k).astype(np.float32))   # stored k x m (ld 95), n x k (ld 55)
    c0 = rng.integers(0, 10, m * n).astype(np.float32)
    to_bf16 = lambda x: (x.view(np.uint32) >> 16).astype(np.uint16)
    da, db, dc = tmm.malloc_device(af.size * 2), tmm.malloc_device(bf.size * 2), tmm.malloc_device(c0.nbytes)
    tmm.copy_to_device(to_bf16(af), da); tmm.copy_to_device(to_bf16(bf), db); tmm.copy_to_device(c0, dc)
    tmm.device_gemm_bf16("T", "T", m, n, k, 2.0, da, 95, db, 55, -1.0, dc, m)
    got = np.empty_like(c0); tmm.copy_to_host(dc, got)
    want = oracle.gemm("T", "T", m, n, k, np.float32(2.0), af, 95, bf, 55, np.float32(-1.0), c0.copy(), m)
    assert np.array_equal(got, want), "bf16 entry point"
    try:
        tmm.device_gemm_bf16("N", "N", m, n, k, 1.0, da, m - 1, db, k, 0.0, dc, m)
        raise AssertionError("ld_a < m must be rejected")
    except ValueError:
        pass
    for p in (da, db, dc):
        tmm.free_device(p)
    check_clean("bf16 entry")
    assert lib.emul_live_device_bytes(0) == 0, "device memory leaked after the contexts were destroyed"
    print("EMUL_OK single")


def run_grid(n_dev, plane):
    assert tmm.device_count() == n_dev, tmm.device_count()
    pr, pc = tmm.grid_shape(n_dev)
    for dtype, alpha, beta in [(np.float64, 2.0, -1.0), (np.complex128, 1 - 2j, 2 + 1j)]:
        es = np.dtype(dtype).itemsize
        with tmm.make_context(dtype, 2, 64, 64, 64) as ctx:
            ctx.set_devices(n_dev)
            assert ctx.num_devices() == n_dev
            for tt in ("NN", "TN", "NT", "CC"):
                m, n, k = 301, 403, 209
                st = case(ctx, dtype, tt, m, n, k, alpha, beta, (1, 2, 3), copy_modes=(True,), seed=7)
                # every shared panel element crosses "PCIe" exactly once over the whole grid, the rest travels GPU to GPU
                assert st.regime == 0 and st.h2d_bytes == es * (m * k + k * n + m * n), (tt, st.h2d_bytes)
                want_peer = es * (m * k * (pc - 1) + k * n * (pr - 1))
                assert st.peer_bytes == want_peer, (tt, st.peer_bytes, want_peer)
            check_clean(f"grid {pr}x{pc} resident {np.dtype(dtype)}")
            ctx.set_devices(n_dev)                                         # fresh children
            ctx.set_device_budget(3 << 20)
            ctx.set_devices(n_dev)                                         # children inherit the budget: streaming ring + acks
            for tt in ("NN", "TT"):
                st = case(ctx, dtype, tt, 900, 700, 1100, alpha, beta, (1, 2, 3), copy_modes=(True,), seed=8)
                assert st.regime == 1, st.regime
            check_clean(f"grid {pr}x{pc} streaming {np.dtype(dtype)}")
            ctx.set_device_budget(0)
            ctx.set_devices(n_dev)
            # ragged blocks: fewer rows / columns than a balanced split would like, and a shape too small for the grid
            for shape in [(pr, pc, 5), (pr + 1, 2 * pc + 1, 33), (7, 200, 64), (1, 1, 1)]:
                case(ctx, dtype, "NT", *shape, alpha, beta, (0, 1, 0), copy_modes=(True,), seed=9)
            check_clean(f"grid {pr}x{pc} ragged {np.dtype(dtype)}")
    for d in range(n_dev):
        assert lib.emul_live_device_bytes(d) == 0, f"device {d} memory leaked"
    print(f"EMUL_OK grid {pr}x{pc} {plane}")


def run_sweep(n_dev, n_cases, seed):
    """Randomised sweep in the manner of the one that pinned the reference's valid domain: random element type, op
    pair (upper / lower case), dims 1-260, ld padding 0-5, alpha / beta (incl. beta = 0 over a NaN-filled C), tile hints 1-300, 1-4
    streams, both copy modes, tight device budgets (streaming regime with C super-blocks), contexts REUSED across calls."""
    rng = np.random.default_rng(seed)
    dtypes = [np.float64, np.complex128, np.float32, np.complex64]
    if os.environ.get("TMM_EMUL_DTYPES"):   # e.g. "c": complex<float> only
        dtypes = [{"d": np.float64, "z": np.complex128, "s": np.float32, "c": np.complex64}[ch] for ch in os.environ["TMM_EMUL_DTYPES"]] * 4
    done = 0
    while done < n_cases:
        dtype = dtypes[int(rng.integers(0, 4))]
        cplx = np.dtype(dtype).kind == "c"
        ctx = tmm.make_context(dtype, int(rng.integers(1, 5)), *(int(x) for x in rng.integers(1, 300, 3)))
        if n_dev > 1:
            ctx.set_devices(n_dev)
        for _ in range(int(rng.integers(3, 9))):                      # several calls on one context: grow and shrink
            if n_dev == 1 or rng.random() < 0.5:
                budget = int(rng.choice([0, 0, 96 << 10, 256 << 10, 1 << 20]))
                ctx.set_device_budget(budget)
                if n_dev > 1:
                    ctx.set_devices(n_dev)                           # children pick the budget up when they are created
            tt = "".join(rng.choice(list("NTCntc"), 2))
            m, n, k = (int(x) for x in rng.integers(1, 261, 3))
            if rng.random() < 0.15:
                k = int(rng.integers(1, 2500))                        # long k: many chunks
            pad = tuple(int(x) for x in rng.integers(0, 6, 3))
            alpha = complex(*rng.integers(-2, 3, 2)) if cplx else float(rng.integers(-2, 3))
            beta = [0.0, 1.0, complex(1, -1) if cplx else -1.5][int(rng.integers(0, 3))]
            modes = (True,) if n_dev > 1 else ((True, False) if rng.random() < 0.5 else (bool(rng.integers(0, 2)),))
            try:
                nan_c = beta == 0.0
                case(ctx, dtype, tt, m, n, k, alpha, beta, pad, copy_modes=modes, seed=int(rng.integers(0, 1 << 30)), pageable=rng.random() < 0.2, nan_c=nan_c)
            except RuntimeError as e:
                if "budget too small" not in str(e):
                    raise
            done += 1
        ctx.close()
        check_clean(f"sweep seed {seed} after {done} cases ({np.dtype(dtype)}, {n_dev} devices)")
    for d in range(n_dev):
        assert lib.emul_live_device_bytes(d) == 0
    if os.environ.get("TMM_EMUL_C32_TC") == "1":
        lib.emul_prepared_cgemm_launches.restype = ctypes.c_uint64
        launches = lib.emul_prepared_cgemm_launches()
        assert launches > done, f"the prepared-operand path took only {launches} launches in {done} complex<float> calls"
        print(f"prepared-operand launches: {launches}")
    print(f"EMUL_OK sweep {n_dev} devices, {done} cases")


def run_replan():
    """Free HBM shrinks between two calls (another allocation appears): the cached budget is stale, a panel allocation fails before
    anything is enqueued, and the call must recover by re-planning into the streaming regime instead of failing."""
    with tmm.make_context(np.float64, 2, 5000, 5000, 5000) as ctx:
        st = case(ctx, np.float64, "NN", 500, 500, 500, 1.0, 1.0, (0, 0, 0), copy_modes=(True,), seed=1)
        assert st.regime == 0
        foreign = tmm.malloc_device(14 << 20)                         # somebody else takes 14 of the 32 MiB
        st = case(ctx, np.float64, "TN", 900, 900, 900, 2.0, -1.0, (1, 2, 3), copy_modes=(True,), seed=2)
        assert st.regime == 1, "expected the retry to land in the streaming regime"
        st = case(ctx, np.float64, "NT", 900, 900, 900, 1.0, 0.0, (0, 0, 0), copy_modes=(True, False), seed=3)   # budget now re-read: no failure
        tmm.free_device(foreign)
        st = case(ctx, np.float64, "NN", 300, 300, 300, 1.0, 1.0, (0, 0, 0), copy_modes=(True,), seed=4)
        check_clean("re-plan after allocation failure")
    assert lib.emul_live_device_bytes(0) == 0
    print("EMUL_OK replan")


def run_threads(n_threads, per_thread):
    """One context per host thread, all on the same device, calls in flight concurrently (the reference's intended use: 'one handle
    per host thread').  ctypes releases the GIL during tmm_gemm, so the library's process-wide state is really shared."""
    import threading
    errors = []

    def body(tid):
        try:
            rng = np.random.default_rng(1000 + tid)
            dtype = [np.float64, np.complex128, np.float32, np.complex64][tid % 4]
            with tmm.make_context(dtype, 2, 64, 64, 64) as ctx:
                for i in range(per_thread):
                    m, n, k = (int(x) for x in rng.integers(1, 200, 3))
                    if i % 7 == 3:
                        ctx.set_device_budget(256 << 10)
                    elif i % 7 == 5:
                        ctx.set_device_budget(0)
                    try:
                        case(ctx, dtype, "".join(rng.choice(list("NTC"), 2)), m, n, k, 1.0, [0.0, 1.0][i % 2], (1, 0, 2), copy_modes=(bool(i % 3),), seed=tid * 10000 + i)
                    except RuntimeError as e:
                        if "budget too small" not in str(e):
                            raise
        except BaseException as e:  # noqa: BLE001 - reported by the main thread
            errors.append(f"thread {tid}: {type(e).__name__}: {e}")

    threads = [threading.Thread(target=body, args=(t,)) for t in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    check_clean("concurrent contexts")
    assert lib.emul_live_device_bytes(0) == 0
    print(f"EMUL_OK threads {n_threads} x {per_thread}")


def run_drysweep(n_dev, n_cases, seed):
    """Address-only sweep over REALISTIC sizes (dims up to 60000, budgets from 1 GiB to all of HBM): every plan shape the scheduler can
    produce at scale - several column stripes, many k-chunks, phase-2 blocks, C super-blocks with two buffers, ring reuse - is walked with
    bounds, ordering and byte-count checks, without arithmetic."""
    assert lib.emul_dry_run() == 1
    rng = np.random.default_rng(seed)
    pr, pc = tmm.grid_shape(n_dev)
    A, B, C = 0x200000000000, 0x300000000000, 0x400000000000
    dtypes = [np.float64, np.complex128, np.float32, np.complex64]
    if os.environ.get("TMM_EMUL_DTYPES"):   # e.g. "c": complex<float> only
        dtypes = [{"d": np.float64, "z": np.complex128, "s": np.float32, "c": np.complex64}[ch] for ch in os.environ["TMM_EMUL_DTYPES"]] * 4
    done = 0
    while done < n_cases:
        dtype = dtypes[int(rng.integers(0, 4))]
        es = np.dtype(dtype).itemsize
        with tmm.make_context(dtype, int(rng.integers(1, 5)), *(int(x) for x in rng.integers(64, 8000, 3))) as ctx:
            if n_dev > 1:
                ctx.set_devices(n_dev)
            for _ in range(4):
                budget = int(rng.choice([0, 0, 1 << 30, 6 << 30, 40 << 30]))
                ctx.set_device_budget(budget)
                if n_dev > 1:
                    ctx.set_devices(n_dev)
                m, n, k = (int(x) for x in np.exp(rng.uniform(np.log(64), np.log(60000), 3)))
                ta, tb = rng.choice(list("NTC"), 2)
                lda = (m if ta == "N" else k) + int(rng.integers(0, 9))
                ldb = (k if tb == "N" else n) + int(rng.integers(0, 9))
                ldc = m + int(rng.integers(0, 9))
                beta = float(rng.integers(0, 2))
                back = True if n_dev > 1 else bool(rng.integers(0, 2))
                try:
                    tmm.gemm(ctx, ta, tb, m, n, k, 1.0, A, lda, B, ldb, beta, C, ldc, pin_host_buffers=False, copy_c_back=back)
                except RuntimeError as e:
                    if "budget too small" not in str(e) and "out of memory" not in str(e):
                        raise
                    continue
                st = ctx.last_stats()
                once = es * (m * k + k * n + (m * n if beta else 0))
                assert st.d2h_bytes == (es * m * n if back else 0), (m, n, k, st.d2h_bytes)
                if st.regime == 0 and (n_dev == 1 or (m >= pr and n >= pc)):
                    assert st.h2d_bytes == once, (np.dtype(dtype), ta + tb, m, n, k, st.h2d_bytes, once)
                else:
                    assert st.h2d_bytes >= once or n_dev > 1, (m, n, k, st.h2d_bytes, once)
                done += 1
        assert lib.emul_violations() == 0, lib.emul_first_violation().decode()
        assert lib.emul_tma_contract_violations() == 0
        assert lib.emul_races() == 0, lib.emul_first_race().decode()
    for d in range(n_dev):
        assert lib.emul_live_device_bytes(d) == 0
    print(f"EMUL_OK drysweep {n_dev} devices, {done} cases")


def run_auto(n_dev):
    """TMM_DEVICES=n: contexts created by the application drive n GPUs without any code change; copy_c_back=false calls land on the
    first device and get_full_device_buffer_c() follows them there."""
    assert int(os.environ["TMM_DEVICES"]) == n_dev
    for dtype, alpha, beta in [(np.float64, 2.0, -1.0), (np.complex64, 1 - 1j, 1j)]:
        with tmm.make_context(dtype, 2, 64, 64, 64) as ctx:
            assert ctx.num_devices() == n_dev
            for tt in ("NN", "CT"):
                st = case(ctx, dtype, tt, 301, 403, 209, alpha, beta, (1, 2, 3), copy_modes=(True,), seed=20)
                assert st.peer_bytes > 0, "the grid was not used"
                st = case(ctx, dtype, tt, 301, 403, 209, alpha, beta, (1, 2, 3), copy_modes=(False,), seed=21)   # device-resident C
                assert st.peer_bytes == 0
            ctx.set_full_sizes(50, 60, 1)
            assert ctx.get_full_device_buffer_c().size() == 3000 and ctx.get_full_device_buffer_c().data() != 0
    check_clean("TMM_DEVICES")
    for d in range(n_dev):
        assert lib.emul_live_device_bytes(d) == 0
    print(f"EMUL_OK auto {n_dev}")


def run_faults():
    """Fault injection (the reference has none): the n-th device allocation / the n-th 2-D copy of a call fails.  The call
    must come back with an error or recover (allocation failures before anything is enqueued are re-planned), never crash or hang; the
    context must stay usable - the next call is correct - and nothing may leak."""
    lib.emul_inject_fault.argtypes = [ctypes.c_int, ctypes.c_long]
    outcomes = {"error": 0, "recovered": 0, "not reached": 0}
    for kind, name in ((0, "cudaMalloc"), (1, "cudaMemcpy2DAsync")):
        for budget in (0, 2 << 20):
            for nth in range(1, 12):
                ctx = tmm.make_context(np.float64, 2, 64, 64, 64)
                ctx.set_device_budget(budget)
                case(ctx, np.float64, "NN", 150, 140, 130, 1.0, 1.0, (1, 2, 3), copy_modes=(True,), seed=30)      # warm: buffers exist
                ctx.set_device_budget(budget)
                lib.emul_inject_fault(kind, nth)
                try:
                    case(ctx, np.float64, "TN", 400, 380, 500, 2.0, -1.0, (1, 2, 3), copy_modes=(True, False), seed=31)   # larger: buffers must grow
                    outcomes["recovered" if kind == 0 and nth <= 4 else "not reached"] += 1
                except RuntimeError as e:
                    assert "GPU ERROR" in str(e), str(e)
                    outcomes["error"] += 1
                lib.emul_inject_fault(kind, 0)
                case(ctx, np.float64, "NT", 300, 200, 250, 1.0, 0.0, (0, 1, 0), copy_modes=(True, False), seed=32)       # the context still works
                ctx.close()
                assert lib.emul_live_device_bytes(0) == 0, f"{name} #{nth}: device memory leaked"
                assert lib.emul_violations() == 0, lib.emul_first_violation().decode()
    assert outcomes["error"] > 0, outcomes
    print(f"EMUL_OK faults {outcomes}")


def run_grid_faults(n_dev):
    """A device allocation fails on ONE GPU of the grid in the middle of a sequence of calls: every rank must give the call up together
    (nobody may enqueue work that waits for shares which never come - that would be a hang), and the grid must serve the next call."""
    lib.emul_inject_fault.argtypes = [ctypes.c_int, ctypes.c_long]
    import time
    errors = 0
    for plane_budget in (0, 3 << 20):
        for nth in (1, 2, 3, 5, 8):
            ctx = tmm.make_context(np.float64, 2, 64, 64, 64)
            ctx.set_device_budget(plane_budget)
            ctx.set_devices(n_dev)
            case(ctx, np.float64, "NN", 200, 180, 160, 1.0, 0.0, (0, 0, 0), copy_modes=(True,), seed=40)
            lib.emul_inject_fault(0, nth)
            t0 = time.time()
            try:
                case(ctx, np.float64, "TN", 900, 700, 1100, 2.0, -1.0, (1, 2, 3), copy_modes=(True,), seed=41)   # larger: panels must grow
            except RuntimeError as e:
                assert "GPU ERROR" in str(e), str(e)
                errors += 1
            assert time.time() - t0 < 20, "the grid did not give the call up together (a rank waited for its peers)"
            lib.emul_inject_fault(0, 0)
            case(ctx, np.float64, "NT", 500, 400, 300, 1.0, 1.0, (0, 1, 0), copy_modes=(True,), seed=42)
            ctx.close()
            for d in range(n_dev):
                assert lib.emul_live_device_bytes(d) == 0, f"device {d}: memory leaked"
            assert lib.emul_violations() == 0, lib.emul_first_violation().decode()
    assert errors > 0
    print(f"EMUL_OK gridfaults {n_dev} devices, {errors} failed calls handled")


def run_ranks(world):
    """The one-process-per-GPU entry point (tmm_context_attach_grid, what torchrun / MPI jobs use), played by `world` host threads with one
    emulated device each.  With TMM_DIST_FORCE_IPC=1 the peers' panels are imported through the (emulated) CUDA IPC calls, so the
    bookkeeping of imports - close before re-import when a peer's buffer changed, and never free an allocation that a peer still has
    imported (outgrown buffers are retired) - is checked: the emulation reports a cudaFree of an imported allocation."""
    import threading
    from tiled_mm_b200 import multi_gpu
    lib.emul_set_device.argtypes = [ctypes.c_int]
    lib.emul_open_ipc_imports.restype = ctypes.c_uint64
    pr, pc = tmm.grid_shape(world)
    barrier = threading.Barrier(world)
    ids, errors = {}, []
    cases = [(np.float64, "NN", 300, 260, 200, (0, 0, 0), 1.0, 0.0, 0), (np.float64, "TN", 513, 300, 777, (3, 5, 7), 2.0, -1.0, 0),       # panels grow
             (np.float64, "NT", 900, 700, 1100, (1, 2, 3), 1.0, 1.0, 3 << 20),                                                         # streaming ring
             (np.float64, "CN", 200, 150, 100, (1, 0, 2), 1.0, 1.0, 0), (np.float64, "NN", 1200, 1000, 640, (0, 0, 0), 1.0, 0.0, 0),    # shrink, grow again
             (np.complex128, "CT", 301, 403, 209, (1, 2, 3), 1 - 2j, 2 + 1j, 0)]

    def body(rank):
        try:
            assert lib.emul_set_device(rank) == 0
            row, col = rank // pc, rank % pc
            if pc > 1 and col == 0:
                ids[("row", row)] = tmm.dist_unique_id()
            if pr > 1 and row == 0:
                ids[("col", col)] = tmm.dist_unique_id()
            barrier.wait()
            ctxs = {}
            for ci, (dtype, tt, m, n, k, pad, alpha, beta, budget) in enumerate(cases):
                key = np.dtype(dtype)
                if key not in ctxs:
                    ctxs[key] = tmm.make_context(dtype, 2, 64, 64, 64)
                    ctxs[key].attach_grid(pr, pc, row, col, ids.get(("row", row)), ids.get(("col", col)))
                ctx = ctxs[key]
                ctx.set_device_budget(budget)
                ta, tb = tt
                ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
                lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
                rng = np.random.default_rng(100 + ci)                   # the same full problem on every rank
                a0, b0, c0 = gen(rng, dtype, lda * ac), gen(rng, dtype, ldb * bc), gen(rng, dtype, ldc * n)
                expect = oracle.gemm(ta, tb, m, n, k, alpha, a0, lda, b0, ldb, beta, c0.copy(), ldc).reshape(n, ldc)
                i0, i1, j0, j1 = multi_gpu.block_of(rank, world, m, n)
                es = np.dtype(dtype).itemsize
                oa = i0 if ta == "N" else i0 * lda
                ob = j0 * ldb if tb == "N" else j0
                ap = tmm.malloc_pinned(dtype, a0.size); ap[:] = a0
                bp = tmm.malloc_pinned(dtype, b0.size); bp[:] = b0
                cp = tmm.malloc_pinned(dtype, c0.size); cp[:] = c0
                tmm.gemm(ctx, ta, tb, i1 - i0, j1 - j0, k, alpha, ap.ctypes.data + oa * es, lda, bp.ctypes.data + ob * es, ldb, beta,
                         cp.ctypes.data + (j0 * ldc + i0) * es, ldc, pin_host_buffers=False, copy_c_back=True)
                got = np.asarray(cp).reshape(n, ldc)
                assert np.array_equal(got[j0:j1, i0:i1], expect[j0:j1, i0:i1]), f"rank {rank} case {ci}: block differs from the oracle"
                mask = np.ones((n, ldc), dtype=bool); mask[j0:j1, i0:i1] = False
                assert np.array_equal(got[mask], c0.reshape(n, ldc)[mask]), f"rank {rank} case {ci}: wrote outside its block"
                assert ctx.last_stats().regime == (1 if budget else 0)
            barrier.wait()
            if rank == 0:
                check_clean("ranks")
                if os.environ.get("TMM_DIST_FORCE_IPC") == "1":
                    assert lib.emul_open_ipc_imports() > 0, "the IPC branch was not taken"
                # teardown is not collective (a rank may free its panels while a peer still has them imported - as on hardware, where the
                # processes are about to exit); only the steady state above is held to the no-free-while-imported rule
                os.environ["TMM_EMUL_ALLOW_FREE_WHILE_IMPORTED"] = "1"
            barrier.wait()
            for c in ctxs.values():
                c.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(f"rank {rank}: {type(e).__name__}: {e}")
            barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert lib.emul_violations() == 0, lib.emul_first_violation().decode()
    assert lib.emul_open_ipc_imports() == 0, "imports left open"
    for d in range(world):
        assert lib.emul_live_device_bytes(d) == 0
    print(f"EMUL_OK ranks {pr}x{pc}")


def run_rank_fails_early(world):
    """One rank of the grid rejects its arguments BEFORE the agreement round (ld_a too small there only); the others must not wait for it:
    every rank gets an error for that call, and the next call on the same contexts works (the rounds stayed in step)."""
    import threading
    from tiled_mm_b200 import multi_gpu
    lib.emul_set_device.argtypes = [ctypes.c_int]
    pr, pc = tmm.grid_shape(world)
    barrier = threading.Barrier(world)
    ids, errors, failed_calls = {}, [], []

    def body(rank):
        try:
            assert lib.emul_set_device(rank) == 0
            row, col = rank // pc, rank % pc
            if pc > 1 and col == 0:
                ids[("row", row)] = tmm.dist_unique_id()
            if pr > 1 and row == 0:
                ids[("col", col)] = tmm.dist_unique_id()
            barrier.wait()
            dtype, m, n, k = np.float64, 300, 260, 200
            ctx = tmm.make_context(dtype, 2, 64, 64, 64)
            ctx.attach_grid(pr, pc, row, col, ids.get(("row", row)), ids.get(("col", col)))
            rng = np.random.default_rng(3)
            a0, b0, c0 = gen(rng, dtype, m * k), gen(rng, dtype, k * n), gen(rng, dtype, m * n)
            expect = oracle.gemm("N", "N", m, n, k, 1.0, a0, m, b0, k, 0.0, c0.copy(), m).reshape(n, m)
            i0, i1, j0, j1 = multi_gpu.block_of(rank, world, m, n)
            ap = tmm.malloc_pinned(dtype, a0.size); ap[:] = a0
            bp = tmm.malloc_pinned(dtype, b0.size); bp[:] = b0
            cp = tmm.malloc_pinned(dtype, c0.size); cp[:] = c0
            for attempt, bad_rank in enumerate([world - 1, 0, None]):
                lda = (i1 - i0 - 1) if rank == bad_rank else m           # invalid on one rank only: rejected before anything collective
                try:
                    tmm.gemm(ctx, "N", "N", i1 - i0, j1 - j0, k, 1.0, ap.ctypes.data + i0 * 8, lda, bp.ctypes.data + j0 * k * 8, k, 0.0,
                             cp.ctypes.data + (j0 * m + i0) * 8, m, pin_host_buffers=False, copy_c_back=True)
                    ok = True
                except (RuntimeError, ValueError, MemoryError) as e:
                    ok = False
                    failed_calls.append((attempt, rank, str(e)[:60]))
                assert ok == (bad_rank is None), f"rank {rank} attempt {attempt}: call {'succeeded' if ok else 'failed'}"
            got = np.asarray(cp).reshape(n, m)
            assert np.array_equal(got[j0:j1, i0:i1], expect[j0:j1, i0:i1])
            barrier.wait()
            os.environ["TMM_EMUL_ALLOW_FREE_WHILE_IMPORTED"] = "1"
            barrier.wait()
            ctx.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(f"rank {rank}: {type(e).__name__}: {e}")
            barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert len(failed_calls) == 2 * world, failed_calls
    assert lib.emul_violations() == 0, lib.emul_first_violation().decode()
    print(f"EMUL_OK rank-fails-early world {world}: {len(failed_calls)} calls given up together, next call exact")


def run_dry(n_dev):
    """Full-size walk through the real scheduler with address-only memory (TMM_EMUL_DRY=1): BASELINE configs[3] and [4] and a C that
    needs super-blocks.  No arithmetic, no data movement - bounds, 64-bit offsets, ordering, protocol progress and byte counts."""
    assert lib.emul_dry_run() == 1 and tmm.device_count() == n_dev
    pr, pc = tmm.grid_shape(n_dev)
    A, B, C = 0x200000000000, 0x300000000000, 0x400000000000      # host "buffers": never dereferenced in a dry run
    cases = [
        ("headline dgemm 10000^3 beta=1 (four column stripes, staggered C upload)", np.float64, "NN", 10000, 10000, 10000, 1.0),
        ("C5 dgemm 100000^3", np.float64, "NN", 100000, 100000, 100000, 0.0),
        ("C4 zgemm 20000x20000x500000", np.complex128, "NN", 20000, 20000, 500000, 0.0),
        ("C4 zgemm CN beta=1", np.complex128, "CN", 20000, 20000, 500000, 1.0),
        ("dgemm TN 150000x130000x30000 (C alone exceeds one HBM)", np.float64, "TN", 150000, 130000, 30000, 1.0),
    ]
    for name, dtype, tt, m, n, k, beta in cases:
        es = np.dtype(dtype).itemsize
        ta, tb = tt
        lda = (m if ta == "N" else k) + 8
        ldb = (k if tb == "N" else n) + 8
        ldc = m + 8
        with tmm.make_context(dtype, 2, 5000, 5000, 5000) as ctx:
            if n_dev > 1:
                ctx.set_devices(n_dev)
            tmm.gemm(ctx, ta, tb, m, n, k, 1.0, A, lda, B, ldb, beta, C, ldc, pin_host_buffers=False, copy_c_back=True)
            st = ctx.last_stats()
            abc = es * (m * k + k * n + (m * n if beta != 0 else 0))
            assert st.d2h_bytes == es * m * n, (name, st.d2h_bytes)
            if st.regime == 0:    # resident on every GPU: each element crosses PCIe once over the whole grid
                assert st.h2d_bytes == abc, (name, st.h2d_bytes, abc)
                assert st.peer_bytes == es * (m * k * (pc - 1) + k * n * (pr - 1)), (name, st.peer_bytes)
            else:                 # streaming: panels are re-sent once per row / column of C super-blocks
                assert st.h2d_bytes >= abc and st.h2d_bytes % es == 0, (name, st.h2d_bytes, abc)
            print(f"  dry {name} on {pr}x{pc}: regime {st.regime}, {st.kernel_launches} launches, {st.c_blocks} C blocks, {st.k_chunks} k-chunks, "
                  f"H2D {st.h2d_bytes / 1e9:.1f} GB, D2H {st.d2h_bytes / 1e9:.1f} GB, NVLink {st.peer_bytes / 1e9:.1f} GB", flush=True)
        assert lib.emul_violations() == 0, f"{name}: {lib.emul_first_violation().decode()}"
        assert lib.emul_tma_contract_violations() == 0, name
        assert lib.emul_races() == 0, f"{name}: {lib.emul_first_race().decode()}"
    for d in range(n_dev):
        assert lib.emul_live_device_bytes(d) == 0
    print(f"EMUL_OK dry {pr}x{pc}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "ranks":
        run_ranks(int(sys.argv[2]))
    elif mode == "rankfails":
        run_rank_fails_early(int(sys.argv[2]))
    elif mode == "gridfaults":
        run_grid_faults(int(sys.argv[2]))
    elif mode == "faults":
        run_faults()
    elif mode == "auto":
        run_auto(int(sys.argv[2]))
    elif mode == "drysweep":
        run_drysweep(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif mode == "threads":
        run_threads(int(sys.argv[2]), int(sys.argv[3]))
    elif mode == "replan":
        run_replan()
    elif mode == "dry":
        run_dry(int(sys.argv[2]))
    elif mode == "sweep":
        run_sweep(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif mode == "single":
        run_single()
    else:
        run_grid(int(sys.argv[2]), sys.argv[3])
