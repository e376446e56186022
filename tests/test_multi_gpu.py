"""Multi-GPU path: host logic on CPU (gloo, world_size 2 and 4), device path on a box with >= 2 GPUs."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import _util

ROOT = Path(__file__).resolve().parent.parent
WORKER = str(ROOT / "tests" / "_grid_worker.py")


def _launch(world, mode, port):
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           WORKER, mode]
    r = subprocess.run(cmd, cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GRID_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


def test_grid_shape_and_shares(tmm):
    assert [tmm.grid_shape(n) for n in (1, 2, 3, 4, 6, 8)] == [(1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (2, 4)]
    for extent, parts in [(10, 3), (7, 8), (100000, 8), (0, 2), (5, 1)]:
        rs = [tmm.share_range(extent, parts, g) for g in range(parts)]
        assert rs[0][0] == 0 and rs[-1][1] == extent
        assert all(rs[g][1] == rs[g + 1][0] for g in range(parts - 1))
        sizes = [hi - lo for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_block_partition_covers_c(tmm):
    from tiled_mm_b200 import multi_gpu
    for world in (1, 2, 4, 8):
        cover = np.zeros((37, 53), dtype=int)
        for r in range(world):
            i0, i1, j0, j1 = multi_gpu.block_of(r, world, 37, 53)
            cover[i0:i1, j0:j1] += 1
        assert (cover == 1).all()


@pytest.mark.parametrize("world", [2, 4])
def test_grid_host_logic_gloo(tmm, world):
    _launch(world, "cpu", 29621 + world)


@pytest.mark.gpu
def test_single_process_multi_device(gpu_tmm, oracle):
    """gpu::gemm drop-in over several GPUs of the box: tmm_context_set_devices + one tmm_gemm."""
    tmm = gpu_tmm
    ndev = tmm.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    n_use = 4 if ndev >= 4 else 2
    rng = np.random.default_rng(11)
    for dtype, tt, (m, n, k), pad, alpha, beta, budget in [
        (np.float64, "NN", (1000, 1000, 1000), (0, 0, 0), 1.0, 1.0, 0), (np.float64, "TN", (1234, 777, 1357), (3, 5, 7), 2.0, -1.0, 0),
        (np.complex128, "CT", (301, 403, 209), (1, 2, 3), 1 - 2j, 2 + 1j, 0), (np.float64, "NT", (1500, 1300, 2100), (1, 2, 3), 1.0, 0.0, 16 << 20),
        (np.float32, "NN", (300, 200, 150), (0, 0, 0), 1.0, 1.0, 0), (np.float64, "NN", (1, 3, 5), (0, 0, 0), 1.0, 0.0, 0),
    ]:
        ta, tb = tt
        ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
        lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
        def gen(count):
            v = rng.integers(0, 10, count).astype(np.float64)
            return (v + 1j * rng.integers(0, 10, count)).astype(dtype) if np.dtype(dtype).kind == "c" else v.astype(dtype)
        a0, b0, c0 = gen(lda * ac), gen(ldb * bc), gen(ldc * n)
        expect = oracle.gemm(ta, tb, m, n, k, alpha, a0, lda, b0, ldb, beta, c0.copy(), ldc)
        a = tmm.malloc_pinned(dtype, a0.size); a[:] = a0
        b = tmm.malloc_pinned(dtype, b0.size); b[:] = b0
        c = tmm.malloc_pinned(dtype, c0.size); c[:] = c0
        with tmm.make_context(dtype, 2, 512, 512, 512) as ctx:
            if budget:
                ctx.set_device_budget(budget)
            ctx.set_devices(n_use)
            assert ctx.num_devices() == n_use
            tmm.gemm(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers=False, copy_c_back=True)
            st = ctx.last_stats()
            assert np.array_equal(np.asarray(c), expect), (np.dtype(dtype), tt, m, n, k)
            if m >= 64 and not budget:
                es = np.dtype(dtype).itemsize
                # each shared panel element crosses PCIe exactly once over the whole grid
                assert st.h2d_bytes == es * (m * k + k * n + (m * n if beta != 0 else 0)), st.h2d_bytes
                assert st.peer_bytes > 0
            # pageable buffers + pin_host_buffers=True through the parent
            c2 = c0.copy()
            tmm.gemm(ctx, ta, tb, m, n, k, alpha, a0, lda, b0, ldb, beta, c2, ldc, pin_host_buffers=True, copy_c_back=True)
            assert np.array_equal(c2, expect)


@pytest.mark.gpu
def test_one_process_per_gpu_nccl(gpu_tmm):
    ndev = gpu_tmm.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    _launch(4 if ndev >= 4 else 2, "gpu", 29671)
