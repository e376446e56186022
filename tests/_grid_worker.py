"""Worker for the multi-rank tests (launched by torch.distributed.run, one rank per process).

  mode cpu : gloo, no GPU - checks the host logic of the grid path: id exchange, block partition, upload shares reassemble
             to the shared panel, and the per-block products (oracle) tile the full product.
  mode gpu : nccl, one GPU per rank - every rank computes its C block through tmm_gemm on an attached grid; rank 0 gathers the
             blocks and compares with the oracle on the full matrices.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _util  # noqa: E402
import tiled_mm_b200 as tmm  # noqa: E402
from tiled_mm_b200 import multi_gpu  # noqa: E402


def full_problem(dtype, tt, m, n, k, pad, seed):
    ta, tb = tt
    ar, ac = _util.stored_shape(ta, m, k); br, bc = _util.stored_shape(tb, k, n)
    lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
    rng = np.random.default_rng(seed)
    def gen(count):
        v = rng.integers(0, 10, count).astype(np.float64)
        return (v + 1j * rng.integers(0, 10, count)).astype(dtype) if np.dtype(dtype).kind == "c" else v.astype(dtype)
    return gen(lda * ac), gen(ldb * bc), gen(ldc * n), lda, ldb, ldc


def panel_offsets(ta, tb, lda, ldb, i0, j0):
    """element offsets of this rank's A row-panel / B column-panel inside the full stored matrices"""
    return (i0 if ta == "N" else i0 * lda), (j0 * ldb if tb == "N" else j0)


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    else:
        dist.init_process_group("gloo")
    oracle = _util.Oracle()
    pr, pc, row, col = multi_gpu.grid_position(rank, world)
    cases = [(np.float64, "NN", 700, 900, 500, (0, 0, 0), 1.0, 0.0, None), (np.float64, "TN", 513, 300, 777, (3, 5, 7), 2.0, -1.0, None),
             (np.complex128, "CT", 301, 403, 209, (1, 2, 3), 1 - 2j, 2 + 1j, None), (np.float64, "NT", 1500, 1300, 2100, (1, 2, 3), 1.0, 1.0, 16 << 20)]
    if mode == "cpu":
        # 1. id exchange: the first rank of each row / column makes the id, everyone in that row / column receives the same one
        row_id, col_id = multi_gpu.exchange_ids(dist, rank, world, make_id=lambda: bytes([rank + 1]) * 128)
        assert (row_id is None) == (pc == 1) and (col_id is None) == (pr == 1)
        if row_id is not None:
            assert row_id == bytes([row * pc + 1]) * 128
        if col_id is not None:
            assert col_id == bytes([col + 1]) * 128
        row_groups = [dist.new_group([r * pc + g for g in range(pc)]) for r in range(pr)]  # collective: every rank creates every group
    else:
        ctx_by_dtype = {}
    for ci, (dtype, tt, m, n, k, pad, alpha, beta, budget) in enumerate(cases):
        ta, tb = tt
        a, b, c, lda, ldb, ldc = full_problem(dtype, tt, m, n, k, pad, seed=ci)
        expect = oracle.gemm(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c.copy(), ldc).reshape(n, ldc)
        i0, i1, j0, j1 = multi_gpu.block_of(rank, world, m, n)
        mb, nb = i1 - i0, j1 - j0
        oa, ob = panel_offsets(ta, tb, lda, ldb, i0, j0)
        if mode == "cpu":
            # 2. upload shares of the shared A panel reassemble to the panel (same split the device path uses: stored columns)
            ar, ac = _util.stored_shape(ta, mb, k)
            panel = a[oa:].copy()
            lo, hi = tmm.share_range(ac, pc, col)
            mine = np.zeros((-(-ac // pc), ar), dtype=dtype)
            for j in range(lo, hi):
                mine[j - lo] = panel[j * lda: j * lda + ar]
            t = torch.from_numpy(mine.view(np.float64).copy())
            out = [torch.empty_like(t) for _ in range(pc)]
            dist.all_gather(out, t, group=row_groups[row])
            rebuilt = np.zeros((ac, ar), dtype=dtype)
            for g in range(pc):
                glo, ghi = tmm.share_range(ac, pc, g)
                rebuilt[glo:ghi] = out[g].numpy().view(dtype)[: ghi - glo]
            want = np.stack([panel[j * lda: j * lda + ar] for j in range(ac)])
            assert np.array_equal(rebuilt, want), "shares do not reassemble to the A panel"
            # 3. my block through the oracle, called exactly like the device path is (panel pointers + full leading dimensions)
            cblk = c.copy()
            oc = j0 * ldc + i0
            got = oracle.gemm(ta, tb, mb, nb, k, alpha, a[oa:], lda, b[ob:], ldb, beta, cblk[oc:], ldc)
            got_blk = np.stack([got[j * ldc: j * ldc + mb] for j in range(nb)])
        else:
            key = np.dtype(dtype)
            if key not in ctx_by_dtype:
                ctx = tmm.make_context(dtype, 2, 512, 512, 512)
                ctx_by_dtype[key] = (ctx, multi_gpu.GridGemm(ctx, dist))
            ctx, grid = ctx_by_dtype[key]
            ctx.set_device_budget(budget or 0)
            ap = tmm.malloc_pinned(dtype, a.size); ap[:] = a
            bp = tmm.malloc_pinned(dtype, b.size); bp[:] = b
            cp = tmm.malloc_pinned(dtype, c.size); cp[:] = c
            es = np.dtype(dtype).itemsize
            grid.gemm(ta, tb, mb, nb, k, alpha, ap.ctypes.data + oa * es, lda, bp.ctypes.data + ob * es, ldb, beta,
                      cp.ctypes.data + (j0 * ldc + i0) * es, ldc)
            st = ctx.last_stats()
            assert st.regime == (1 if budget else 0), (st.regime, budget)
            got_full = np.asarray(cp).reshape(n, ldc)
            got_blk = got_full[j0:j1, i0:i1].copy()
            # nothing outside my block was written
            mask = np.ones((n, ldc), dtype=bool); mask[j0:j1, i0:i1] = False
            assert np.array_equal(got_full[mask], c.reshape(n, ldc)[mask]), "wrote outside this rank's C block"
            # every shared panel element crossed PCIe once across the grid: my share is 1/pc of A plus 1/pr of B (+ my C block if beta != 0)
            if not budget:
                ar, ac = _util.stored_shape(ta, mb, k); br, bc = _util.stored_shape(tb, k, nb)
                assert st.h2d_bytes <= es * (ar * -(-ac // pc) + br * -(-bc // pr) + (mb * nb if beta != 0 else 0)) + 64 * es * (ar + br), st.h2d_bytes
                if world > 1:
                    assert st.peer_bytes > 0
        assert np.array_equal(got_blk, expect[j0:j1, i0:i1]), f"rank {rank} case {ci} {tt}: block differs from the oracle"
    dist.barrier()
    if rank == 0:
        print(f"GRID_OK mode={mode} world={world} grid={pr}x{pc}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
