"""One process per GPU (torchrun): C tile-blocks over a p_r x p_c grid of ranks.

torch.distributed is plumbing only: it carries the 128-byte NCCL ids from the first rank of every grid row / column to
its peers.  Everything on the data path - each rank's H2D share of a shared panel, the all-gather of the shares over NVLink,
the DMMA kernels, the D2H of the finished C blocks - is enqueued by the C++ scheduler (csrc/tmm_dist.cu, tmm_context.cu).

    grid = GridGemm(ctx, torch.distributed)            # collective: every rank
    grid.gemm('N', 'N', m_blk, n_blk, k, alpha, a_panel, ld_a, b_panel, ld_b, beta, c_blk, ld_c)

where a_panel holds this rank's rows of op(A) (full k), b_panel its columns of op(B), c_blk its block of C.
"""
from __future__ import annotations

import numpy as np

from . import gemm as _gemm, grid_shape, share_range, dist_unique_id


def grid_position(rank: int, world: int):
    """(grid_rows, grid_cols, my_row, my_col): ranks are laid out row-major over the grid."""
    pr, pc = grid_shape(world)
    return pr, pc, rank // pc, rank % pc


def block_of(rank: int, world: int, m: int, n: int):
    """This rank's block of an m x n C: (i_lo, i_hi, j_lo, j_hi)."""
    pr, pc, row, col = grid_position(rank, world)
    i0, i1 = share_range(m, pr, row)
    j0, j1 = share_range(n, pc, col)
    return i0, i1, j0, j1


def exchange_ids(dist, rank: int, world: int, make_id=dist_unique_id):
    """Collective.  Returns (row_id, col_id) for this rank: the id made by the first rank of its grid row / grid column.
    Works over any torch.distributed backend (gloo on CPU for tests, nccl on the box)."""
    pr, pc, row, col = grid_position(rank, world)
    mine = {}
    if pc > 1 and col == 0:
        mine[("row", row)] = make_id()
    if pr > 1 and row == 0:
        mine[("col", col)] = make_id()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ids = {}
    for d in gathered:
        ids.update(d)
    return ids.get(("row", row)), ids.get(("col", col))


class GridGemm:
    def __init__(self, ctx, dist):
        self.ctx, self.dist = ctx, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.pr, self.pc, self.row, self.col = grid_position(self.rank, self.world)
        row_id, col_id = exchange_ids(dist, self.rank, self.world)
        ctx.attach_grid(self.pr, self.pc, self.row, self.col, row_id, col_id)

    def gemm(self, trans_a, trans_b, m_blk, n_blk, k, alpha, a_panel, ld_a, b_panel, ld_b, beta, c_blk, ld_c, pin_host_buffers=False, copy_c_back=True):
        _gemm(self.ctx, trans_a, trans_b, m_blk, n_blk, k, alpha, a_panel, ld_a, b_panel, ld_b, beta, c_blk, ld_c,
              pin_host_buffers=pin_host_buffers, copy_c_back=copy_c_back)
