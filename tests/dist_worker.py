"""Worker for tests/test_distributed_cpu.py: world_size ranks over gloo on CPU.
Checks that the grid plan + panel all-gather deliver to every rank exactly the A row-panels and
B column-panels its C blocks need (the host-side logic of matrel_b200/distributed.py); the block
products themselves are done with the oracle here because there is no GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matrel_b200.distributed import (GridGroups, GridPlan, exchange_repartition, exchange_transpose, gather_panels,  # noqa: E402
                                     panel_blocks_A, panel_blocks_B, repartition_routes, transpose_routes)
from oracle import matrel_oracle as O  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, k, m, blk = 7 * 16 - 3, 5 * 16, 6 * 16 - 5, 16          # ragged edges, rectangular
    planA, planB, planC = GridPlan(world, n, k, blk), GridPlan(world, k, m, blk), GridPlan(world, n, m, blk)
    groups = GridGroups(planA, rank)
    A = O.rand_dense_dataset(n, k, blk, 42)
    B = O.rand_dense_dataset(k, m, blk, 43)

    def local_slab(ds, plan):
        slab = torch.zeros((plan.local_slots, plan.slot_elems), dtype=torch.float64)
        for (i, j) in plan.owned(rank):
            v = ds[(i, j)].values
            slab[plan.slot(i, j), :v.size] = torch.from_numpy(v)
        return slab

    gA = gather_panels(local_slab(A, planA), groups.row_group, planA.pc)
    gB = gather_panels(local_slab(B, planB), groups.col_group, planA.pr)
    pa, pb = {}, {}
    for i, kk, src, slot in panel_blocks_A(planA, rank):
        r, c = planA.block_shape(i, kk)
        pa[(i, kk)] = O.DenseMatrix(r, c, gA[src, slot, :r * c].numpy().copy())
    for kk, j, src, slot in panel_blocks_B(planB, rank):
        r, c = planB.block_shape(kk, j)
        pb[(kk, j)] = O.DenseMatrix(r, c, gB[src, slot, :r * c].numpy().copy())
    # every gathered block is bit-identical to the source block
    for key, blkm in pa.items():
        assert np.array_equal(blkm.values, A[key].values), ("A", key)
    for key, blkm in pb.items():
        assert np.array_equal(blkm.values, B[key].values), ("B", key)
    mine = O.matrix_multiply(pa, n, k, pb, k, m, blk)
    assert sorted(mine) == sorted(planC.owned(rank)), (sorted(mine), planC.owned(rank))
    full = O.matrix_multiply(A, n, k, B, k, m, blk)
    for key, blkm in mine.items():
        assert planC.owner(*key) == rank
        assert np.array_equal(blkm.values, full[key].values), key
    # union over ranks covers the whole product exactly once
    cnt = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(cnt)
    assert int(cnt.item()) == len(full)
    # transpose on the grid: every block (j, i) of A lands, untouched, in the slot of block (i, j) of A^T on its new owner
    slabT, planT = exchange_transpose(local_slab(A, planA), planA, rank)
    assert (planT.nrows, planT.ncols) == (k, n)
    AT = O.transpose(A)
    for (i, j) in planT.owned(rank):
        v = AT[(i, j)].values
        assert np.array_equal(slabT[planT.slot(i, j), :v.size].numpy(), v), ("A^T", i, j)
    sends, recvs = transpose_routes(planA, rank)
    assert sum(len(v) for v in sends.values()) == len(planA.owned(rank))
    assert sum(len(v) for v in recvs.values()) == len(planT.owned(rank))
    # re-partitioning (repartitionWithTargetPartitioner): grid -> RowPartitioner layout (P x 1) -> ColumnPartitioner layout (1 x P) -> grid;
    # every block arrives untouched at its owner under each layout, and the round trip is the identity
    slab0 = local_slab(A, planA)
    cur, cur_plan = slab0, planA
    for (pr_, pc_) in [(world, 1), (1, world), (planA.pr, planA.pc)]:
        nxt_plan = GridPlan(world, n, k, blk, pr_, pc_)
        sends, recvs = repartition_routes(cur_plan, nxt_plan, rank)
        assert sum(len(v) for v in sends.values()) == len(cur_plan.owned(rank))
        assert sum(len(v) for v in recvs.values()) == len(nxt_plan.owned(rank))
        nxt = exchange_repartition(cur, cur_plan, nxt_plan, rank)
        for (i, j) in nxt_plan.owned(rank):
            assert nxt_plan.coords(rank) == (i % pr_, j % pc_)
            v = A[(i, j)].values
            assert np.array_equal(nxt[nxt_plan.slot(i, j), :v.size].numpy(), v), ("repartition", pr_, pc_, i, j)
        cur, cur_plan = nxt, nxt_plan
    assert torch.equal(cur, slab0)
    dist.barrier()
    if rank == 0:
        print(f"OK world={world} grid={planA.pr}x{planA.pc} blocks={len(full)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
