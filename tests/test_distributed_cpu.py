"""CPU (gloo) coverage of the N > 1 path: placement arithmetic and the panel exchange."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grid_plan_arithmetic():
    from matrel_b200.distributed import GridPlan, grid_shape, panel_blocks_A, panel_blocks_B
    assert [grid_shape(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    for world in (1, 2, 4, 8, 6):
        for (n, m, blk) in [(16384, 16384, 1024), (1000, 700, 128), (4096, 4096, 512), (130, 50, 64)]:
            p = GridPlan(world, n, m, blk)
            seen = {}
            for i in range(p.nbr):
                for j in range(p.nbc):
                    o = p.owner(i, j)
                    assert 0 <= o < world
                    key = (o, p.slot(i, j))
                    assert key not in seen and p.slot(i, j) < p.local_slots
                    seen[key] = (i, j)
                    # the reference's Row/ColumnPartitioner arithmetic
                    assert p.coords(o) == (i % p.pr, j % p.pc)
            assert sum(len(p.owned(r)) for r in range(world)) == p.nbr * p.nbc
            for r in range(world):
                rows = {i for i, _ in p.owned(r)}
                assert {i for i, _, _, _ in panel_blocks_A(p, r)} == rows or not p.owned(r)
                for i, k, src, slot in panel_blocks_A(p, r):
                    assert p.row_group_ranks(r)[src] == p.owner(i, k) and slot == p.slot(i, k)
                for k, j, src, slot in panel_blocks_B(p, r):
                    assert p.col_group_ranks(r)[src] == p.owner(k, j) and slot == p.slot(k, j)


@pytest.mark.parametrize("world", [2, 4])
def test_panel_exchange_gloo(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"OK world={world}" in r.stdout


def test_pull_pieces_partition_the_block_rows_and_columns():
    """pull_chunks mirrors mr_grid_multiply's cut: piece ch of A = local block rows [rows ch / n, rows (ch + 1) / n), piece ch of B
    the local block columns cut the same way; every owned row / column lands in exactly one piece, in ascending order."""
    from matrel_b200.distributed import GridPlan, pull_chunks
    for world, n, k, m, blk in [(1, 512, 512, 512, 128), (2, 1024, 512, 768, 128), (4, 2048, 1024, 1024, 256), (8, 16384, 16384, 16384, 1024),
                                (8, 640, 384, 896, 128)]:
        planA, planB = GridPlan(world, n, k, blk), GridPlan(world, k, m, blk)
        for rank in range(world):
            r, c = planA.coords(rank)
            for req in (1, 3, 4, 100):
                nch, rows, cols = pull_chunks(planA, planB, rank, req)
                assert nch == min(req, 64) and len(rows) == len(cols) == nch
                assert [i for piece in rows for i in piece] == list(range(r, planA.nbr, planA.pr))
                assert [j for piece in cols for j in piece] == list(range(c, planB.nbc, planB.pc))
                my_rows = len(range(r, planA.nbr, planA.pr))
                for ch, piece in enumerate(rows):   # the C side's integer cut
                    assert len(piece) == my_rows * (ch + 1) // nch - my_rows * ch // nch


def test_repartition_routes_are_a_permutation():
    """Every block of the source grid is sent exactly once, to its owner on the target grid, into the slot that owner expects; the
    receive lists name the same slots in the same (block id) order, so one packed message per peer suffices."""
    from matrel_b200.distributed import GridPlan, repartition_routes
    world, n, m, blk = 8, 1280, 896, 128
    src, dst = GridPlan(world, n, m, blk, 2, 4), GridPlan(world, n, m, blk, 4, 2)
    routes = [repartition_routes(src, dst, rank) for rank in range(world)]
    moved = 0
    for rank in range(world):
        sends, _ = routes[rank]
        owned = sorted(src.owned(rank))
        assert sorted(s_slot for lst in sends.values() for s_slot, _ in lst) == sorted(src.slot(i, j) for i, j in owned)
        for peer, lst in sends.items():
            want = [(src.slot(i, j), dst.slot(i, j)) for (i, j) in owned if dst.owner(i, j) == peer]
            assert lst == want
            assert routes[peer][1][rank] == [d for _, d in lst]        # the peer expects exactly these slots, in this order
            moved += len(lst)
    assert moved == src.nbr * src.nbc
    for rank in range(world):                                            # every target slot is filled exactly once
        _, recvs = routes[rank]
        assert sorted(d for lst in recvs.values() for d in lst) == sorted(dst.slot(i, j) for i, j in dst.owned(rank))
