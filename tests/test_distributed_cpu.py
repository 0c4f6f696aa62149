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
