"""The single-process multi-GPU entry points of the C ABI (mr_init_grid / mr_dmatrix_*) through ctypes, against the oracle.
Runs on the 1 x 1 grid on a one-GPU box and on every grid the box can hold (2, 4, 8 GPUs) otherwise."""
import ctypes as C

import numpy as np
import pytest

from matrel_b200 import _native as N
from oracle import matrel_oracle as O

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _put(dm, rid, cid, m):
    d = N.mr_block_desc()
    d.type, d.numRows, d.numCols, d.isTransposed = 1, m.numRows, m.numCols, 1 if m.isTransposed else 0
    d.values = m.values.ctypes.data_as(C.POINTER(C.c_double))
    d.valuesLen = m.values.size
    N.check(N.lib.mr_dmatrix_put_block(dm, rid, cid, C.byref(d)))


def _get(dm, rid, cid):
    d = N.mr_block_desc()
    N.check(N.lib.mr_dmatrix_get_block(dm, rid, cid, C.byref(d)))
    v = np.empty(d.valuesLen)
    d.values = v.ctypes.data_as(C.POINTER(C.c_double))
    N.check(N.lib.mr_dmatrix_get_block(dm, rid, cid, C.byref(d)))
    assert d.type == 1 and not d.isTransposed
    return v.reshape(d.numCols, d.numRows).T


def _assemble(dm, n, m, blk):
    out = np.zeros((n, m))
    for i in range(-(-n // blk)):
        for j in range(-(-m // blk)):
            has = C.c_int32()
            N.check(N.lib.mr_dmatrix_has_block(dm, i, j, C.byref(has)))
            if has.value:
                b = _get(dm, i, j)
                out[i * blk:i * blk + b.shape[0], j * blk:j * blk + b.shape[1]] = b
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("algo,n,k,m,blk", [(1, 7 * 64 - 6, 5 * 64, 6 * 64 - 10, 64), (4, 1024, 768, 1280, 128), (0, 2048, 2048, 2048, 256)])
def test_grid_multiply_matches_oracle(world, algo, n, k, m, blk):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    opts = N.mr_options(-1, 1, algo, 0, None)
    g = C.c_void_p()
    N.check(N.lib.mr_init_grid(C.byref(opts), world, C.byref(g)))
    try:
        A, B, Cm = C.c_void_p(), C.c_void_p(), C.c_void_p()
        if algo == 0:   # device-generated operands (java.util.Random streams), checked against the oracle's generator
            N.check(N.lib.mr_dmatrix_rand(g, n, k, blk, 42, C.byref(A)))
            N.check(N.lib.mr_dmatrix_rand(g, k, m, blk, 43, C.byref(B)))
            Ao, Bo = O.rand_dense_dataset(n, k, blk, 42), O.rand_dense_dataset(k, m, blk, 43)
        else:
            rng = np.random.default_rng(world * 100 + algo)
            Ao = {(i, j): O.DenseMatrix(min(blk, n - i * blk), min(blk, k - j * blk), rng.uniform(-1, 1, min(blk, n - i * blk) * min(blk, k - j * blk)))
                  for i in range(-(-n // blk)) for j in range(-(-k // blk))}
            Bo = {(i, j): O.DenseMatrix(min(blk, k - i * blk), min(blk, m - j * blk), rng.uniform(-1, 1, min(blk, k - i * blk) * min(blk, m - j * blk)))
                  for i in range(-(-k // blk)) for j in range(-(-m // blk))}
            N.check(N.lib.mr_dmatrix_create(g, n, k, blk, C.byref(A)))
            N.check(N.lib.mr_dmatrix_create(g, k, m, blk, C.byref(B)))
            for (i, j), blkm in Ao.items():
                _put(A, i, j, blkm)
            for (i, j), blkm in Bo.items():
                _put(B, i, j, blkm)
        N.check(N.lib.mr_dmatrix_multiply(A, B, C.byref(Cm)))
        got = _assemble(Cm, n, m, blk)
        want = O.assemble(O.matrix_multiply(Ao, n, k, Bo, k, m, blk), n, m, blk)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
        nblocks = C.c_int64()
        N.check(N.lib.mr_dmatrix_num_blocks(Cm, C.byref(nblocks)))
        assert nblocks.value == (-(-n // blk)) * (-(-m // blk))
        # placement: the reference's RowPartitioner x ColumnPartitioner arithmetic on the grid
        pr, pc = C.c_int32(), C.c_int32()
        N.check(N.lib.mr_grid_info(g, None, C.byref(pr), C.byref(pc), None))
        owner = C.c_int32()
        N.check(N.lib.mr_dmatrix_owner(Cm, 3, 2, C.byref(owner)))
        assert owner.value == (3 % pr.value) * pc.value + (2 % pc.value)
        # (A B) on operands that start on DIFFERENT layouts: re-partition B to the column layout and back (NCCL all-to-all)
        if world > 1:
            Bc, Bg, C2 = C.c_void_p(), C.c_void_p(), C.c_void_p()
            N.check(N.lib.mr_dmatrix_repartition(B, 1, world, C.byref(Bc)))
            N.check(N.lib.mr_dmatrix_repartition(Bc, pr.value, pc.value, C.byref(Bg)))
            N.check(N.lib.mr_dmatrix_multiply(A, Bg, C.byref(C2)))
            assert np.array_equal(_assemble(C2, n, m, blk), got)
            total = C.c_double()
            N.check(N.lib.mr_dmatrix_reduce_scalar(Cm, 0, C.byref(total)))
            assert abs(total.value - want.sum()) <= 1e-10 * np.abs(want).sum()
            for h in (Bc, Bg, C2):
                N.check(N.lib.mr_dmatrix_free(h))
        for h in (A, B, Cm):
            N.check(N.lib.mr_dmatrix_free(h))
    finally:
        N.check(N.lib.mr_grid_shutdown(g))


def test_sharded_put_block_rejects_foreign_blocks(session):
    from matrel_b200.dataset import create_sharded
    import matrel_b200 as mb
    ds = create_sharded(session, 256, 256, 64, 2, 2, 1, 0)           # rank (1, 0) of a 2 x 2 grid
    assert sorted(ds.block_ids()) == [(1, 0), (1, 2), (3, 0), (3, 2)]
    ds.put_block(1, 2, mb.DenseMatrix(64, 64, np.arange(4096.0)))
    assert np.array_equal(ds.get_block(1, 2).values, np.arange(4096.0))
    assert not ds.get_block(3, 0).values.any()                         # never written: zeros
    with pytest.raises(mb.IllegalArgumentException, match="belongs to rank"):
        ds.put_block(0, 0, mb.DenseMatrix(64, 64, np.zeros(4096)))
    with pytest.raises(mb.IllegalArgumentException, match="layout expects"):
        ds.put_block(1, 0, mb.DenseMatrix(32, 64, np.zeros(2048)))


def _get_any(dm, rid, cid):
    """(numRows, numCols, isTransposed, values) of a block, whatever its layout flag."""
    d = N.mr_block_desc()
    N.check(N.lib.mr_dmatrix_get_block(dm, rid, cid, C.byref(d)))
    v = np.empty(d.valuesLen)
    d.values = v.ctypes.data_as(C.POINTER(C.c_double))
    N.check(N.lib.mr_dmatrix_get_block(dm, rid, cid, C.byref(d)))
    assert d.type == 1
    return d.numRows, d.numCols, bool(d.isTransposed), v


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_grid_transpose_and_scalar_ops_match_oracle(world):
    """Dataset.t and the scalar maps on the single-process grid against the oracle (O.transpose = key swap + flag flip,
    MatfastExecution.scala:215-236): ids, shapes, flags and payloads exactly; the owner of block (j, i) of A^T is the reference's
    Row x ColumnPartitioner arithmetic on the same grid; a transposed operand feeds the multiply."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    n, m, blk = 5 * 64 - 3, 3 * 64 + 10, 64
    rng = np.random.default_rng(7 + world)
    Ao = {(i, j): O.DenseMatrix(min(blk, n - i * blk), min(blk, m - j * blk), rng.uniform(-1, 1, min(blk, n - i * blk) * min(blk, m - j * blk)))
          for i in range(-(-n // blk)) for j in range(-(-m // blk))}
    opts = N.mr_options(-1, 1, 1, 0, None)
    g = C.c_void_p()
    N.check(N.lib.mr_init_grid(C.byref(opts), world, C.byref(g)))
    try:
        pr, pc = C.c_int32(), C.c_int32()
        N.check(N.lib.mr_grid_info(g, None, C.byref(pr), C.byref(pc), None))
        A, At, G, S = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(N.lib.mr_dmatrix_create(g, n, m, blk, C.byref(A)))
        for (i, j), b in Ao.items():
            _put(A, i, j, b)
        N.check(N.lib.mr_dmatrix_transpose(A, C.byref(At)))
        want = O.transpose(Ao)
        for (i, j), w in want.items():
            nr, nc, isT, v = _get_any(At, i, j)
            assert (nr, nc, isT) == (w.numRows, w.numCols, w.isTransposed) and np.array_equal(v, w.values)
            owner = C.c_int32()
            N.check(N.lib.mr_dmatrix_owner(At, i, j, C.byref(owner)))
            assert owner.value == (i % pr.value) * pc.value + (j % pc.value)
        nblocks = C.c_int64()
        N.check(N.lib.mr_dmatrix_num_blocks(At, C.byref(nblocks)))
        assert nblocks.value == len(want)
        # A^T A through the grid multiply (the row-major slabs are normalised on the device)
        N.check(N.lib.mr_dmatrix_multiply(At, A, C.byref(G)))
        wantG = O.assemble(O.matrix_multiply(want, m, n, Ao, n, m, blk), m, m, blk)
        assert np.max(np.abs(_assemble(G, m, m, blk) - wantG)) <= 1e-12 * np.max(np.abs(wantG))
        # scalar maps keep ids, placement and layout flags: (A^T * 3 + 0.25) ^ 2
        T1, T2 = C.c_void_p(), C.c_void_p()
        N.check(N.lib.mr_dmatrix_scalar(1, At, 3.0, C.byref(T1)))
        N.check(N.lib.mr_dmatrix_scalar(0, T1, 0.25, C.byref(T2)))
        N.check(N.lib.mr_dmatrix_scalar(2, T2, 2.0, C.byref(S)))
        wantS = O.power(O.add_scalar(O.multiply_scalar(want, 3.0), 0.25), 2.0)
        for (i, j), w in wantS.items():
            nr, nc, isT, v = _get_any(S, i, j)
            assert (nr, nc, isT) == (w.numRows, w.numCols, w.isTransposed)
            np.testing.assert_allclose(v, w.values, rtol=1e-14, atol=0)
        # rowSum / colSum: block ids (i, 0) / (0, j), shapes and values against the oracle; owners under the operand's placement
        Rs, Cs = C.c_void_p(), C.c_void_p()
        N.check(N.lib.mr_dmatrix_axis_sum(A, 0, C.byref(Rs)))
        N.check(N.lib.mr_dmatrix_axis_sum(At, 1, C.byref(Cs)))      # colSum of A^T = (rowSum of A)^T, from row-major blocks
        wr, wc = O.row_sum(Ao, n, m), O.col_sum(want, m, n)
        for dm, w in ((Rs, wr), (Cs, wc)):
            N.check(N.lib.mr_dmatrix_num_blocks(dm, C.byref(nblocks)))
            assert nblocks.value == len(w)
            for (i, j), blkm in w.items():
                nr, nc, isT, v = _get_any(dm, i, j)
                assert (nr, nc) == (blkm.numRows, blkm.numCols)
                np.testing.assert_allclose(v, blkm.values, rtol=0, atol=1e-12 * m)
                owner = C.c_int32()
                N.check(N.lib.mr_dmatrix_owner(dm, i, j, C.byref(owner)))
                assert owner.value == (i % pr.value) * pc.value + (j % pc.value)
        # project / selection re-key the pieces to (0, cid) / (rid, 0) / (0, 0) and place them at their owners
        Pr, Pc, Se = C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(N.lib.mr_dmatrix_project(A, 1, 130, C.byref(Pr)))
        N.check(N.lib.mr_dmatrix_project(At, 0, 131, C.byref(Pc)))   # column 131 of A^T = row 131 of A, from row-major blocks
        N.check(N.lib.mr_dmatrix_selection(A, n - 1, m - 1, C.byref(Se)))
        for dm, w in ((Pr, O.project(Ao, n, m, blk, True, 130)), (Pc, O.project(want, m, n, blk, False, 131)),
                      (Se, O.selection(Ao, n, m, blk, n - 1, m - 1))):
            N.check(N.lib.mr_dmatrix_num_blocks(dm, C.byref(nblocks)))
            assert nblocks.value == len(w) > 0
            for (i, j), blkm in w.items():
                nr, nc, isT, v = _get_any(dm, i, j)
                assert (nr, nc) == (blkm.numRows, blkm.numCols) and np.array_equal(v, blkm.to_numpy().ravel(order="F"))
        bad = C.c_void_p()
        assert N.lib.mr_dmatrix_project(A, 0, m, C.byref(bad)) == N.MR_EINVAL
        assert "col index should be smaller than #cols" in N.last_error()
        for h in (A, At, G, S, T1, T2, Rs, Cs, Pr, Pc, Se):
            N.check(N.lib.mr_dmatrix_free(h))
    finally:
        N.check(N.lib.mr_grid_shutdown(g))
