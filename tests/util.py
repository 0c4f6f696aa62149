"""Shared helpers for the GPU parity tests: oracle <-> product conversions and comparisons."""
import numpy as np

import matrel_b200 as mb
from oracle import matrel_oracle as O

# north_star tolerance: 1e-5 relative fp64.  The native DMMA path is plain fp64 FMA arithmetic, so the
# tests additionally hold it to a much tighter bound to catch indexing/race bugs that a loose bound hides.
REL_TOL = 1e-5
TIGHT_TOL = 1e-11


def to_product_matrix(m):
    if isinstance(m, O.DenseMatrix):
        return mb.DenseMatrix(m.numRows, m.numCols, m.values, m.isTransposed)
    return mb.SparseMatrix(m.numRows, m.numCols, m.colPtrs, m.rowIndices, m.values, m.isTransposed)


def to_dataset(session, block_dict):
    return session.createDataset(mb.MatrixBlock(i, j, to_product_matrix(m)) for (i, j), m in block_dict.items())


def from_dataset(ds):
    return {(b.rid, b.cid): b.matrix for b in ds.collect()}


def rel_err(got, want):
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    if scale == 0.0:
        scale = 1.0
    return float(np.max(np.abs(got - want))) / scale if want.size else 0.0


def assert_same_dataset(got, want, tol=TIGHT_TOL, exact_storage=False):
    """Block ids, block presence, types, shapes and isTransposed flags must match exactly;
    values within `tol` relative to the block's max magnitude."""
    assert sorted(got) == sorted(want), (sorted(got), sorted(want))
    for key, w in want.items():
        g = got[key]
        assert isinstance(g, mb.SparseMatrix) == isinstance(w, O.SparseMatrix), (key, g, w)
        assert (g.numRows, g.numCols, g.isTransposed) == (w.numRows, w.numCols, w.isTransposed), (key, g, w)
        if isinstance(w, O.SparseMatrix):
            assert g.colPtrs.tolist() == w.colPtrs.tolist() and g.rowIndices.tolist() == w.rowIndices.tolist()
        a, b = g.values, w.values
        assert a.shape == b.shape
        if exact_storage:
            assert np.array_equal(a, b, equal_nan=True), key
        else:
            finite = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), finite), key
            assert np.array_equal(a[~finite], b[~finite], equal_nan=True), key
            assert rel_err(a[finite], b[finite]) <= tol, (key, rel_err(a[finite], b[finite]))


def random_block_dataset(rng, nrows, ncols, blk, density=1.0, p_transposed=0.0, p_sparse=0.0, sparse_density=0.05,
                         lo=-1.0, hi=1.0):
    """Random block dict: each block present with prob `density`, row-major with prob p_transposed,
    sparse (CSC or CSR) with prob p_sparse."""
    ds = {}
    for i in range(-(-nrows // blk)):
        for j in range(-(-ncols // blk)):
            if rng.random() >= density:
                continue
            r, c = min(blk, nrows - i * blk), min(blk, ncols - j * blk)
            a = rng.uniform(lo, hi, (r, c))
            if rng.random() < p_sparse:
                a = a * (rng.random((r, c)) < sparse_density)
                csc = O.DenseMatrix(r, c, np.ascontiguousarray(a.T).reshape(-1)).toSparse()
                if rng.random() < 0.5:
                    ds[(i, j)] = csc
                else:   # CSR of the same matrix
                    ds[(i, j)] = O.DenseMatrix(c, r, np.ascontiguousarray(a).reshape(-1)).toSparse().transpose()
            elif rng.random() < p_transposed:
                ds[(i, j)] = O.DenseMatrix(r, c, np.ascontiguousarray(a).reshape(-1), True)
            else:
                ds[(i, j)] = O.DenseMatrix(r, c, np.ascontiguousarray(a.T).reshape(-1), False)
    return ds
