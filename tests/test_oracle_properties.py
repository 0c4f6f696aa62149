"""Hypothesis property tests of the CPU oracle (oracle plan item 4): random block-presence
patterns, random storage flags (row-/column-major, CSC/CSR), ragged edge blocks.  The blocked restatement of
`matrixMultiplyGeneral` and of the element-wise joins must agree with plain dense numpy on the assembled matrices,
and the structural rules of the reference (presence, output layout, serializer struct) must hold."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import matrel_oracle as O
from tests.util import random_block_dataset

SET = dict(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
dims = st.tuples(st.integers(1, 23), st.integers(1, 23), st.integers(1, 23), st.integers(2, 9))


@settings(**SET)
@given(d=dims, seed=st.integers(0, 2 ** 20), dens=st.sampled_from([1.0, 0.7, 0.4]), pt=st.sampled_from([0.0, 0.5, 1.0]),
       ps=st.sampled_from([0.0, 0.4]))
def test_blocked_multiply_equals_dense_product(d, seed, dens, pt, ps):
    n, k, m, blk = d
    rng = np.random.default_rng(seed)
    A = random_block_dataset(rng, n, k, blk, density=dens, p_transposed=pt, p_sparse=ps, sparse_density=0.3)
    B = random_block_dataset(rng, k, m, blk, density=dens, p_transposed=pt)
    C = O.matrix_multiply(A, n, k, B, k, m, blk)
    want = O.assemble(A, n, k, blk) @ O.assemble(B, k, m, blk)
    np.testing.assert_allclose(O.assemble(C, n, m, blk), want, rtol=0, atol=1e-12 * max(1, k))
    for (i, j), c in C.items():
        # a result block exists iff some k has both A(i,k) and B(k,j) (join on k); outer product: every (i, j) pair
        ks = [kk for kk in range(-(-k // blk)) if (i, kk) in A and (kk, j) in B]
        assert ks, (i, j)
        if isinstance(c, O.DenseMatrix):
            assert not c.isTransposed          # MLMatrix.scala:101: products are column-major
        assert (c.numRows, c.numCols) == (min(blk, n - i * blk), min(blk, m - j * blk))
    expect = {(i, j) for (i, kk) in A for (k2, j) in B if k2 == kk}
    assert set(C) == expect


@settings(**SET)
@given(d=dims, seed=st.integers(0, 2 ** 20), dens=st.sampled_from([1.0, 0.6]), pt=st.sampled_from([0.0, 0.5]),
       ps=st.sampled_from([0.0, 0.5]))
def test_elementwise_joins(d, seed, dens, pt, ps):
    n, m, _, blk = d
    rng = np.random.default_rng(seed)
    A = random_block_dataset(rng, n, m, blk, density=dens, p_transposed=pt, p_sparse=ps, sparse_density=0.4, lo=0.5, hi=1.5)
    B = random_block_dataset(rng, n, m, blk, density=dens, p_transposed=pt, p_sparse=ps, sparse_density=0.4, lo=0.5, hi=1.5)
    fa, fb = O.assemble(A, n, m, blk), O.assemble(B, n, m, blk)
    S = O.add_element(A, n, m, B, n, m, blk)
    assert set(S) == set(A) | set(B)                                         # outer join
    np.testing.assert_allclose(O.assemble(S, n, m, blk), fa + fb, rtol=0, atol=1e-14)
    for key in set(A) - set(B):
        assert S[key] is A[key]                                              # one-sided blocks pass through untouched
    Pm = O.multiply_element(A, n, m, B, n, m, blk, compat_bugs=False)
    assert set(Pm) == set(A) & set(B)                                        # inner join
    mask = np.zeros((n, m), dtype=bool)
    for (i, j) in Pm:
        mask[i * blk:(i + 1) * blk, j * blk:(j + 1) * blk] = True
    np.testing.assert_allclose(O.assemble(Pm, n, m, blk), np.where(mask, fa * fb, 0.0), rtol=0, atol=1e-14)


@settings(**SET)
@given(d=dims, seed=st.integers(0, 2 ** 20), pt=st.sampled_from([0.0, 0.5]), ps=st.sampled_from([0.0, 0.5]))
def test_transpose_scalar_aggregate_slicing(d, seed, pt, ps):
    n, m, _, blk = d
    rng = np.random.default_rng(seed)
    A = random_block_dataset(rng, n, m, blk, p_transposed=pt, p_sparse=ps, sparse_density=0.4)
    fa = O.assemble(A, n, m, blk)
    T = O.transpose(A)
    np.testing.assert_array_equal(O.assemble(T, m, n, blk), fa.T)
    for (i, j), b in A.items():
        assert np.array_equal(T[(j, i)].values, b.values) and T[(j, i)].isTransposed != b.isTransposed   # same arrays, flag flipped
    np.testing.assert_allclose(O.assemble(O.multiply_scalar(A, 2.5), n, m, blk), 2.5 * fa, rtol=1e-15)
    np.testing.assert_allclose(O.assemble(O.row_sum(A, n, m), n, 1, blk)[:, 0], fa.sum(axis=1), atol=1e-12)
    np.testing.assert_allclose(O.assemble(O.col_sum(A, n, m), 1, m, blk)[0], fa.sum(axis=0), atol=1e-12)
    np.testing.assert_allclose(O.total_sum(A, n, m)[(0, 0)].values[0], fa.sum(), atol=1e-12)
    i, j = int(rng.integers(0, n)), int(rng.integers(0, m))
    np.testing.assert_array_equal(O.assemble(O.project(A, n, m, blk, True, i), 1, m, blk)[0], fa[i])
    np.testing.assert_array_equal(O.assemble(O.project(A, n, m, blk, False, j), n, 1, blk)[:, 0], fa[:, j])
    assert O.selection(A, n, m, blk, i, j)[(0, 0)].values[0] == fa[i, j]


@settings(**SET)
@given(r=st.integers(1, 12), c=st.integers(1, 12), seed=st.integers(0, 2 ** 20), dens=st.sampled_from([0.0, 0.2, 0.7]))
def test_block_formats_and_serializer_round_trip(r, c, seed, dens):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (r, c)) * (rng.random((r, c)) < dens)
    col = O.DenseMatrix(r, c, np.ascontiguousarray(a.T).reshape(-1), False)
    row = O.DenseMatrix(r, c, np.ascontiguousarray(a).reshape(-1), True)
    csc = col.toSparse()
    csr = O.DenseMatrix(c, r, np.ascontiguousarray(a).reshape(-1)).toSparse().transpose()
    for m in (col, row, csc, csr):
        np.testing.assert_array_equal(m.to_numpy(), a)
        assert all(m.apply(i, j) == a[i, j] for i in range(r) for j in range(c))
        back = O.deserialize(O.serialize(m))                                   # MLMatrixSerializer.scala:26-69
        assert type(back) is type(m) and back.isTransposed == m.isTransposed
        np.testing.assert_array_equal(back.to_numpy(), a)
        assert len(O.serialize(m)) == 7
    assert csc.colPtrs[-1] == len(csc.values) == np.count_nonzero(a)
    assert len(csr.colPtrs) == r + 1 and len(csc.colPtrs) == c + 1            # MLMatrix.scala:535-541


@settings(**SET)
@given(nr=st.integers(1, 5000), nc=st.integers(1, 5000), blk=st.sampled_from([1, 7, 64, 256, 1024]), i=st.integers(0, 80),
       j=st.integers(0, 80), p=st.integers(1, 64))
def test_partitioner_ranges(nr, nc, blk, i, j, p):
    assert O.row_partition(i, j, p) == i % p and O.column_partition(i, j, p) == j % p
    R, C, r, c = O.gen_block_cyclic_partitioner(nr, nc, blk)
    assert R == -(-nr // blk) and C == -(-nc // blk) and r >= 1 and c >= 1
    part = O.BlockCyclicPartitioner(R, C, r, c)
    assert 0 <= part.getPartition(i, j) < part.numPartitions
    # periodic in the per-partition block counts (BlockCyclicPartitioner.scala:52-58)
    assert part.getPartition(i, j) == part.getPartition(i + part.num_row_part, j + part.num_col_part)


@settings(**SET)
@given(r=st.integers(1, 14), k=st.integers(1, 14), c=st.integers(1, 14), seed=st.integers(0, 2 ** 20),
       dens=st.sampled_from([0.02, 0.09]), aT=st.booleans(), bT=st.booleans())
def test_multiply_sparse_sparse_formats(r, k, c, seed, dens, aT, bT):
    """LocalMatrix.multiplySparseSparse (LocalMatrix.scala:143-323): the product of the stored entries, in the storage format
    its four loop nests end with."""
    rng = np.random.default_rng(seed)

    def sparse(rows, cols, csr):
        a = rng.uniform(0.5, 1.5, (rows, cols)) * (rng.random((rows, cols)) < dens)
        if csr:
            return a, O.DenseMatrix(cols, rows, np.ascontiguousarray(a).reshape(-1)).toSparse().transpose()
        return a, O.DenseMatrix(rows, cols, np.ascontiguousarray(a.T).reshape(-1)).toSparse()
    fa, A = sparse(r, k, aT)
    fb, B = sparse(k, c, bT)
    C = O.multiplySparseSparse(A, B)
    np.testing.assert_allclose(C.to_numpy(), fa @ fb, atol=1e-13)
    nnz = int(np.count_nonzero(fa @ fb))
    if aT and not bT:
        assert isinstance(C, O.SparseMatrix) and not C.isTransposed
    elif aT and bT:
        assert isinstance(C, O.SparseMatrix) == (r * c > 2 * nnz + r + 1)
        assert isinstance(C, O.DenseMatrix) or C.isTransposed
    elif not aT and not bT:
        assert isinstance(C, O.SparseMatrix) == (r * c > 2 * nnz + c + 1)
    else:
        assert isinstance(C, O.DenseMatrix) == (r * c <= 2 * nnz + c)
    if isinstance(C, O.SparseMatrix):
        assert int(C.colPtrs[-1]) == nnz == C.values.size                       # exact zeros are not stored
