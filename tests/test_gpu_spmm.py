"""GPU tests of the pipelined CSR x dense kernel (matrel_b200/csrc/spmm.cu) against BLAS.gemmsdd restated in the oracle
(M/matrix/BLAS.scala:352-458), and against the simple shared-memory kernel it replaces for large blocks."""
import numpy as np
import pytest

import matrel_b200 as mb
from oracle import matrel_oracle as O
from tests.util import assert_same_dataset, from_dataset, to_dataset

pytestmark = pytest.mark.gpu


def csr_block(rng, rows, cols, density, sort=True):
    """A rows x cols CSR block (isTransposed = true) with ~density non-zeros."""
    mask = rng.random((rows, cols)) < density
    ptr = np.zeros(rows + 1, dtype=np.int32)
    idx, val = [], []
    for r in range(rows):
        c = np.flatnonzero(mask[r]).astype(np.int32)
        if not sort:
            rng.shuffle(c)
        idx.append(c)
        val.append(rng.uniform(-1, 1, c.size))
        ptr[r + 1] = ptr[r] + c.size
    return O.SparseMatrix(rows, cols, ptr, np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32),
                          np.concatenate(val) if val else np.zeros(0), True)


def dense_blocks(rng, n, m, blk, p_rowmajor):
    out = {}
    for i in range(-(-n // blk)):
        for j in range(-(-m // blk)):
            r, c = min(blk, n - i * blk), min(blk, m - j * blk)
            a = rng.uniform(-1, 1, (r, c))
            if rng.random() < p_rowmajor:
                out[(i, j)] = O.DenseMatrix(r, c, np.ascontiguousarray(a).reshape(-1), True)
            else:
                out[(i, j)] = O.DenseMatrix(r, c, np.ascontiguousarray(a.T).reshape(-1), False)
    return out


@pytest.mark.parametrize("n,k,m,blk,density,sort,p_rm", [
    (1024, 1024, 1024, 512, 0.01, True, 0.0),      # 2 x 2 x 2 blocks, column-major B (transposed once per multiply)
    (1024, 1024, 1024, 512, 0.01, True, 1.0),      # row-major B: tensor maps straight on the blocks
    (700, 900, 1100, 512, 0.02, True, 0.5),        # ragged rows (188), ragged k (388), ragged columns (76), mixed B layouts
    (1024, 512, 256, 1024, 0.05, True, 0.0),       # 1024-row blocks (two strips); segments above the shared-memory capacity
    (512, 1024, 96, 512, 0.01, False, 0.0),        # unsorted column indices inside the rows
    (2048, 2048, 64, 1024, 0.003, True, 0.5),      # very sparse: many empty rows
])
def test_spmm2_vs_oracle(session, n, k, m, blk, density, sort, p_rm):
    rng = np.random.default_rng(n + 3 * k + 5 * m + int(1000 * density))
    A = {(i, j): csr_block(rng, min(blk, n - i * blk), min(blk, k - j * blk), density, sort)
         for i in range(-(-n // blk)) for j in range(-(-k // blk))}
    B = dense_blocks(rng, k, m, blk, p_rm)
    want = O.matrix_multiply(A, n, k, B, k, m, blk)
    session.set_option("spmm_algo", 0)
    session.reset_stats()
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, k, to_dataset(session, B), k, m, blk))
    launches = session.stats()["kernel_launches"]
    assert_same_dataset(got, want, tol=1e-13)
    assert launches <= 3                                   # preparation, (transposes), ONE multiply launch
    session.set_option("spmm_algo", 1)                     # the simple kernel: same sums in the same order when the rows are sorted
    try:
        old = from_dataset(to_dataset(session, A).matrixMultiply(n, k, to_dataset(session, B), k, m, blk))
    finally:
        session.set_option("spmm_algo", 0)
    assert_same_dataset(old, want, tol=1e-13)
    if sort:
        for key in got:
            assert np.array_equal(got[key].values, old[key].values), key


def test_spmm2_accumulates_onto_dense_partials(session):
    """A block row that mixes dense and CSR blocks: the GEMM writes the dense partial sums, the sparse kernel adds on top."""
    rng = np.random.default_rng(77)
    n, blk = 1024, 512
    A = {(0, 0): csr_block(rng, blk, blk, 0.02), (1, 1): csr_block(rng, blk, blk, 0.02)}
    for key in [(0, 1), (1, 0)]:
        A[key] = O.DenseMatrix(blk, blk, rng.uniform(-1, 1, blk * blk))
    B = dense_blocks(rng, n, n, blk, 0.5)
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    with mb.MatfastSession(device=0, gemm_algo=1) as s:
        got = from_dataset(to_dataset(s, A).matrixMultiply(n, n, to_dataset(s, B), n, n, blk))
    assert_same_dataset(got, want, tol=1e-13)


def test_config5_blocks_from_sprand(session):
    """BASELINE configs[4] inputs at reduced N: A blocks = SparseMatrix.sprand(1024, 1024, 0.01, new java.util.Random(seed))
    (M/matrix/MLMatrix.scala:791-856, restated by the oracle) presented as CSR (its transpose's arrays), B dense U(0,1)."""
    n, blk = 2048, 1024
    nb = n // blk
    A = {}
    for i in range(nb):
        for k in range(nb):
            csc = O.sprand(blk, blk, 0.01, O.JavaRandom(1000 + i * nb + k))
            assert csc.values.size == int(np.ceil(blk * blk * 0.01))
            A[(i, k)] = csc.transpose()                    # CSR of the transposed pattern: same arrays, isTransposed = true
            assert A[(i, k)].isTransposed
    B = O.rand_dense_dataset(n, n, blk, 43)
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
    assert_same_dataset(got, want, tol=1e-13)


@pytest.mark.parametrize("nrows,ncols,blk,density,csr", [(2048, 2048, 1024, 0.01, True), (512, 1024, 256, 0.05, False),
                                                           (1024, 512, 512, 0.002, True)])
def test_device_sprand_is_bit_identical_to_the_jvm_stream(session, nrows, ncols, blk, density, csr):
    """mr_matrix_sprand restates SparseMatrix.sprand (M/matrix/MLMatrix.scala:791-856) in parallel on the device: the same
    coordinates (first nnz distinct draws of the java.util.Random stream, column-major) and the same U(0,1) values, bit for bit,
    as the sequential oracle."""
    from matrel_b200.dataset import sprand
    ds = from_dataset(sprand(session, nrows, ncols, blk, density, 77, csr=csr))
    nbc = -(-ncols // blk)
    assert sorted(ds) == [(i, j) for i in range(-(-nrows // blk)) for j in range(nbc)]
    for (i, j), got in ds.items():
        r, c = min(blk, nrows - i * blk), min(blk, ncols - j * blk)
        want = O.sprand(c, r, density, O.JavaRandom(77 + i * nbc + j)).transpose() if csr else O.sprand(r, c, density, O.JavaRandom(77 + i * nbc + j))
        assert isinstance(got, mb.SparseMatrix) and (got.numRows, got.numCols, got.isTransposed) == (r, c, csr)
        assert got.colPtrs.tolist() == want.colPtrs.tolist()
        assert got.rowIndices.tolist() == want.rowIndices.tolist()
        assert np.array_equal(got.values, want.values)
