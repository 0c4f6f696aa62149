"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and
the committed golden vectors.  Block ids / presence / shapes / flags bit-exact, values <= 1e-5
relative (north_star) -- and <= 1e-11 for the native fp64 DMMA kernel."""
import json
import os

import numpy as np
import pytest

import matrel_b200 as mb
from oracle import matrel_oracle as O
from tests.util import (REL_TOL, TIGHT_TOL, assert_same_dataset, from_dataset, random_block_dataset, rel_err,
                        to_dataset, to_product_matrix)

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def mk(b):
    if b["type"] == "dense":
        return O.DenseMatrix(b["numRows"], b["numCols"], b["values"])
    return O.SparseMatrix(b["numRows"], b["numCols"], b["colPtrs"], b["rowIndices"], b["values"])


# ---------------------------------------------------------------------------------- golden fixtures
def test_golden_basic_matrix_ops(session):
    g = load("basic_matrix_ops")
    e = g["expected"]
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    mat1 = to_dataset(session, {(r, c): blocks[n] for r, c, n in g["mat1"]})
    mat2 = to_dataset(session, {(r, c): blocks[n] for r, c, n in g["mat2"]})
    prod = from_dataset(mat1.matrixMultiply(4, 4, mat2, 4, 4, 2))
    assert sorted(f"{i},{j}" for i, j in prod) == sorted(e["multiply_blocks_colmajor"])
    for key, vals in e["multiply_blocks_colmajor"].items():
        m = prod[tuple(map(int, key.split(",")))]
        assert isinstance(m, mb.DenseMatrix) and not m.isTransposed and (m.numRows, m.numCols) == (2, 2)
        assert m.values.tolist() == vals           # small integers: exact
    full = np.zeros((4, 4))
    for (i, j), m in prod.items():
        full[2 * i:2 * i + 2, 2 * j:2 * j + 2] = m.to_numpy()
    assert float(np.trace(full)) == e["trace"] and float(full[0, 3]) == e["selection_0_3"]
    assert full[:, 3].tolist() == e["column_3"]
    # transpose: flag flip + index swap
    t = from_dataset(mat1.t())
    assert sorted(t) == [(0, 0), (1, 1)] and all(m.isTransposed for m in t.values())
    tf = np.zeros((4, 4))
    for (i, j), m in t.items():
        tf[2 * i:2 * i + 2, 2 * j:2 * j + 2] = m.to_numpy()
    assert tf.tolist() == e["transpose_mat1_full"]
    # addElement: outer join; multiplyElement: inner join
    s = from_dataset(mat1.addElement(4, 4, mat2, 4, 4, 2))
    assert sorted(f"{i},{j}" for i, j in s) == sorted(e["add_present"])
    sf = np.zeros((4, 4))
    for (i, j), m in s.items():
        sf[2 * i:2 * i + 2, 2 * j:2 * j + 2] = m.to_numpy()
    assert sf.tolist() == e["add_full"]
    p = from_dataset(mat1.multiplyElement(4, 4, mat2, 4, 4, 2))
    assert sorted(f"{i},{j}" for i, j in p) == sorted(e["mul_present"])
    pf = np.zeros((4, 4))
    for (i, j), m in p.items():
        pf[2 * i:2 * i + 2, 2 * j:2 * j + 2] = m.to_numpy()
    assert pf.tolist() == e["mul_full_on_common"]


def test_golden_remaining_demos(session):
    """runMatrixTranspose / runMatrixScalar / runMatrixProjection (example/BasicMatrixOps.scala:47-78, 183-210) on the device."""
    g = load("basic_matrix_ops")
    e = g["expected"]
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    t = from_dataset(to_dataset(session, {(r, c): blocks[n] for r, c, n in g["transpose_demo"]["blocks"]}).t())
    assert sorted(f"{i},{j}" for i, j in t) == sorted(g["transpose_demo"]["expected"])
    for key, want in g["transpose_demo"]["expected"].items():
        m = t[tuple(map(int, key.split(",")))]
        assert m.isTransposed and m.to_numpy().tolist() == want
        assert isinstance(m, mb.SparseMatrix) == (key == "2,0")
    p = from_dataset(to_dataset(session, {(r, c): blocks[n] for r, c, n in g["power_demo"]["blocks"]}).power(2.0))
    for key, want in g["power_demo"]["expected"].items():
        m = p[tuple(map(int, key.split(",")))]
        assert m.to_numpy().tolist() == want
        assert isinstance(m, mb.SparseMatrix) == (key == "1,3")
    mat1 = to_dataset(session, {(r, c): blocks[n] for r, c, n in g["mat1"]})
    mat2 = to_dataset(session, {(r, c): blocks[n] for r, c, n in g["mat2"]})
    row = from_dataset(mat1.project(4, 4, 2, True, 2))
    assert sorted(row) == [(0, 1)] and row[(0, 1)].to_numpy()[0].tolist() == e["project_row_2_mat1"][2:]
    col = from_dataset(mat2.project(4, 4, 2, False, 3))
    assert sorted(col) == [(0, 0), (1, 0)]
    assert col[(0, 0)].to_numpy()[:, 0].tolist() + col[(1, 0)].to_numpy()[:, 0].tolist() == e["project_col_3_mat2"]


def test_golden_test_sparse(session):
    g = load("test_sparse")
    e = g["expected"]
    s1 = O.SparseMatrix(**g["spmat1"])
    s2 = O.SparseMatrix(**g["spmat2"])
    d2 = s2.toDense()
    csr1 = O.DenseMatrix(3, 3, s1.toArray()).transpose().toSparse().transpose()

    def mul(a, b, n=3):
        da, db = to_dataset(session, {(0, 0): a}), to_dataset(session, {(0, 0): b})
        # two k-blocks declared (blkSize 2 over 3 columns) so the general (join) path runs
        return from_dataset(da.matrixMultiply(3, 3, db, 3, n, 2))[(0, 0)].to_numpy()
    assert mul(s1, d2).tolist() == e["S1_times_S2"]                 # CSC x dense
    assert mul(csr1, d2).tolist() == e["S1_times_S2"]               # CSR x dense
    assert mul(s1.toDense(), s2).tolist() == e["S1_times_S2"]       # dense x sparse (densified)
    assert mul(s1, s2.transpose().toDense()).tolist() == e["S1_times_S2t"]
    assert mul(s2, s2.transpose().toDense()).tolist() == e["S2_times_S2t"]
    assert mul(s1, O.DenseMatrix(3, 1, g["denV"]), 1)[:, 0].tolist() == e["S1_times_v"]   # SpMV row of the dispatch
    # the demo's own fixtures are > 0.1 dense, so sparse x sparse takes the densify branch (LocalMatrix.scala:903)
    assert mul(s1, s2).tolist() == e["S1_times_S2"]
    # dense (+) sparse element-wise through the join operators
    da, db = to_dataset(session, {(0, 0): s1.toDense()}), to_dataset(session, {(0, 0): s2})
    assert from_dataset(da.addElement(3, 3, db, 3, 3, 3))[(0, 0)].to_numpy().tolist() == e["S1_plus_S2"]
    assert from_dataset(da.multiplyElement(3, 3, db, 3, 3, 3))[(0, 0)].to_numpy().tolist() == e["S1_hadamard_S2"]


def test_java_random_on_device(session):
    g = load("java_random")
    for seed, want in g["nextDouble"].items():
        ds = session.rand(3, 3, 3, int(seed))
        assert ds.get_block(0, 0).values[0] == want
    got = from_dataset(session.rand(300, 200, 128, 42))
    want = O.rand_dense_dataset(300, 200, 128, 42)
    assert_same_dataset(got, want, exact_storage=True)       # bit-identical to the JVM stream


# ---------------------------------------------------------------------------------- multiply parity
@pytest.mark.parametrize("n,k,m,blk,pt", [
    (256, 256, 256, 128, 0.0),      # tiles exactly, column-major only
    (256, 256, 256, 128, 0.5),      # mixed isTransposed: all four T/N combinations
    (300, 200, 260, 128, 0.5),      # ragged edge blocks (even sizes -> bulk-copy path with zero fill)
    (131, 77, 93, 64, 0.5),         # odd leading dimensions -> guarded element-load path
    (1024, 1024, 1024, 256, 0.3),   # BASELINE config[0] shape
    (96, 512, 64, 32, 0.5),         # many k-blocks, small tiles
    (40, 40, 40, 7, 0.5),           # tiny odd blocks
])
def test_multiply_dense_vs_oracle(session, n, k, m, blk, pt):
    rng = np.random.default_rng(n * 7 + k * 3 + m + blk)
    A = random_block_dataset(rng, n, k, blk, p_transposed=pt)
    B = random_block_dataset(rng, k, m, blk, p_transposed=pt)
    want = O.matrix_multiply(A, n, k, B, k, m, blk)
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, k, to_dataset(session, B), k, m, blk))
    assert_same_dataset(got, want, tol=TIGHT_TOL)
    # independent dense check of the whole product
    full = np.zeros((n, m))
    for (i, j), blkm in got.items():
        full[i * blk:i * blk + blkm.numRows, j * blk:j * blk + blkm.numCols] = blkm.to_numpy()
    dense = O.assemble(A, n, k, blk) @ O.assemble(B, k, m, blk)
    assert rel_err(full, dense) <= REL_TOL and rel_err(full, dense) <= TIGHT_TOL


@pytest.mark.parametrize("variant", [0, 1])
def test_multiply_both_tile_variants(session, variant):
    session.set_option("gemm_variant", variant)
    try:
        rng = np.random.default_rng(variant)
        n, blk = 520, 200
        A = random_block_dataset(rng, n, n, blk, p_transposed=0.5)
        B = random_block_dataset(rng, n, n, blk, p_transposed=0.5)
        want = O.matrix_multiply(A, n, n, B, n, n, blk)
        got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
        assert_same_dataset(got, want, tol=TIGHT_TOL)
    finally:
        session.set_option("gemm_variant", -1)


def test_multiply_block_sparse_presence(session):
    """Output block (i,j) exists iff some k has both A(i,k) and B(k,j) (join semantics)."""
    rng = np.random.default_rng(5)
    n, blk = 6 * 48, 48
    for trial in range(4):
        A = random_block_dataset(rng, n, n, blk, density=0.4, p_transposed=0.3)
        B = random_block_dataset(rng, n, n, blk, density=0.4, p_transposed=0.3)
        want = O.matrix_multiply(A, n, n, B, n, n, blk)
        got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
        assert_same_dataset(got, want, tol=TIGHT_TOL)
    # empty inputs -> empty product
    empty = session.emptyDataset()
    assert from_dataset(empty.matrixMultiply(n, n, to_dataset(session, B), n, n, blk)) == {}


def test_has_block_follows_join_presence(session):
    rng = np.random.default_rng(4)
    n, blk = 4 * 32, 32
    A = random_block_dataset(rng, n, n, blk, density=0.5)
    B = random_block_dataset(rng, n, n, blk, density=0.5)
    C = to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk)
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    for i in range(4):
        for j in range(4):
            assert C.has_block(i, j) == ((i, j) in want)
    assert not C.has_block(7, 0)


def test_multiply_with_sparse_blocks(session):
    rng = np.random.default_rng(11)
    n, blk = 4 * 64, 64
    A = random_block_dataset(rng, n, n, blk, p_transposed=0.3, p_sparse=0.5, sparse_density=0.3)
    B = random_block_dataset(rng, n, n, blk, p_transposed=0.3)
    want = O.matrix_multiply(A, n, n, B, n, n, blk)      # sparse x dense -> gemmsdd, dense x dense -> dgemm
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
    assert_same_dataset(got, want, tol=TIGHT_TOL)
    # dense x sparse: the sparse side is densified (LocalMatrix.scala:892)
    want = O.matrix_multiply(B, n, n, A, n, n, blk)
    got = from_dataset(to_dataset(session, B).matrixMultiply(n, n, to_dataset(session, A), n, n, blk))
    assert_same_dataset(got, want, tol=TIGHT_TOL)


def _sparse_blocks(rng, n, m, blk, density, fmt, presence=1.0):
    """All blocks sparse: fmt 'csc' / 'csr' / 'mix'."""
    out = {}
    for i in range(-(-n // blk)):
        for j in range(-(-m // blk)):
            if rng.random() >= presence:
                continue
            r, c = min(blk, n - i * blk), min(blk, m - j * blk)
            a = rng.uniform(0.5, 1.5, (r, c)) * (rng.random((r, c)) < density)
            csr = fmt == "csr" or (fmt == "mix" and rng.random() < 0.5)
            if csr:
                out[(i, j)] = O.DenseMatrix(c, r, np.ascontiguousarray(a).reshape(-1)).toSparse().transpose()
            else:
                out[(i, j)] = O.DenseMatrix(r, c, np.ascontiguousarray(a.T).reshape(-1)).toSparse()
    return out


@pytest.mark.parametrize("fa,fb", [("csc", "csc"), ("csr", "csr"), ("csr", "csc"), ("csc", "csr")])
@pytest.mark.parametrize("density", [0.004, 0.05])
def test_sparse_times_sparse_single_block_pair_formats(session, fa, fb, density):
    """LocalMatrix.multiplySparseSparse (LocalMatrix.scala:143-323), one k-block (outer-product path, no reduce): the four
    loop nests end in four different storage rules -- CSC / CSR / always-CSC / dense-or-CSC -- reproduced exactly
    (type, isTransposed, index arrays); values to fp64 tolerance."""
    rng = np.random.default_rng(int(density * 1000) + len(fa + fb))
    blk, nb = 64, 3
    A = _sparse_blocks(rng, nb * blk, blk, blk, density, fa)          # nb x 1 blocks
    B = _sparse_blocks(rng, blk, nb * blk - 7, blk, density, fb)      # 1 x nb blocks, ragged last column block
    want = O.matrix_multiply(A, nb * blk, blk, B, blk, nb * blk - 7, blk)
    got = from_dataset(to_dataset(session, A).matrixMultiply(nb * blk, blk, to_dataset(session, B), blk, nb * blk - 7, blk))
    assert_same_dataset(got, want, tol=TIGHT_TOL)
    kinds = {(isinstance(w, O.SparseMatrix), w.isTransposed) for w in want.values()}
    if (fa, fb) == ("csr", "csc"):
        assert kinds == {(True, False)}                              # always CSC, whatever the density
    if (fa, fb) == ("csr", "csr") and density < 0.01:
        assert kinds == {(True, True)}                               # CSR result


@pytest.mark.parametrize("fmt", ["csc", "csr", "mix"])
def test_sparse_times_sparse_chains_follow_the_add_rules(session, fmt):
    """Several k-blocks: every partial is a multiplySparseSparse result and `reduceByKey(LocalMatrix.add)` re-decides the
    format at each sparse + sparse step (LocalMatrix.scala:74-139); replayed in ascending k like the oracle."""
    rng = np.random.default_rng({"csc": 1, "csr": 2, "mix": 3}[fmt])
    n, blk = 4 * 48 - 5, 48
    for density, presence in [(0.01, 1.0), (0.06, 1.0), (0.02, 0.6)]:
        A = _sparse_blocks(rng, n, n, blk, density, fmt, presence)
        B = _sparse_blocks(rng, n, n, blk, density, fmt, presence)
        want = O.matrix_multiply(A, n, n, B, n, n, blk)
        got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
        assert_same_dataset(got, want, tol=TIGHT_TOL)
        np.testing.assert_allclose(O.assemble(want, n, n, blk), O.assemble(A, n, n, blk) @ O.assemble(B, n, n, blk), atol=1e-12)


def test_sparse_times_sparse_next_to_dense_partials_is_dense(session):
    """A block sum that contains any dense partial product is dense (LocalMatrix.add), so low-density sparse x sparse pairs
    in such a sum only contribute values."""
    rng = np.random.default_rng(9)
    n, blk = 3 * 40, 40
    A = _sparse_blocks(rng, n, n, blk, 0.03, "mix")
    B = _sparse_blocks(rng, n, n, blk, 0.03, "mix")
    dense = rng.uniform(-1, 1, (blk, blk))
    A[(0, 1)] = O.DenseMatrix(blk, blk, np.ascontiguousarray(dense.T).reshape(-1))     # dense x sparse -> densify (:892)
    B[(2, 2)] = O.DenseMatrix(blk, blk, np.ascontiguousarray(dense).reshape(-1), True)  # sparse x dense -> gemmsdd
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
    assert_same_dataset(got, want, tol=TIGHT_TOL)
    assert all(isinstance(want[(0, j)], O.DenseMatrix) for j in range(3))
    assert all(isinstance(want[(i, 2)], O.DenseMatrix) for i in range(3))


def test_config5_shape_csr_times_dense(session):
    """BASELINE configs[4] at reduced N: 1%-sparse CSR blocks x dense, 1024-blocks (fused shared-memory SpMM kernel),
    against BLAS.gemmsdd restated in the oracle."""
    import scipy.sparse as sp
    rng = np.random.default_rng(55)
    n, blk = 2048, 1024
    A, B = {}, O.rand_dense_dataset(n, n, blk, 43, transposed_mask=lambda i, j: (i + j) % 2 == 1)
    for i in range(2):
        for k in range(2):
            m = sp.random(blk, blk, density=0.01, format="csr", random_state=rng, dtype=np.float64)
            m.sort_indices()
            A[(i, k)] = O.SparseMatrix(blk, blk, m.indptr, m.indices, m.data, True)
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    session.reset_stats()
    got = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
    assert session.stats()["kernel_launches"] <= 3          # CSR re-packing, B transposes, ONE fused SpMM launch over all blocks
    assert_same_dataset(got, want, tol=1e-14)


def test_multiply_outer_product_paths(session):
    """Inner dimension of one block: multiplyOuterProductDuplicate{Left,Right} (rank-k update);
    defect B1 (Left variant throws in the reference) is intentionally not reproduced."""
    rng = np.random.default_rng(3)
    for (n, m, k, blk) in [(5 * 32, 2 * 32, 32, 32), (2 * 32, 5 * 32, 20, 32)]:
        A = random_block_dataset(rng, n, k, blk, p_transposed=0.5)
        B = random_block_dataset(rng, k, m, blk, p_transposed=0.5)
        want = O.matrix_multiply(A, n, k, B, k, m, blk)
        got = from_dataset(to_dataset(session, A).matrixMultiply(n, k, to_dataset(session, B), k, m, blk))
        assert len(got) == len(A) * len(B)
        assert_same_dataset(got, want, tol=TIGHT_TOL)


def test_multiply_linearity_and_transpose_identity(session):
    """Size-independent properties: (A+A)B == 2(AB); (AB)^T == B^T A^T through flag-only transposes."""
    n, blk = 512, 128
    A, B = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    AB = A.matrixMultiply(n, n, B, n, n, blk)
    lhs = from_dataset(A.addElement(n, n, A, n, n, blk).matrixMultiply(n, n, B, n, n, blk))
    rhs = from_dataset(AB.multiplyScalar(2.0))
    for key in rhs:
        assert rel_err(lhs[key].values, rhs[key].values) <= 1e-14
    BtAt = from_dataset(B.t().matrixMultiply(n, n, A.t(), n, n, blk))
    ABt = from_dataset(AB.t().materialize())
    for key in ABt:
        assert not ABt[key].isTransposed
        assert rel_err(BtAt[key].values, ABt[key].values) <= 1e-13


# ---------------------------------------------------------------------------------- siblings
@pytest.mark.parametrize("op", ["add", "mul", "div"])
@pytest.mark.parametrize("pt,ps", [(0.0, 0.0), (0.5, 0.0), (0.4, 0.4)])
def test_elementwise_vs_oracle(session, op, pt, ps):
    rng = np.random.default_rng(hash((op, pt, ps)) % 1000)
    n, m, blk = 333, 270, 100
    A = random_block_dataset(rng, n, m, blk, density=0.8, p_transposed=pt, p_sparse=ps, sparse_density=0.3, lo=0.5, hi=2.0)
    B = random_block_dataset(rng, n, m, blk, density=0.8, p_transposed=pt, lo=0.5, hi=2.0)
    dA, dB = to_dataset(session, A), to_dataset(session, B)
    if op == "add":
        want, got = O.add_element(A, n, m, B, n, m, blk), dA.addElement(n, m, dB, n, m, blk)
    elif op == "mul":
        want, got = O.multiply_element(A, n, m, B, n, m, blk), dA.multiplyElement(n, m, dB, n, m, blk)
    else:
        want, got = O.divide_element(A, n, m, B, n, m, blk), dA.divideElement(n, m, dB, n, m, blk)
    assert_same_dataset(from_dataset(got), want, tol=1e-15)


@pytest.mark.parametrize("op", ["add", "mul", "div"])
def test_sparse_op_sparse_output_format_rule(session, op):
    """sparse (op) sparse: dense result turned into CSC iff rows*cols > 2*nnz + cols + 1 (LocalMatrix.scala:74-139,521-602)."""
    rng = np.random.default_rng({"add": 1, "mul": 2, "div": 3}[op])
    n, m, blk = 200, 150, 64
    for dens in (0.02, 0.3, 0.7):
        A = random_block_dataset(rng, n, m, blk, p_sparse=1.0, sparse_density=dens, lo=0.5, hi=2.0)
        B = random_block_dataset(rng, n, m, blk, p_sparse=1.0, sparse_density=dens, lo=0.5, hi=2.0)
        dA, dB = to_dataset(session, A), to_dataset(session, B)
        if op == "add":
            want, got = O.add_element(A, n, m, B, n, m, blk), dA.addElement(n, m, dB, n, m, blk)
        elif op == "mul":
            want, got = O.multiply_element(A, n, m, B, n, m, blk), dA.multiplyElement(n, m, dB, n, m, blk)
        else:
            want, got = O.divide_element(A, n, m, B, n, m, blk), dA.divideElement(n, m, dB, n, m, blk)
        got = from_dataset(got)
        kinds = {type(v).__name__ for v in want.values()}
        assert_same_dataset(got, want, tol=1e-15)       # types (sparse vs dense), colPtrs, rowIndices exact
        if op == "mul" and dens == 0.02:
            assert kinds == {"SparseMatrix"}
        if op == "div":
            assert kinds == {"DenseMatrix"}             # 0/0 = NaN counts as non-zero everywhere


def test_divide_compat_switch(session):
    """Defect B4: (Sparse, Dense) divide = dense/sparse under compat_bugs, sparse/dense otherwise."""
    s = O.SparseMatrix(2, 2, [0, 1, 2], [0, 1], [2.0, 4.0])
    d = O.DenseMatrix(2, 2, [8.0, 1.0, 1.0, 8.0])
    A, B = {(0, 0): s}, {(0, 0): d}
    got = from_dataset(to_dataset(session, A).divideElement(2, 2, to_dataset(session, B), 2, 2, 2))
    assert_same_dataset(got, O.divide_element(A, 2, 2, B, 2, 2, 2, compat_bugs=True), tol=0)
    with mb.MatfastSession(device=0, compat_bugs=False) as s2:
        got = from_dataset(to_dataset(s2, A).divideElement(2, 2, to_dataset(s2, B), 2, 2, 2))
        assert_same_dataset(got, O.divide_element(A, 2, 2, B, 2, 2, 2, compat_bugs=False), tol=0)


def test_scalar_ops_and_power_preserve_layout(session):
    rng = np.random.default_rng(9)
    A = random_block_dataset(rng, 257, 130, 64, density=0.9, p_transposed=0.5, p_sparse=0.3, sparse_density=0.2, lo=0.1, hi=3.0)
    dA = to_dataset(session, A)
    assert_same_dataset(from_dataset(dA.addScalar(1.5)), O.add_scalar(A, 1.5), exact_storage=True)
    assert_same_dataset(from_dataset(dA.multiplyScalar(-2.25)), O.multiply_scalar(A, -2.25), exact_storage=True)
    assert_same_dataset(from_dataset(dA.power(2.0)), O.power(A, 2.0), tol=1e-15)
    assert_same_dataset(from_dataset(dA.power(0.5)), O.power(A, 0.5), tol=1e-15)
    assert_same_dataset(from_dataset(dA.power(-1.3)), O.power(A, -1.3), tol=1e-14)


def test_transpose_is_metadata_only_and_materialize(session):
    rng = np.random.default_rng(2)
    A = random_block_dataset(rng, 200, 120, 64, density=0.8, p_transposed=0.4, p_sparse=0.2, sparse_density=0.2)
    dA = to_dataset(session, A)
    dT = dA.t()
    assert_same_dataset(from_dataset(dT), O.transpose(A), exact_storage=True)
    # shares device memory with its source (MLMatrix.scala:312 "sharing the same underlying data")
    key = next(k for k, m in A.items() if isinstance(m, O.DenseMatrix))
    assert dT.block_device_ptr(key[1], key[0]) == dA.block_device_ptr(*key)
    mat = from_dataset(dT.materialize())
    for (i, j), m in O.transpose(A).items():
        g = mat[(i, j)]
        assert isinstance(g, mb.DenseMatrix) and not g.isTransposed
        assert np.array_equal(g.to_numpy(), m.to_numpy())
    assert_same_dataset(from_dataset(dA.t().t()), A, exact_storage=True)


def test_rank_one_update(session):
    rng = np.random.default_rng(4)
    n, blk = 150, 64
    A = random_block_dataset(rng, n, n, blk, p_transposed=0.5)
    v = random_block_dataset(rng, n, 1, blk)
    # intended semantics A + v v^T (compat off)
    with mb.MatfastSession(device=0, compat_bugs=False) as s2:
        got = from_dataset(to_dataset(s2, A).matrixRankOneUpdate(n, n, to_dataset(s2, v), n, 1, blk))
        want = O.rank_one_update(A, n, n, v, n, 1, blk, compat_bugs=False)
        assert_same_dataset(got, want, tol=1e-15)
        full = np.zeros((n, n))
        for (i, j), m in got.items():
            full[i * blk:i * blk + m.numRows, j * blk:j * blk + m.numCols] = m.to_numpy()
        vv = O.assemble(v, n, 1, blk)
        assert rel_err(full, O.assemble(A, n, n, blk) + vv @ vv.T) <= 1e-15
    # compat: the reference's requires only admit 1-row matrices, and B3 drops A
    with pytest.raises(mb.IllegalArgumentException, match="requirement failed: Vector column size is not 1, but #cols = 150"):
        to_dataset(session, A).matrixRankOneUpdate(n, n, to_dataset(session, v), n, 1, blk)
    A1 = {(0, 0): O.DenseMatrix(1, 1, [5.0])}
    v1 = {(0, 0): O.DenseMatrix(1, 1, [3.0])}
    got = from_dataset(to_dataset(session, A1).matrixRankOneUpdate(1, 1, to_dataset(session, v1), 1, 1, 1))
    assert_same_dataset(got, O.rank_one_update(A1, 1, 1, v1, 1, 1, 1, compat_bugs=True), tol=0)
    assert got[(0, 0)].values.tolist() == [9.0]


# ---------------------------------------------------------------------------------- errors
def test_error_messages_match_reference(session):
    a = to_dataset(session, {(0, 0): O.DenseMatrix(2, 2, [1, 2, 3, 4])})
    with pytest.raises(mb.IllegalArgumentException) as e:
        a.matrixMultiply(4, 4, a, 6, 4, 2)
    assert str(e.value) == "requirement failed: Matrix dimension not match, leftColNum = 4, rightRowNum = 6"
    with pytest.raises(mb.IllegalArgumentException) as e:
        a.addElement(4, 4, a, 5, 4, 2)
    assert str(e.value) == "requirement failed: Row number not match, leftRowNum = 4, rightRowNum = 5"
    with pytest.raises(mb.IllegalArgumentException) as e:
        a.multiplyElement(4, 4, a, 4, 3, 2)
    assert str(e.value) == "requirement failed: Col number not match, leftColNum = 4, rightColNum = 3"
    with pytest.raises(mb.IllegalArgumentException, match="The number of values supplied doesn't match the size of the matrix! values.length: 3, numRows \\* numCols: 4"):
        session.createDataset([mb.MatrixBlock(0, 0, mb.DenseMatrix(2, 2, [1, 2, 3]))])
    with pytest.raises(mb.IllegalArgumentException, match="Expecting 3 colPtrs when numCols = 2 but got 2"):
        session.createDataset([mb.MatrixBlock(0, 0, mb.SparseMatrix(2, 2, [0, 1], [0], [1.0]))])
    with pytest.raises(mb.IllegalArgumentException, match="The columns of A don't match the rows of B. A: 2, B: 3"):
        b = to_dataset(session, {(0, 0): O.DenseMatrix(3, 2, np.arange(6.0))})
        a.matrixMultiply(2, 4, b, 4, 2, 2)
    with pytest.raises(KeyError):
        a.get_block(3, 3)


# ---------------------------------------------------------------------------------- BASELINE sizes
def test_config1_1024_blk256(session):
    """BASELINE configs[0]: 1024^2 fp64, 256-blocks, device-generated U(0,1) inputs vs the oracle."""
    n, blk = 1024, 256
    A, B = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    got = from_dataset(A.matrixMultiply(n, n, B, n, n, blk))
    want = O.matrix_multiply(O.rand_dense_dataset(n, n, blk, 42), n, n, O.rand_dense_dataset(n, n, blk, 43), n, n, blk)
    assert_same_dataset(got, want, tol=TIGHT_TOL)


def test_config2_4096_blk512(session):
    """BASELINE configs[1]: 4096^2 fp64, 512-blocks."""
    n, blk = 4096, 512
    A, B = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    C = A.matrixMultiply(n, n, B, n, n, blk)
    Af = O.assemble(O.rand_dense_dataset(n, n, blk, 42), n, n, blk)
    Bf = O.assemble(O.rand_dense_dataset(n, n, blk, 43), n, n, blk)
    want = Af @ Bf
    got = np.zeros((n, n))
    for b in C.collect():
        assert isinstance(b.matrix, mb.DenseMatrix) and not b.matrix.isTransposed
        got[b.rid * blk:(b.rid + 1) * blk, b.cid * blk:(b.cid + 1) * blk] = b.matrix.to_numpy()
    assert len(C.block_ids()) == 64
    assert rel_err(got, want) <= TIGHT_TOL


def test_full_size_16384_properties(session):
    """BASELINE metric size (16384^2, 1024-blocks): checked through size-independent properties --
    a row/column checksum identity (e^T (A B) = (e^T A) B, (A B) e = A (B e)) and sampled blocks
    against float64 numpy on the host."""
    n, blk = 16384, 1024
    nb = n // blk
    A, B = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    C = A.matrixMultiply(n, n, B, n, n, blk)
    assert len(C.block_ids()) == nb * nb
    hostA = {k: A.get_block(*k) for k in A.block_ids() if k[0] in (0, nb - 1)}
    hostB = {k: B.get_block(*k) for k in B.block_ids() if k[1] in (0, nb - 1)}
    for (i, j) in [(0, 0), (nb - 1, nb - 1), (0, nb - 1)]:
        want = sum(hostA[(i, k)].to_numpy() @ hostB[(k, j)].to_numpy() for k in range(nb))
        got = C.get_block(i, j)
        assert not got.isTransposed and (got.numRows, got.numCols) == (blk, blk)
        assert rel_err(got.to_numpy(), want) <= 1e-12
    # checksum of checksums over one block row of C: sum_j C(0,j) e  ==  sum_k A(0,k) (sum_j B(k,j) e)
    rowsum_C = sum(C.get_block(0, j).to_numpy().sum(axis=1) for j in range(nb))
    Be = [sum(B.get_block(k, j).to_numpy().sum(axis=1) for j in range(nb)) for k in range(nb)]
    want = sum(hostA[(0, k)].to_numpy() @ Be[k] for k in range(nb))
    assert rel_err(rowsum_C, want) <= 1e-12


# ---------------------------------------------------------------------------------- aggregates (section 8f-2)
def test_aggregates_golden_and_oracle(session):
    g = load("basic_matrix_ops")
    e = g["expected"]
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    A = {(r, c): blocks[n] for r, c, n in g["mat1"]}
    B = {(r, c): blocks[n] for r, c, n in g["mat2"]}
    mat1, mat2 = to_dataset(session, A), to_dataset(session, B)
    rs = from_dataset(mat1.t().rowSum(4, 4))                       # BasicMatrixOps.scala:159
    assert sorted(rs) == [(0, 0), (1, 0)]
    assert [x for k in sorted(rs) for x in rs[k].values.tolist()] == e["rowSum_mat1_t"]
    assert all((m.numRows, m.numCols) == (2, 1) for m in rs.values())
    cs = from_dataset(mat2.colSum(4, 4))                           # :166
    assert sorted(cs) == [(0, 0), (0, 1)]
    assert [x for k in sorted(cs) for x in cs[k].values.tolist()] == e["colSum_mat2"]
    tr = from_dataset(mat1.matrixMultiply(4, 4, mat2, 4, 4, 2).trace(4, 4))   # :174
    assert tr[(0, 0)].values.tolist() == [e["trace"]]
    with pytest.raises(mb.IllegalArgumentException, match="Cannot perform trace\\(\\) on a rectangle matrix"):
        mat1.trace(4, 5)
    # random dense / transposed / sparse blocks vs the oracle's intended reductions
    rng = np.random.default_rng(8)
    n, m, blk = 300, 260, 64
    D = random_block_dataset(rng, n, m, blk, density=0.8, p_transposed=0.5, p_sparse=0.3, sparse_density=0.2)
    dD = to_dataset(session, D)
    assert_same_dataset(from_dataset(dD.rowSum(n, m)), O.row_sum(D, n, m), tol=1e-14)
    assert_same_dataset(from_dataset(dD.colSum(n, m)), O.col_sum(D, n, m), tol=1e-14)
    assert_same_dataset(from_dataset(dD.sum(n, m)), O.total_sum(D, n, m), tol=1e-13)
    S = random_block_dataset(rng, 256, 256, 64, density=0.7, p_transposed=0.5, p_sparse=0.3, sparse_density=0.2)
    assert_same_dataset(from_dataset(to_dataset(session, S).trace(256, 256)), O.trace(S, 256, 256), tol=1e-14)
    # where the reference's literal index arithmetic is sound (square dense blocks) it agrees with the intended sums
    Q = random_block_dataset(rng, 192, 192, 64, p_transposed=0.5)
    for key, v in O.row_sum(Q, 192, 192, literal=True).items():
        assert np.allclose(v.values, O.row_sum(Q, 192, 192)[key].values, rtol=1e-13)
    for key, v in O.col_sum(Q, 192, 192, literal=True).items():
        assert np.allclose(v.values, O.col_sum(Q, 192, 192)[key].values, rtol=1e-13)


def test_planner_identities_on_device(session):
    """The rewrites of M/execution/MatfastPlanner.scala hold on the device results:
    trace(A B) = sum(A^T o B) (:238-241) and sum(A B) = colSum(A) . rowSum(B) (:220-225)."""
    n, blk = 512, 128
    A, B = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    AB = A.matrixMultiply(n, n, B, n, n, blk)
    tr = AB.trace(n, n).get_block(0, 0).values[0]
    alt = A.t().multiplyElement(n, n, B, n, n, blk).sum(n, n).get_block(0, 0).values[0]
    assert abs(tr - alt) / abs(tr) < 1e-13
    total = AB.sum(n, n).get_block(0, 0).values[0]
    cs, rs = A.colSum(n, n), B.rowSum(n, n)
    dot = cs.matrixMultiply(1, n, rs, n, 1, blk).get_block(0, 0).values[0]
    assert abs(total - dot) / abs(total) < 1e-13


def test_lazy_plan_rewrites_match_direct_execution(session):
    """the planner's aggregate push-downs (MatfastPlanner.scala:168-244) give the same answers as the
    un-rewritten plan, and trace(A B) / sum(A B) / rowSum(A B) no longer launch the O(N^3) block GEMM."""
    from matrel_b200.plan import LazyDataset, Planner
    n, blk = 512, 128
    A, B = LazyDataset.of(session.rand(n, n, blk, 42)), LazyDataset.of(session.rand(n, n, blk, 43))
    queries = {
        "trace(AB)": A.matrixMultiply(n, n, B, n, n, blk).trace(n, n),
        "sum(AB)": A.matrixMultiply(n, n, B, n, n, blk).sum(n, n),
        "rowSum(AB)": A.matrixMultiply(n, n, B, n, n, blk).rowSum(n, n),
        "colSum(AB)": A.matrixMultiply(n, n, B, n, n, blk).colSum(n, n),
        "sum(2A+3)": A.multiplyScalar(2.0).addScalar(3.0).sum(n, n),
        "rowSum(A^T)": A.t().rowSum(n, n),
        "trace(A+B)": A.addElement(n, n, B, n, n, blk).trace(n, n),
        "colSum(A+1.5)": A.addScalar(1.5).colSum(n, n),
    }
    for name, q in queries.items():
        p1, p0 = Planner(True), Planner(False)
        session.reset_stats()
        got = from_dataset(q.execute(planner=p1))
        full_gemm_tiles = session.stats()["last_gemm_flops"] if session.stats()["gemm_launches"] else 0
        want = from_dataset(q.execute(planner=p0))
        assert sorted(got) == sorted(want), name
        for key in want:
            assert (got[key].numRows, got[key].numCols) == (want[key].numRows, want[key].numCols), (name, key)
            assert rel_err(got[key].to_numpy(), want[key].to_numpy()) <= 1e-12, name
        if name in ("trace(AB)",):
            assert "MatrixMatrixMultiplicationExecution" not in p1.trace and "SumDirectExecution" in p1.trace
        if name in ("sum(AB)", "rowSum(AB)", "colSum(AB)"):
            assert full_gemm_tiles < 2.0 * n ** 3 / 16          # a matrix-vector sized product, not N^3
        assert "MatrixMatrixMultiplicationExecution" in p0.trace or "AB" not in name


def test_pipelined_ingest_multiply_egress(session):
    """Section 8f-4: blocks put from pinned host memory are copied on the ingest stream; a multiply issued while
    copies are in flight is chunked per block row and must give exactly the resident result; get_block on the
    egress stream returns finished rows."""
    import torch
    n, blk = 2048, 256
    A = O.rand_dense_dataset(n, n, blk, 42, transposed_mask=lambda i, j: (i * 3 + j) % 4 == 0)
    B = O.rand_dense_dataset(n, n, blk, 43)
    ref = from_dataset(to_dataset(session, A).matrixMultiply(n, n, to_dataset(session, B), n, n, blk))
    pin = lambda m: mb.DenseMatrix(m.numRows, m.numCols, torch.from_numpy(m.values).pin_memory().numpy(), m.isTransposed)  # noqa: E731
    pA = [mb.MatrixBlock(i, j, pin(m)) for (i, j), m in A.items()]
    pB = [mb.MatrixBlock(i, j, pin(m)) for (i, j), m in B.items()]
    for rep in range(3):
        session.reset_stats()
        dB = session.createDataset(pB)
        dA = session.createDataset(pA)
        dC = dA.matrixMultiply(n, n, dB, n, n, blk)
        got = from_dataset(dC)
        assert_same_dataset(got, {k: O.DenseMatrix(v.numRows, v.numCols, v.values) for k, v in ref.items()}, exact_storage=True)
    with mb.MatfastSession(device=0) as s2:
        s2.set_option("pipeline", 0)                      # same answer with the overlap machinery off
        got = from_dataset(to_dataset(s2, A).matrixMultiply(n, n, to_dataset(s2, B), n, n, blk))
        assert_same_dataset(got, {k: O.DenseMatrix(v.numRows, v.numCols, v.values) for k, v in ref.items()}, exact_storage=True)


def test_project_selection_golden_and_pushdown(session):
    """Section 8f-4 slicing ops + their planner push-downs: BasicMatrixOps.scala:212 (column 3 of mat1*mat2) and
    :234 (element (0,3)) from the golden file; select(A B) runs as a 1 x k by k x 1 product."""
    from matrel_b200.plan import LazyDataset, Planner
    g = load("basic_matrix_ops")
    e = g["expected"]
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    A = {(r, c): blocks[n] for r, c, n in g["mat1"]}
    B = {(r, c): blocks[n] for r, c, n in g["mat2"]}
    mat1, mat2 = to_dataset(session, A), to_dataset(session, B)
    prod = mat1.matrixMultiply(4, 4, mat2, 4, 4, 2)
    col3 = from_dataset(prod.project(4, 4, 2, False, 3))
    assert sorted(col3) == [(0, 0), (1, 0)]
    assert [x for k in sorted(col3) for x in col3[k].values.tolist()] == e["column_3"]
    assert from_dataset(prod.selection(4, 4, 2, 0, 3))[(0, 0)].values.tolist() == [e["selection_0_3"]]
    for rewrite in (True, False):
        q = LazyDataset.of(mat1).matrixMultiply(4, 4, LazyDataset.of(mat2), 4, 4, 2)
        assert from_dataset(q.selection(4, 4, 2, 0, 3).execute(rewrite))[(0, 0)].values.tolist() == [e["selection_0_3"]]
        c3 = from_dataset(q.project(4, 4, 2, False, 3).execute(rewrite))
        assert [x for k in sorted(c3) for x in c3[k].values.tolist()] == e["column_3"]
    with pytest.raises(mb.IllegalArgumentException, match="row index should be smaller than #rows, index=9, #rows=4"):
        mat1.project(4, 4, 2, True, 9)
    # random blocks vs the oracle, and the O(N) selection push-down on a large product
    rng = np.random.default_rng(31)
    D = random_block_dataset(rng, 300, 200, 64, density=0.8, p_transposed=0.5, p_sparse=0.2, sparse_density=0.3)
    dD = to_dataset(session, D)
    assert_same_dataset(from_dataset(dD.project(300, 200, 64, True, 137)), O.project(D, 300, 200, 64, True, 137), exact_storage=True)
    assert_same_dataset(from_dataset(dD.project(300, 200, 64, False, 99)), O.project(D, 300, 200, 64, False, 99), exact_storage=True)
    assert_same_dataset(from_dataset(dD.selection(300, 200, 64, 250, 131)), O.selection(D, 300, 200, 64, 250, 131), exact_storage=True)
    n, blk = 1024, 256
    X, Y = session.rand(n, n, blk, 1), session.rand(n, n, blk, 2)
    q = LazyDataset.of(X).matrixMultiply(n, n, LazyDataset.of(Y), n, n, blk).selection(n, n, blk, 700, 300)
    p1 = Planner(True)
    session.reset_stats()
    fast = from_dataset(q.execute(planner=p1))[(0, 0)].values[0]
    assert session.stats()["last_gemm_flops"] <= 2 * n                      # one dot product, not 2 N^3
    slow = from_dataset(q.execute(False))[(0, 0)].values[0]
    assert abs(fast - slow) / abs(slow) < 1e-13


def test_vec_reshape(session):
    rng = np.random.default_rng(77)
    n, m, blk = 100, 70, 32
    D = random_block_dataset(rng, n, m, blk, density=0.8, p_transposed=0.5, p_sparse=0.2, sparse_density=0.3)
    got = from_dataset(to_dataset(session, D).vec(n, m, blk))
    assert_same_dataset(got, O.vec(D, n, m, blk), exact_storage=True)
    # stacking the vector blocks in key order reproduces numpy's column-major flattening block column by block column
    full = O.assemble(D, n, m, blk)
    R = -(-n // blk)
    for (key, _), v in got.items():
        col, i = divmod(key, R)
        assert np.array_equal(v.values, full[i * blk:i * blk + v.numRows, col])
