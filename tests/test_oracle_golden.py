"""CPU tests: the numpy oracle against the committed golden vectors (tests/golden/*.json, derived
independently by tests/golden/make_golden.py from the reference's own demo fixtures)."""
import json
import os

import numpy as np
import pytest

from oracle import matrel_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def mk(b):
    if b["type"] == "dense":
        return O.DenseMatrix(b["numRows"], b["numCols"], b["values"])
    return O.SparseMatrix(b["numRows"], b["numCols"], b["colPtrs"], b["rowIndices"], b["values"])


@pytest.fixture(scope="module")
def basic():
    g = load("basic_matrix_ops")
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    mat1 = {(r, c): blocks[n] for r, c, n in g["mat1"]}
    mat2 = {(r, c): blocks[n] for r, c, n in g["mat2"]}
    return g, mat1, mat2


def test_multiply_golden(basic):
    g, mat1, mat2 = basic
    e = g["expected"]
    out = O.matrix_multiply(mat1, 4, 4, mat2, 4, 4, 2)
    assert sorted(f"{i},{j}" for i, j in out) == sorted(e["multiply_blocks_colmajor"])
    for key, vals in e["multiply_blocks_colmajor"].items():
        i, j = map(int, key.split(","))
        m = out[(i, j)]
        assert isinstance(m, O.DenseMatrix) and not m.isTransposed
        assert m.values.tolist() == vals
    for key in e["multiply_absent"]:
        assert tuple(map(int, key.split(","))) not in out
    full = O.assemble(out, 4, 4, 2)
    assert full.tolist() == e["multiply_full"]
    assert float(np.trace(full)) == e["trace"]
    assert float(full[0, 3]) == e["selection_0_3"]
    assert full[:, 3].tolist() == e["column_3"]


def test_transpose_add_mul_golden(basic):
    g, mat1, mat2 = basic
    e = g["expected"]
    t = O.transpose(mat1)
    assert O.assemble(t, 4, 4, 2).tolist() == e["transpose_mat1_full"]
    assert O.assemble(t, 4, 4, 2).sum(axis=1).tolist() == e["rowSum_mat1_t"]
    assert O.assemble(mat2, 4, 4, 2).sum(axis=0).tolist() == e["colSum_mat2"]
    for m in t.values():
        assert m.isTransposed            # flag flip only (MLMatrix.scala:312)
    s = O.add_element(mat1, 4, 4, mat2, 4, 4, 2)
    assert sorted(f"{i},{j}" for i, j in s) == sorted(e["add_present"])
    assert O.assemble(s, 4, 4, 2).tolist() == e["add_full"]
    assert isinstance(s[(0, 1)], O.DenseMatrix)      # one-sided block passes through untouched
    p = O.multiply_element(mat1, 4, 4, mat2, 4, 4, 2)
    assert sorted(f"{i},{j}" for i, j in p) == sorted(e["mul_present"])
    assert O.assemble(p, 4, 4, 2).tolist() == e["mul_full_on_common"]


def test_remaining_demos_golden(basic):
    """runMatrixTranspose / runMatrixScalar / runMatrixProjection of example/BasicMatrixOps.scala:47-78, 183-210."""
    g, mat1, mat2 = basic
    e = g["expected"]
    blocks = {k: mk(v) for k, v in g["blocks"].items()}
    demo = {(r, c): blocks[n] for r, c, n in g["transpose_demo"]["blocks"]}
    t = O.transpose(demo)
    assert sorted(f"{i},{j}" for i, j in t) == sorted(g["transpose_demo"]["expected"])
    for key, want in g["transpose_demo"]["expected"].items():
        m = t[tuple(map(int, key.split(",")))]
        assert m.isTransposed and m.to_numpy().tolist() == want
        assert isinstance(m, O.SparseMatrix) == (key == "2,0")           # the sparse block stays sparse (CSR now)
    demo = {(r, c): blocks[n] for r, c, n in g["power_demo"]["blocks"]}
    p = O.power(demo, 2.0)
    for key, want in g["power_demo"]["expected"].items():
        m = p[tuple(map(int, key.split(",")))]
        assert m.to_numpy().tolist() == want
        assert isinstance(m, O.SparseMatrix) == (key == "1,3")           # map over the stored values only
    row = O.project(mat1, 4, 4, 2, True, 2)
    assert sorted(row) == [(0, 1)] and O.assemble(row, 1, 4, 2)[0].tolist() == e["project_row_2_mat1"]
    col = O.project(mat2, 4, 4, 2, False, 3)
    assert sorted(col) == [(0, 0), (1, 0)] and O.assemble(col, 4, 1, 2)[:, 0].tolist() == e["project_col_3_mat2"]


def test_test_sparse_golden():
    g = load("test_sparse")
    e = g["expected"]
    s1 = O.SparseMatrix(**g["spmat1"])
    s2 = O.SparseMatrix(**g["spmat2"])
    assert s1.to_numpy().tolist() == g["S1_full"]
    assert s2.to_numpy().tolist() == g["S2_full"]
    # sparse x (dense view of the other operand): the dispatch rows the hot path keeps
    assert O.matrixMultiplication(s1, s2.toDense()).to_numpy().tolist() == e["S1_times_S2"]
    assert O.matrixMultiplication(s1.toDense(), s2).to_numpy().tolist() == e["S1_times_S2"]
    assert O.gemmsdd_loops(s1, s2.toDense()).to_numpy().tolist() == e["S1_times_S2"]
    assert O.gemmsdd_loops(s1, s2.transpose().toDense()).to_numpy().tolist() == e["S1_times_S2t"]
    assert O.gemmsdd_loops(s2, s2.transpose().toDense()).to_numpy().tolist() == e["S2_times_S2t"]
    # CSR variant of the same matrix (transpose of the transpose stored explicitly)
    csr = O.DenseMatrix(3, 3, s1.toArray()).transpose().toSparse().transpose()
    assert csr.isTransposed and csr.to_numpy().tolist() == g["S1_full"]
    assert O.gemmsdd_loops(csr, s2.toDense()).to_numpy().tolist() == e["S1_times_S2"]
    assert O.gemmsdd(csr, s2.toDense()).to_numpy().tolist() == e["S1_times_S2"]
    v = O.DenseMatrix(3, 1, g["denV"])
    assert O.matrixMultiplication(s1, v).to_numpy()[:, 0].tolist() == e["S1_times_v"]
    assert O.elementWiseMultiply(s1, s2).to_numpy().tolist() == e["S1_hadamard_S2"]
    assert O.add(s1, s2).to_numpy().tolist() == e["S1_plus_S2"]
    # multiplySparseSparse called directly by the demo (LocalMatrix.scala:1112, :1132): CSC x CSC and CSC x CSR
    assert O.multiplySparseSparse(s1, s2).to_numpy().tolist() == e["S1_times_S2"]
    assert O.multiplySparseSparse(s1, s2.transpose()).to_numpy().tolist() == e["S1_times_S2t"]
    assert isinstance(O.multiplySparseSparse(s1, s2), O.DenseMatrix)        # 9 > 2*7 + 4 is false -> densified
    # sparse+sparse density rule (LocalMatrix.scala:133): 9 > 2*9 + 4 is false -> dense result
    assert isinstance(O.add(s1, s2), O.DenseMatrix)
    assert isinstance(O.elementWiseMultiply(s1, s2), O.SparseMatrix)   # 1 nnz: 9 > 2 + 4


def test_partitioners_golden():
    g = load("partitioners")
    for c in g["cases"]:
        p = O.gen_block_cyclic_partitioner(c["nrows"], c["ncols"], c["blkSize"])
        assert list(p) == c["params"]
        bc = O.BlockCyclicPartitioner(*p)
        assert bc.numPartitions == c["numPartitions"]
        used = set()
        for i in range(p[0]):
            for j in range(p[1]):
                v = bc.getPartition(i, j)
                used.add(v)
                if c["table"] is not None:
                    assert v == c["table"][i][j]
        assert sorted(used) == c["used"]
    k = g["known"]
    assert list(O.gen_block_cyclic_partitioner(16384, 16384, 1024)) == k["16384/1024"]["params"]
    for i, p, want in g["row"]:
        assert O.row_partition(i, 5, p) == want
    for j, p, want in g["col"]:
        assert O.column_partition(5, j, p) == want


def test_java_random_known_answers():
    g = load("java_random")
    for seed, want in g["nextInt"].items():
        assert O.JavaRandom(int(seed)).next(32) == want
    for seed, want in g["nextDouble"].items():
        assert O.JavaRandom(int(seed)).nextDouble() == want
        assert O.JavaRandom(int(seed)).next_doubles(5)[0] == want
    # vectorised skip-ahead == sequential stream
    r1, r2 = O.JavaRandom(12345), O.JavaRandom(12345)
    seq = [r1.nextDouble() for _ in range(1000)]
    assert r2.next_doubles(1000).tolist() == seq
    assert r1.nextDouble() == r2.nextDouble()        # state advanced identically
    assert 0 <= O.JavaRandom(7).nextInt(10) < 10


def test_oracle_blocked_equals_dense_product():
    rng = np.random.default_rng(0)
    n, m, k, blk = 70, 50, 90, 32
    A = rng.uniform(-1, 1, (n, k))
    B = rng.uniform(-1, 1, (k, m))

    def blocks(M, tmask):
        out = {}
        for i in range(-(-M.shape[0] // blk)):
            for j in range(-(-M.shape[1] // blk)):
                sub = M[i * blk:(i + 1) * blk, j * blk:(j + 1) * blk]
                if tmask(i, j):
                    out[(i, j)] = O.DenseMatrix(sub.shape[0], sub.shape[1], np.ascontiguousarray(sub).reshape(-1), True)
                else:
                    out[(i, j)] = O.DenseMatrix(sub.shape[0], sub.shape[1], np.ascontiguousarray(sub.T).reshape(-1))
        return out
    C = O.matrix_multiply(blocks(A, lambda i, j: (i + j) % 2 == 0), n, k, blocks(B, lambda i, j: j % 2 == 1), k, m, blk)
    np.testing.assert_allclose(O.assemble(C, n, m, blk), A @ B, rtol=0, atol=1e-12)
    for blkm in C.values():
        assert not blkm.isTransposed


def test_oracle_requires_and_compat():
    a = {(0, 0): O.DenseMatrix(2, 2, [1, 2, 3, 4])}
    with pytest.raises(O.IllegalArgumentException, match="requirement failed: Matrix dimension not match, leftColNum = 4, rightRowNum = 6"):
        O.matrix_multiply(a, 4, 4, a, 6, 4, 2)
    with pytest.raises(O.IllegalArgumentException, match="Row number not match, leftRowNum = 4, rightRowNum = 5"):
        O.add_element(a, 4, 4, a, 5, 4, 2)
    with pytest.raises(O.IllegalArgumentException, match="The number of values supplied"):
        O.DenseMatrix(2, 2, [1, 2, 3])
    # B4: (Sparse, Dense) divide computes dense / sparse in compat mode
    s = O.SparseMatrix(2, 2, [0, 1, 2], [0, 1], [2.0, 4.0])
    d = O.DenseMatrix(2, 2, [8.0, 1.0, 1.0, 8.0])
    got = O.elementWiseDivide(s, d, compat_bugs=True).to_numpy()
    assert got[0, 0] == 4.0 and got[1, 1] == 2.0 and np.isinf(got[1, 0])
    got = O.elementWiseDivide(s, d, compat_bugs=False).to_numpy()
    assert got[0, 0] == 0.25 and got[1, 0] == 0.0
    # B3: rankOneAdd drops A in compat mode
    A = O.DenseMatrix(2, 2, [10, 20, 30, 40])
    x = O.DenseMatrix(2, 1, [1, 2])
    assert O.rankOneAdd(A, x, x, True).to_numpy().tolist() == [[1, 2], [2, 4]]
    assert O.rankOneAdd(A, x, x, False).to_numpy().tolist() == [[11, 32], [22, 44]]
    # scalar ops preserve flag and touch stored values only
    t = O.DenseMatrix(2, 3, np.arange(6.0), True)
    assert O.multiplyScalar(2.0, t).isTransposed and O.addScalar(t, 1.0).values.tolist() == (np.arange(6.0) + 1).tolist()
    assert O.addScalar(s, 1.0).to_numpy().tolist() == [[3.0, 0.0], [0.0, 5.0]]


def test_c_port_matches_numpy_oracle():
    """oracle/oracle.c (F2J-style dgemm restatement used for the timed CPU baseline) == numpy oracle."""
    from oracle import c_port
    if not c_port.available():
        import __graft_entry__ as g
        g.build()
    assert c_port.rand_block(1000, 42).tolist() == O.JavaRandom(42).next_doubles(1000).tolist()
    nb, blk = 3, 40
    A = O.rand_dense_dataset(nb * blk, nb * blk, blk, 42)
    B = O.rand_dense_dataset(nb * blk, nb * blk, blk, 43)
    got = c_port.block_multiply_f2j([A[(i, j)].values for i in range(nb) for j in range(nb)],
                                    [B[(i, j)].values for i in range(nb) for j in range(nb)], nb, blk, nb * nb, 2)
    want = O.matrix_multiply(A, nb * blk, nb * blk, B, nb * blk, nb * blk, blk)
    for t, c in enumerate(got):
        w = want[(t // nb, t % nb)].values
        assert np.max(np.abs(c - w)) / np.max(np.abs(w)) < 1e-14
