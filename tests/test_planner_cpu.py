"""CPU tests of the planner rewrites (matrel_b200/plan.py, restating M/execution/MatfastPlanner.scala:42-279).

The planner is written against the `Dataset` method interface, so here it drives an oracle-backed stand-in
(tests/oracle_dataset.py): for random expression trees the rewritten plan, the un-rewritten plan and a plain numpy
evaluation of the same expression must agree.  The same identities are checked on the device in
tests/test_gpu_parity.py."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from matrel_b200 import plan as P
from oracle import matrel_oracle as O
from tests.oracle_dataset import OracleDataset


def leaf_blocks(rng, n, blk, flags_seed):
    """Full dense n x n matrix in blk-blocks (ragged last block), random isTransposed per block, values in [0.5, 1.5]."""
    full = rng.uniform(0.5, 1.5, (n, n))
    frng = np.random.default_rng(flags_seed)
    blocks = {}
    for i in range(-(-n // blk)):
        for j in range(-(-n // blk)):
            sub = full[i * blk:(i + 1) * blk, j * blk:(j + 1) * blk]
            if frng.random() < 0.5:
                blocks[(i, j)] = O.DenseMatrix(sub.shape[0], sub.shape[1], np.ascontiguousarray(sub).reshape(-1), True)
            else:
                blocks[(i, j)] = O.DenseMatrix(sub.shape[0], sub.shape[1], np.ascontiguousarray(sub.T).reshape(-1), False)
    return full, blocks


# expression trees over square n x n matrices: ("leaf", k) | ("t", e) | ("adds", e, a) | ("muls", e, a) |
# ("add", e, e) | ("mul", e, e) | ("mm", e, e)
def exprs(nleaves):
    leaf = st.tuples(st.just("leaf"), st.integers(0, nleaves - 1))
    alpha = st.sampled_from([0.5, 2.0, -1.25, 3.0])
    return st.recursive(
        leaf,
        lambda ch: st.one_of(
            st.tuples(st.just("t"), ch),
            st.tuples(st.just("adds"), ch, alpha),
            st.tuples(st.just("muls"), ch, alpha),
            st.tuples(st.just("add"), ch, ch),
            st.tuples(st.just("mul"), ch, ch),
            st.tuples(st.just("mm"), ch, ch)),
        max_leaves=5)


def build(e, leaves, n, blk):
    """-> (plan node, numpy value)"""
    k = e[0]
    if k == "leaf":
        full, blocks = leaves[e[1]]
        return P.Leaf(OracleDataset(blocks)), full
    if k == "t":
        c, v = build(e[1], leaves, n, blk)
        return P.TransposeOperator(c), v.T
    if k == "adds":
        c, v = build(e[1], leaves, n, blk)
        return P.MatrixScalarAddOperator(c, e[2]), v + e[2]
    if k == "muls":
        c, v = build(e[1], leaves, n, blk)
        return P.MatrixScalarMultiplyOperator(c, e[2]), v * e[2]
    l, lv = build(e[1], leaves, n, blk)
    r, rv = build(e[2], leaves, n, blk)
    if k == "add":
        return P.MatrixElementAddOperator(l, n, n, r, n, n, blk), lv + rv
    if k == "mul":
        return P.MatrixElementMultiplyOperator(l, n, n, r, n, n, blk), lv * rv
    return P.MatrixMatrixMultiplicationOperator(l, n, n, r, n, n, blk), lv @ rv


ROOTS = ["none", "rowSum", "colSum", "sum", "trace", "projectRow", "projectCol", "select"]


def wrap(root, node, val, n, blk, idx, jdx):
    if root == "none":
        return node, val, (n, n)
    if root == "rowSum":
        return P.RowSumOperator(node, n, n), val.sum(axis=1, keepdims=True), (n, 1)
    if root == "colSum":
        return P.ColumnSumOperator(node, n, n), val.sum(axis=0, keepdims=True), (1, n)
    if root == "sum":
        return P.SumOperator(node, n, n), val.sum().reshape(1, 1), (1, 1)
    if root == "trace":
        return P.TraceOperator(node, n, n), np.trace(val).reshape(1, 1), (1, 1)
    if root == "projectRow":
        return P.ProjectOperator(node, n, n, blk, True, idx), val[idx:idx + 1, :], (1, n)
    if root == "projectCol":
        return P.ProjectOperator(node, n, n, blk, False, jdx), val[:, jdx:jdx + 1], (n, 1)
    return P.SelectOperator(node, n, n, blk, idx, jdx), val[idx, jdx].reshape(1, 1), (1, 1)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(e=exprs(3), root=st.sampled_from(ROOTS), shape=st.sampled_from([(8, 4), (10, 4), (9, 3), (6, 6), (7, 16)]),
       seed=st.integers(0, 2 ** 16), i=st.integers(0, 5), j=st.integers(0, 5))
def test_rewritten_plan_equals_direct_plan_and_numpy(e, root, shape, seed, i, j):
    n, blk = shape
    rng = np.random.default_rng(seed)
    leaves = [leaf_blocks(rng, n, blk, seed + 17 * k) for k in range(3)]
    node, val = build(e, leaves, n, blk)
    node, val, (r, c) = wrap(root, node, val, n, blk, i % n, j % n)
    p1, p0 = P.Planner(True), P.Planner(False)
    got = O.assemble(p1.run(node).blocks, r, c, blk)
    want = O.assemble(p0.run(node).blocks, r, c, blk)
    scale = max(1.0, float(np.max(np.abs(val))))
    assert np.max(np.abs(want - val)) <= 1e-10 * scale, ("direct plan vs numpy", e, root)
    assert np.max(np.abs(got - val)) <= 1e-10 * scale, ("rewritten plan vs numpy", e, root, p1.trace)


def test_rewrites_remove_the_cubic_product():
    """trace(AB) -> sum(A^T o B); sum(AB) -> colSum(A) . rowSum(B); rowSum(AB) -> A . rowSum(B); select(AB) -> row . col
    (MatfastPlanner.scala:125-166, 168-244): the physical plan holds no n x n x n multiply any more."""
    n, blk = 12, 4
    rng = np.random.default_rng(3)
    (fa, ba), (fb, bb) = leaf_blocks(rng, n, blk, 1), leaf_blocks(rng, n, blk, 2)
    calls = []

    class Spy(OracleDataset):
        def _new(self, blocks):
            return Spy(blocks)

        def matrixMultiply(self, lr, lc, right, rr, rc, blk_):
            calls.append((lr, lc, rr, rc))
            return super().matrixMultiply(lr, lc, right, rr, rc, blk_)

    A, B = P.Leaf(Spy(ba)), P.Leaf(Spy(bb))
    AB = P.MatrixMatrixMultiplicationOperator(A, n, n, B, n, n, blk)
    cases = {
        "trace": (P.TraceOperator(AB, n, n), np.trace(fa @ fb)),
        "sum": (P.SumOperator(AB, n, n), (fa @ fb).sum()),
        "select": (P.SelectOperator(AB, n, n, blk, 7, 2), (fa @ fb)[7, 2]),
    }
    for name, (node, want) in cases.items():
        calls.clear()
        out = P.Planner(True).run(node).blocks
        assert abs(out[(0, 0)].values[0] - want) <= 1e-10 * abs(want), name
        assert all(min(lr, rc) == 1 for lr, lc, rr, rc in calls), (name, calls)   # vector-sized products only
    calls.clear()
    out = P.Planner(True).run(P.RowSumOperator(AB, n, n)).blocks
    assert np.allclose(O.assemble(out, n, 1, blk)[:, 0], (fa @ fb).sum(axis=1), rtol=1e-12)
    assert calls == [(n, n, n, 1)]
