"""A `Dataset`-shaped object backed by the CPU oracle: the same method names and argument order as
matrel_b200.dataset.Dataset (M/Dataset.scala:57-152), so host-side logic that is written against that interface
(the planner, matrel_b200/plan.py) can be exercised without a GPU.  Test infrastructure only."""
from oracle import matrel_oracle as O


class OracleDataset:
    def __init__(self, blocks):
        self.blocks = dict(blocks)

    def _new(self, blocks):
        return OracleDataset(blocks)

    def matrixMultiply(self, lr, lc, right, rr, rc, blk):
        return self._new(O.matrix_multiply(self.blocks, lr, lc, right.blocks, rr, rc, blk))

    def addElement(self, lr, lc, right, rr, rc, blk):
        return self._new(O.add_element(self.blocks, lr, lc, right.blocks, rr, rc, blk))

    def multiplyElement(self, lr, lc, right, rr, rc, blk):
        return self._new(O.multiply_element(self.blocks, lr, lc, right.blocks, rr, rc, blk, compat_bugs=False))

    def divideElement(self, lr, lc, right, rr, rc, blk):
        return self._new(O.divide_element(self.blocks, lr, lc, right.blocks, rr, rc, blk, compat_bugs=False))

    def matrixRankOneUpdate(self, lr, lc, right, rr, rc, blk):
        return self._new(O.rank_one_update(self.blocks, lr, lc, right.blocks, rr, rc, blk, compat_bugs=False))

    def transpose(self):
        return self._new(O.transpose(self.blocks))

    t = transpose

    def addScalar(self, alpha):
        return self._new(O.add_scalar(self.blocks, alpha))

    def multiplyScalar(self, alpha):
        return self._new(O.multiply_scalar(self.blocks, alpha))

    def power(self, alpha):
        return self._new(O.power(self.blocks, alpha))

    def project(self, nrows, ncols, blkSize, rowOrCol, index):
        return self._new(O.project(self.blocks, nrows, ncols, blkSize, rowOrCol, index))

    def selection(self, nrows, ncols, blkSize, rowIdx, colIdx):
        return self._new(O.selection(self.blocks, nrows, ncols, blkSize, rowIdx, colIdx))

    def vec(self, nrows, ncols, blkSize):
        return self._new(O.vec(self.blocks, nrows, ncols, blkSize))

    def rowSum(self, nrows, ncols):
        return self._new(O.row_sum(self.blocks, nrows, ncols))

    def colSum(self, nrows, ncols):
        return self._new(O.col_sum(self.blocks, nrows, ncols))

    def sum(self, nrows, ncols):
        return self._new(O.total_sum(self.blocks, nrows, ncols))

    def trace(self, nrows, ncols):
        return self._new(O.trace(self.blocks, nrows, ncols))
