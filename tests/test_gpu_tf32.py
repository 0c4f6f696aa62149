"""GPU tests of the fp32 multiply on tcgen05 kind::tf32 (gemm_algo = 3, 3xTF32 split) -- BASELINE configs[3] shape.

fp32 is not a reference feature (blocks are Array[Double]); parity is defined as: inputs rounded to fp32, oracle =
fp64 product of those inputs, tolerance 2e-5 relative to max|C| (fp32 accumulation over K plus the dropped lo*lo term;
a single-pass TF32 product would be ~1e-3)."""
import numpy as np
import pytest

import matrel_b200 as mb
from oracle import matrel_oracle as O
from tests.util import assert_same_dataset, from_dataset, random_block_dataset, rel_err, to_dataset

pytestmark = pytest.mark.gpu
TF32X3_TOL = 2e-5


def f32_round(ds):
    out = {}
    for k, m in ds.items():
        out[k] = O.DenseMatrix(m.numRows, m.numCols, m.values.astype(np.float32).astype(np.float64), m.isTransposed)
    return out


@pytest.mark.parametrize("n,k,m,blk,pt", [(256, 256, 256, 128, 0.0), (512, 640, 384, 128, 0.5), (300, 200, 260, 128, 0.5),
                                          (2048, 2048, 2048, 512, 0.3)])
def test_tf32x3_multiply(n, k, m, blk, pt):
    rng = np.random.default_rng(n + k + m)
    A = f32_round(random_block_dataset(rng, n, k, blk, p_transposed=pt))
    B = f32_round(random_block_dataset(rng, k, m, blk, p_transposed=pt))
    want = O.matrix_multiply(A, n, k, B, k, m, blk)
    with mb.MatfastSession(device=0, gemm_algo=3) as s:
        got = from_dataset(to_dataset(s, A).matrixMultiply(n, k, to_dataset(s, B), k, m, blk))
        assert s.stats()["kernel_launches"] == 3            # two slicing passes + ONE tcgen05 launch
    assert_same_dataset(got, want, tol=TF32X3_TOL)          # ids / presence / shapes / flags exact
    full_g = O.assemble({k_: O.DenseMatrix(v.numRows, v.numCols, v.values) for k_, v in got.items()}, n, m, blk)
    err = rel_err(full_g, O.assemble(want, n, m, blk))
    assert err <= TF32X3_TOL, err
    assert err > 0                                           # it really is fp32 arithmetic
    # every stored result is an fp32 value
    for v in got.values():
        assert np.array_equal(v.values, v.values.astype(np.float32).astype(np.float64))
