"""GPU tests of the Ozaki scheme II fp64 multiply (gemm_algo = 4: int8 tcgen05 GEMMs of residue matrices + CRT).

The integer product of the scaled operands is recovered EXACTLY, so (a) integer-valued inputs give bit-exact
results, and (b) the only error is the rounding of the operands to alpha bits relative to their row / column
maximum plus one final rounding: with 16 moduli (alpha >= 53 for K <= 2^16) the result is at least as accurate
as an fp64 dgemm for U(-1,1) data.  Bars: bit-exact for integer data; 1e-14 relative to |A||B| at 16 moduli."""
import numpy as np
import pytest

import matrel_b200 as mb
from oracle import matrel_oracle as O
from tests.util import REL_TOL, assert_same_dataset, from_dataset, random_block_dataset, rel_err, to_dataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oz2():
    s = mb.MatfastSession(device=0, gemm_algo=4)
    yield s
    s.stop()


def full(ds, n, m, blk):
    out = np.zeros((n, m))
    for (i, j), b in ds.items():
        out[i * blk:i * blk + b.numRows, j * blk:j * blk + b.numCols] = b.to_numpy()
    return out


def blocks_of(M, blk):
    n, m = M.shape
    out = {}
    for i in range((n + blk - 1) // blk):
        for j in range((m + blk - 1) // blk):
            sub = M[i * blk:(i + 1) * blk, j * blk:(j + 1) * blk]
            out[(i, j)] = O.DenseMatrix(sub.shape[0], sub.shape[1], np.ascontiguousarray(sub.T).reshape(-1))
    return out


@pytest.mark.parametrize("n,k,m,blk,pt", [
    (256, 256, 256, 128, 0.0),
    (512, 384, 640, 128, 0.5),      # all four T/N combinations, several 128x256 tiles
    (300, 200, 260, 128, 0.5),      # ragged edges
    (131, 77, 93, 64, 0.5),         # odd dims: byte-wise residue stores at block edges
    (1024, 1024, 1024, 256, 0.3),   # BASELINE config[0] shape
])
def test_crt_multiply_vs_oracle(oz2, n, k, m, blk, pt):
    rng = np.random.default_rng(n + 3 * k + 7 * m + blk)
    A = random_block_dataset(rng, n, k, blk, p_transposed=pt)
    B = random_block_dataset(rng, k, m, blk, p_transposed=pt)
    want = O.matrix_multiply(A, n, k, B, k, m, blk)
    oz2.reset_stats()
    got = from_dataset(to_dataset(oz2, A).matrixMultiply(n, k, to_dataset(oz2, B), k, m, blk))
    assert oz2.stats()["kernel_launches"] >= 7             # absmax x2, exp, residues x2, ONE GEMM launch, CRT
    assert_same_dataset(got, want, tol=REL_TOL)            # ids / presence / shapes / flags exact
    err = rel_err(full(got, n, m, blk), full(want, n, m, blk))
    assert err <= 1e-14, err


@pytest.mark.parametrize("n,k,m,blk", [(256, 384, 256, 128), (200, 333, 150, 64)])
def test_crt_is_bit_exact_on_integer_data(oz2, n, k, m, blk):
    """Integers below 2^20: every product and sum is exact in fp64 AND in the residue arithmetic -> identical bits."""
    rng = np.random.default_rng(5)
    Af = rng.integers(-(1 << 20), 1 << 20, (n, k)).astype(np.float64)
    Bf = rng.integers(-(1 << 20), 1 << 20, (k, m)).astype(np.float64)
    got = full(from_dataset(to_dataset(oz2, blocks_of(Af, blk)).matrixMultiply(n, k, to_dataset(oz2, blocks_of(Bf, blk)), k, m, blk)), n, m, blk)
    want = (Af.astype(object) @ Bf.astype(object)).astype(np.float64)   # exact integer product (|c| < 2^50)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("moduli,bound", [(8, 2e-7), (10, 1e-9), (12, 5e-12), (14, 2e-14), (16, 5e-15)])
def test_crt_error_scales_with_moduli(moduli, bound):
    rng = np.random.default_rng(moduli)
    n, blk = 512, 128
    A = random_block_dataset(rng, n, n, blk, lo=0.0, hi=1.0)
    B = random_block_dataset(rng, n, n, blk, lo=0.0, hi=1.0)
    want = full(O.matrix_multiply(A, n, n, B, n, n, blk), n, n, blk)
    with mb.MatfastSession(device=0, gemm_algo=4) as s:
        s.set_option("crt_moduli", moduli)
        got = full(from_dataset(to_dataset(s, A).matrixMultiply(n, n, to_dataset(s, B), n, n, blk)), n, n, blk)
    assert rel_err(got, want) <= bound, rel_err(got, want)


def test_crt_wide_dynamic_range_rows_and_columns(oz2):
    rng = np.random.default_rng(17)
    n, blk = 384, 128
    Af = rng.uniform(-1, 1, (n, n)) * np.exp2(rng.integers(-40, 40, n))[:, None]
    Bf = rng.uniform(-1, 1, (n, n)) * np.exp2(rng.integers(-40, 40, n))[None, :]
    got = full(from_dataset(to_dataset(oz2, blocks_of(Af, blk)).matrixMultiply(n, n, to_dataset(oz2, blocks_of(Bf, blk)), n, n, blk)), n, n, blk)
    want = Af @ Bf
    scale = np.abs(Af) @ np.abs(Bf)
    assert np.max(np.abs(got - want) / scale) <= 1e-14


def test_crt_block_sparse_presence_and_zero_rows(oz2):
    rng = np.random.default_rng(23)
    n, blk = 5 * 64, 64
    A = random_block_dataset(rng, n, n, blk, density=0.5, p_transposed=0.3)
    B = random_block_dataset(rng, n, n, blk, density=0.5, p_transposed=0.3)
    key = next(iter(A))
    A[key] = O.DenseMatrix(A[key].numRows, A[key].numCols, np.zeros(A[key].numRows * A[key].numCols))  # all-zero block
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    got = from_dataset(to_dataset(oz2, A).matrixMultiply(n, n, to_dataset(oz2, B), n, n, blk))
    assert_same_dataset(got, want, tol=1e-13)


def test_crt_nonfinite_falls_back_to_exact_kernel(oz2):
    n, blk = 256, 128
    A = O.rand_dense_dataset(n, n, blk, 1)
    B = O.rand_dense_dataset(n, n, blk, 2)
    A[(0, 0)].values[5] = np.nan
    want = O.matrix_multiply(A, n, n, B, n, n, blk)
    got = from_dataset(to_dataset(oz2, A).matrixMultiply(n, n, to_dataset(oz2, B), n, n, blk))
    assert_same_dataset(got, want, tol=1e-12)


def test_crt_matches_dmma_at_4096(oz2, session):
    n, blk = 4096, 512
    A1, B1 = session.rand(n, n, blk, 42), session.rand(n, n, blk, 43)
    A2, B2 = oz2.rand(n, n, blk, 42), oz2.rand(n, n, blk, 43)
    C1 = A1.matrixMultiply(n, n, B1, n, n, blk)
    C2 = A2.matrixMultiply(n, n, B2, n, n, blk)
    worst = 0.0
    for key in C1.block_ids():
        a, b = C1.get_block(*key).values, C2.get_block(*key).values
        worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(a))))
    assert worst <= 2e-14, worst      # the deviation is the DMMA kernel's own accumulated rounding (K = 4096)
