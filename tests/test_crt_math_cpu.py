"""CPU checks of the integer arithmetic the Ozaki-II kernels (matrel_b200/csrc/gemm_ozaki.cu, gemm_algo = 4) rely on.
Nothing here runs the product: the moduli are read from the source, the claims in its comments are re-derived with
Python big integers / numpy float32 emulation of the device expressions."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "matrel_b200", "csrc", "gemm_ozaki.cu")).read()
MODULI = [int(x) for x in re.search(r"kCrtModuli\[CRT_MAX_T\] = \{([^}]*)\}", SRC).group(1).split(",")]
T_MIN, T_MAX = int(re.search(r"CRT_MIN_T = (\d+)", SRC).group(1)), int(re.search(r"CRT_MAX_T = (\d+)", SRC).group(1))


def test_moduli_fit_int8_and_are_pairwise_coprime():
    assert len(MODULI) == T_MAX == 16 and T_MIN >= 2
    assert all(2 <= p <= 256 for p in MODULI)                    # symmetric residues fit int8
    for i, a in enumerate(MODULI):
        for b in MODULI[i + 1:]:
            assert math.gcd(a, b) == 1, (a, b)


def alpha_of(T, K):
    P = math.prod(MODULI[:T])
    lgK = 0
    while (1 << lgK) < K:
        lgK += 1
    return min(62, (P.bit_length() - 1 - 1 - lgK) // 2), P


def test_alpha_rule_keeps_the_integer_product_inside_the_crt_range():
    """alpha = floor((floor(log2 P) - 1 - ceil(log2 K)) / 2): K * (2^alpha)^2 < P / 2 for every supported T and K < 2^17, so the
    symmetric CRT reconstruction is the exact integer product; 16 moduli give >= 53 operand bits up to K = 2^17."""
    for T in range(T_MIN, T_MAX + 1):
        for K in (1, 2, 3, 1000, 16384, 65536, (1 << 17) - 1):
            a, P = alpha_of(T, K)
            assert 2 * K * (1 << a) ** 2 < P, (T, K, a)
            assert a >= 8
    assert alpha_of(16, 16384)[0] == 55 and alpha_of(14, 16384)[0] == 47 and alpha_of(16, (1 << 17) - 1)[0] >= 53
    assert math.prod(MODULI) < 1 << 126                           # the kernel's four 32-bit limbs hold P and the weights


def test_s32_accumulator_bound():
    assert ((1 << 17) - 1) * 128 * 128 < 1 << 31                   # K < 2^17 residue products of magnitude <= 2^14


def test_crt_weights_reconstruct_every_residue_vector():
    rng = np.random.default_rng(0)
    for T in (T_MIN, 9, T_MAX):
        mods = MODULI[:T]
        P = math.prod(mods)
        w = [(P // p) * pow((P // p) % p, -1, p) for p in mods]
        assert all(0 < wi < P for wi in w)
        for _ in range(200):
            x = int(rng.integers(-(1 << 62), 1 << 62)) * int(rng.integers(1, 1 << 40)) % P
            x = x - P if x > P // 2 else x
            acc = sum((x % p) * wi for p, wi in zip(mods, w))      # non-negative residues, as the planes store them
            assert acc < T * 256 * P < 1 << 140                    # five 32-bit limbs after carry normalisation
            y = acc % P
            y = y - P if y > P // 2 else y
            assert y == x


def test_dp4a_fp32_residue_is_congruent_and_fits_int8():
    """residue_kernel: x = sum_i byte_i * (256^i mod p) < 2^19 is exact in fp32; r = x - rint(fp32(x) * fp32(1/p)) * p is congruent
    to x and |r| <= 128 (p = 256) / <= 127 (odd p) -- checked for EVERY x the byte fold can produce."""
    xs = np.arange(0, 8 * 255 * 255 + 1, dtype=np.int64)
    xf = xs.astype(np.float32)
    assert np.array_equal(xf.astype(np.int64), xs)                 # exact in fp32
    for p in MODULI:
        coeffs = [pow(256, i, p) for i in range(8)]
        assert all(c < 256 for c in coeffs)                        # one byte each for dp4a
        assert 255 * sum(coeffs) <= xs[-1]
        inv = np.float32(1.0) / np.float32(p)
        q = np.rint(xf * inv).astype(np.int64)                     # __float2int_rn(__uint2float_rn(x) * ip)
        r = xs - q * p
        assert np.all((r - xs) % p == 0)
        assert np.max(np.abs(r)) <= (128 if p == 256 else 127), (p, int(np.max(np.abs(r))))


def test_fp64_epilogue_residue_is_exact():
    """GEMM epilogue: res = c - rint(c / p) * p with c / p evaluated as double(c) * (1.0 / p): the nearest integer is the true
    one for every |c| < 2^31 because no integer multiple of 1/p lies within 2^-20 of a tie (spot-checked densely near ties)."""
    rng = np.random.default_rng(1)
    for p in MODULI:
        inv = 1.0 / p
        c = np.concatenate([rng.integers(-(1 << 31) + 1, (1 << 31) - 1, 200000),
                            (rng.integers(-(1 << 22), 1 << 22, 50000) * p + p // 2),          # just below / at the tie
                            (rng.integers(-(1 << 22), 1 << 22, 50000) * p + (p + 1) // 2)]).astype(np.int64)
        c = c[np.abs(c) < (1 << 31)]
        q = np.rint(c.astype(np.float64) * inv).astype(np.int64)
        res = c - q * p
        assert np.all((res - c) % p == 0)
        assert np.max(np.abs(res)) <= p // 2                       # [-p/2, p/2]; +-128 only for p = 256
        stored = np.where(res < 0, res + p, res)                   # res += (res >> 31) & p
        assert stored.min() >= 0 and stored.max() <= 255 and np.all((stored - c) % p == 0)
