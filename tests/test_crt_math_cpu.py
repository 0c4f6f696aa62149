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
    assert math.prod(MODULI) < 1 << 126                           # four 40-bit fp64 chunks hold P and the weights (top chunk < 2^6)


def test_s32_accumulator_bound():
    assert ((1 << 17) - 1) * 128 * 128 < 1 << 31                   # K < 2^17 residue products of magnitude <= 2^14


CHUNK = int(re.search(r"CRT_CHUNK_BITS = (\d+)", SRC).group(1))


def crt_tables(T):
    """c_crtf[T] as crt_constants() builds it: the weights and P in CHUNK-bit fp64 chunks (three below 2^120, else four)."""
    mods = MODULI[:T]
    P = math.prod(mods)
    nch = 3 if P < 1 << (3 * CHUNK) else 4
    mask = (1 << CHUNK) - 1

    def chunks(x):
        return [float((x >> (CHUNK * i)) & mask) if i + 1 < nch else float(x >> (CHUNK * i)) for i in range(nch)]
    w = [(P // p) * pow((P // p) % p, -1, p) for p in mods]
    assert all(0 < wi < P for wi in w)
    return P, nch, np.array([chunks(wi) for wi in w]), np.array(chunks(P))


def crt_fp64(res, P, nch, W, Pc):
    """crt_kernel<NCH>, operation by operation, in numpy float64 (an FMA whose exact result fits 53 bits equals the two-step form)."""
    up, down, magic = np.float64(2.0 ** CHUNK), np.float64(2.0 ** -CHUNK), np.float64(1.5 * 2.0 ** 52)
    S = np.zeros(nch)
    for t, r in enumerate(res):
        S = S + np.float64(r) * W[t]
    assert np.all(S < 2.0 ** 52) and np.all(S == np.floor(S))      # exact integer chunk sums
    v = S[nch - 1]
    for i in range(nch - 2, -1, -1):
        v = v * up + S[i]
    q = (v * (np.float64(1.0) / np.float64(P)) + magic) - magic
    assert 0 <= q < 1 << 12 and np.all(q * Pc < 2.0 ** 52)
    D = list(S - q * Pc)
    for i in range(nch - 1):
        c = (D[i] * down + magic) - magic
        D[i] = D[i] - c * up
        D[i + 1] = D[i + 1] + c
        assert abs(D[i]) <= 2.0 ** (CHUNK - 1)
    x = D[nch - 1]
    for i in range(nch - 2, -1, -1):
        x = x * up + D[i]
    return float(x)


def test_fp64_crt_reconstructs_the_symmetric_residue():
    """The fp64 chunk arithmetic of crt_kernel returns the integer |c| <= 2^(floor(log2 P) - 1) its residues encode: exactly when
    c fits 53 bits (integer data give the exact product), else within one rounding (< 1 ulp) of it."""
    import random
    from fractions import Fraction
    random.seed(3)
    assert CHUNK == 40
    for T in range(T_MIN, T_MAX + 1):
        P, nch, W, Pc = crt_tables(T)
        lg = P.bit_length() - 1
        assert W[:, nch - 1].max() * T * 255 < 2.0 ** 52            # the top chunk sums stay exact as well
        bound = 1 << (lg - 1)
        assert bound * (1 + 2.0 ** -30) < P / 2                     # margin that keeps rint(S / P) on the right representative
        for trial in range(600):
            kind = trial % 6
            if kind == 0:
                c = random.randint(-bound, bound)
            elif kind == 1:
                c = random.choice([-1, 1]) * bound
            elif kind == 2:
                sh = random.randint(0, max(0, lg - 55))
                mb = min(53, lg - 1 - sh)
                c = random.randint(-(1 << mb), 1 << mb) << sh        # representable in fp64
            elif kind == 3:
                c = random.randint(-1000, 1000)
            elif kind == 4:
                c = random.choice([-1, 1]) * (bound - random.randint(0, 1000))
            else:
                c = random.randint(-bound, bound) >> random.randint(0, lg)
            x = crt_fp64([c % p for p in MODULI[:T]], P, nch, W, Pc)  # non-negative residues, as the planes store them
            if kind in (2, 3):
                assert x == float(c), (T, c, x)
            elif c:
                assert abs(Fraction(x) - c) < Fraction(math.ulp(float(c))), (T, c, x)


def test_auto_moduli_rule():
    """oz2_moduli_for(K, 0): the smallest T with alpha >= 54 - ceil(lg K / 2), i.e. sqrt(K) 2^-alpha <= K 2^-54."""
    for K in (256, 1000, 4096, 16384, 65536, (1 << 17) - 1):
        lgK = 0
        while (1 << lgK) < K:
            lgK += 1
        want = 54 - (lgK + 1) // 2
        T = next(t for t in range(T_MIN, T_MAX + 1) if alpha_of(t, K)[0] >= want)
        assert T == 14, (K, T)
        assert math.sqrt(K) * 2.0 ** -alpha_of(T, K)[0] <= K * 2.0 ** -54 * 1.5


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
