"""CPU re-derivation of the shared-memory layout claims of the fp64 DMMA kernel (matrel_b200/csrc/gemm_f64.cu):
the formulas are read from the source and checked exhaustively -- a wrong permutation or swizzle would still give correct
products (the tests on the GPU cannot see it) but would bring the bank conflicts back."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "matrel_b200", "csrc", "gemm_f64.cu")).read()


def c_expr(name):
    body = re.search(name + r"\(int (\w)\) \{ return ([^;]*); \}", SRC)
    var, expr = body.group(1), body.group(2)
    return lambda v: eval(expr, {}, {var: v})          # plain C integer expression with >>, &, +, *


kpos_of_k = c_expr("kpos_of_k")
k_of_kpos = c_expr("k_of_kpos")
BK = 16


def kset(s):                                            # k values of MMA step s, in thread-in-group order t = 0..3
    return [2 * s + (t & 1) + 8 * (t >> 1) for t in range(4)]


def test_ksets_partition_the_stage():
    assert "k_t = 2s + (t&1) + 8*(t>>1)" in SRC
    assert sorted(k for s in range(4) for k in kset(s)) == list(range(BK))


def test_padded_layout_permutation_makes_each_kset_four_consecutive_rows():
    assert sorted(kpos_of_k(k) for k in range(BK)) == list(range(BK))            # a permutation
    assert all(k_of_kpos(kpos_of_k(k)) == k for k in range(BK))                 # with the inverse the producer uses
    for s in range(4):
        assert [kpos_of_k(k) for k in kset(s)] == [4 * s, 4 * s + 1, 4 * s + 2, 4 * s + 3]


def test_swizzled_layout_gives_a_half_warp_16_distinct_8_byte_slots():
    """MODE_K tile: rows of 128 bytes (16 doubles of k), 16-byte chunk index ^= row & 7.  A half-warp of an m8n8k4 fragment
    load reads 4 consecutive rows x the 4 k of the step: all 16 accesses must fall into different 8-byte bank pairs."""
    for s in range(4):
        for row0 in range(0, 128, 4):                                           # half-warps start at multiples of 4 rows
            slots = set()
            for g in range(4):
                r = row0 + g
                for k in kset(s):
                    chunk = (k >> 1) ^ (r & 7)                                  # SWIZZLE_128B
                    slots.add(2 * chunk + (k & 1))                              # 8-byte slot inside the 128-byte row
            assert len(slots) == 16, (s, row0)
