"""ctypes loader for oracle/liboracle.so (the C restatement; test infrastructure only)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle.so")


def available() -> bool:
    return os.path.exists(_PATH)


def _lib():
    lib = C.CDLL(_PATH)
    lib.oracle_rand_block.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lib.oracle_dgemm_f2j.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.oracle_block_multiply_f2j.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.oracle_block_multiply_f2j.restype = C.c_int
    lib.oracle_max_threads.restype = C.c_int
    return lib


def rand_block(n: int, seed: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    _lib().oracle_rand_block(out.ctypes.data, n, seed)
    return out


def block_multiply_f2j(A, B, nb: int, blk: int, ntasks: int, threads: int = 0, nk: int = 0):
    """A, B: lists (row-major over the nb x nb grid) of column-major blk*blk float64 arrays.
    Returns the list of the first ntasks output blocks."""
    lib = _lib()
    Cb = [np.zeros(blk * blk) for _ in range(nb * nb)]
    pa = (C.c_void_p * (nb * nb))(*[a.ctypes.data for a in A])
    pb = (C.c_void_p * (nb * nb))(*[b.ctypes.data for b in B])
    pc = (C.c_void_p * (nb * nb))(*[c.ctypes.data for c in Cb])
    rc = lib.oracle_block_multiply_f2j(nb, blk, pa, pb, pc, ntasks, threads, nk)
    if rc != 0:
        raise MemoryError("oracle_block_multiply_f2j: allocation failed")
    return Cb[:ntasks]


def max_threads() -> int:
    return _lib().oracle_max_threads()
