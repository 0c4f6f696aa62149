"""Tiny driver for ncu captures: python tools/run_multiply.py N BLK [reps] [variant]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matrel_b200 as mb

n, blk = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with mb.MatfastSession(device=0) as s:
    if len(sys.argv) > 4 and int(sys.argv[4]) >= 0:
        s.set_option("gemm_variant", int(sys.argv[4]))
    for kv in sys.argv[5:]:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    A, B = s.rand(n, n, blk, 42), s.rand(n, n, blk, 43)
    s.set_option("time_kernels", 1)
    for _ in range(reps):
        C = A.matrixMultiply(n, n, B, n, n, blk)
        st = s.stats()
        print(f"n={n} blk={blk}: gemm {st['last_gemm_ms']:.3f} ms  {st['last_gemm_flops'] / st['last_gemm_ms'] / 1e9:.2f} TFLOP/s")
        del C
