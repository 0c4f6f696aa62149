"""BASELINE configs[4] shape on one GPU: N x N 1%-sparse (CSR blocks) x dense fp64, 1024-blocks.
   python tools/bench_spmm.py [N] [BLK] [density] [reps]
Reports ms, GFLOP/s (2*nnz*N) and the max deviation from scipy on sampled output blocks."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import matrel_b200 as mb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
nb = n // blk
rng = np.random.default_rng(0)
with mb.MatfastSession(device=0) as s:
    A = s.emptyDataset()
    host = {}
    nnz = 0
    for i in range(nb):
        for k in range(nb):
            m = sp.random(blk, blk, density=dens, format="csr", random_state=rng, dtype=np.float64)
            m.sort_indices()
            host[(i, k)] = m
            nnz += m.nnz
            A.put_block(i, k, mb.SparseMatrix(blk, blk, m.indptr, m.indices, m.data, True))   # CSR = isTransposed
    B = s.rand(n, n, blk, 43)
    s.set_option("time_kernels", 1)
    for r in range(reps):
        t0 = time.perf_counter()
        C = A.matrixMultiply(n, n, B, n, n, blk)
        s.sync()
        wall = (time.perf_counter() - t0) * 1e3
        st = s.stats()
        ms = st["last_gemm_ms"]
        print(json.dumps({"op": "spmm", "n": n, "blk": blk, "density": dens, "nnz": nnz, "kernel_ms": round(ms, 3),
                          "wall_ms": round(wall, 3), "GFLOPs": round(2.0 * nnz * n / (ms * 1e-3) / 1e9, 1)}))
        if r < reps - 1:
            del C
    worst = 0.0
    for (i, j) in [(0, 0), (nb - 1, nb // 2)]:
        want = sum(host[(i, k)] @ B.get_block(k, j).to_numpy() for k in range(nb))
        got = C.get_block(i, j).to_numpy()
        worst = max(worst, float(np.max(np.abs(got - want)) / np.max(np.abs(want))))
    print(json.dumps({"check": "vs scipy", "max_rel_err": worst}))
