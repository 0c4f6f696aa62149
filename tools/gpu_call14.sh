#!/bin/bash
# round 2, call 16: spmm2 oversized-segment path (coalesced batches); densities 1 / 3 / 6 %
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py -m gpu -x -q > gpurun_out/t_spmm4.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_spmm4.log
for d in 0.01 0.03 0.06; do timeout 300 python tools/bench_spmm.py 8192 1024 $d 2 2>&1 | tail -2 | tee -a gpurun_out/spmm4.jsonl; done
