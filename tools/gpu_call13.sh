#!/bin/bash
# round 2, call 15: whole-warp-per-row spmm2 (256 x 64 tiles, 3 stages)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_spmm3.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_spmm3.log
timeout 300 python tools/bench_spmm.py 8192 1024 0.01 3 2>&1 | tail -4 | tee gpurun_out/spmm3.jsonl
timeout 300 python tools/bench_spmm.py 8192 1024 0.03 2 2>&1 | tail -3 | tee -a gpurun_out/spmm3.jsonl
timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 1 > gpurun_out/cfg5_e.json 2> gpurun_out/cfg5_e.err; echo "cfg5 rc=$?"; cut -c1-600 gpurun_out/cfg5_e.json; tail -2 gpurun_out/cfg5_e.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm2_kernel -c 1 -o gpurun_out/prof_spmm3_r02 python tools/bench_spmm.py 8192 1024 0.01 1 > gpurun_out/ncu_spmm3.log 2>&1; echo "ncu rc=$?"
