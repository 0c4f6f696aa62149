#!/bin/bash
# round 2, call 21: outlier elements corrected exactly instead of tripping the fallback
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki2.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/t_oz2e.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/t_oz2e.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_f.json; tail -2 gpurun_out/bench_f.err
