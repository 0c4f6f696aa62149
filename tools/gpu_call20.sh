#!/bin/bash
# round 2, final validation on one GPU: full suite, smoke, driver-style bench + reference arm, N = 65536 on one GPU, sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_full_suite2.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/t_full_suite2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_g.json; tail -2 gpurun_out/bench_g.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_g.json
timeout 400 python tools/bench_sizes.py 65536 2048 0 1 2> gpurun_out/sizes_n1c.err | grep '^{' | tee gpurun_out/sizes_r02_n1c.jsonl | cut -c1-400; tail -2 gpurun_out/sizes_n1c.err
timeout 900 bash tools/sanitize.sh gpurun_out > gpurun_out/sanitizer_r02b.txt 2>&1; cat gpurun_out/sanitizer_r02b.txt | cut -c1-200
