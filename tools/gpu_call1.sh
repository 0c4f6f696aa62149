mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi1.txt 2>&1
timeout 180 ./tools/tc_peak > gpurun_out/tc_peaks_r02.jsonl 2> gpurun_out/tc_peak.err; echo "tc_peak rc=$?"
timeout 900 python -m pytest tests/test_gpu_ozaki2.py -x -q -m gpu > gpurun_out/t_oz2.log 2>&1; echo "oz2 rc=$?"; tail -5 gpurun_out/t_oz2.log
timeout 1500 python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_ozaki2.py > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/t_all.log
MATREL_E2E_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
cat gpurun_out/bench_a.json | head -c 3000; tail -12 gpurun_out/bench_a.err
cat gpurun_out/tc_peaks_r02.jsonl
