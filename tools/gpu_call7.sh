mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tf32.py -x -q -m gpu > gpurun_out/t_tf32.log 2>&1; echo "tf32 tests rc=$?"; tail -8 gpurun_out/t_tf32.log
: > gpurun_out/sizes_r02_n1.jsonl
for cfg in "4096 512 0" "4096 512 1" "16384 1024 3" "65536 2048 0 1"; do timeout 900 python tools/bench_sizes.py $cfg >> gpurun_out/sizes_r02_n1.jsonl 2>> gpurun_out/sizes_n1.err; done
cat gpurun_out/sizes_r02_n1.jsonl; tail -3 gpurun_out/sizes_n1.err
