mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_grid.py tests/test_gpu_cpp_facade.py -x -q -m gpu -k "8 or facade or without_python" > gpurun_out/t_dist8.log 2>&1; echo "dist8 tests rc=$?"; tail -6 gpurun_out/t_dist8.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; grep '^{' gpurun_out/bench_n8.json | cut -c1-3500; tail -3 gpurun_out/bench_n8.err
: > gpurun_out/sizes_r02_n8.jsonl
for cfg in "4096 512 0" "4096 512 1" "65536 2048 0 2" "65536 2048 1 1" "65536 2048 3 2" "65536 2048 3 2 8x1"; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 tools/bench_sizes.py $cfg 2>> gpurun_out/sizes_n8.err | grep '^{' >> gpurun_out/sizes_r02_n8.jsonl
done
cat gpurun_out/sizes_r02_n8.jsonl; tail -5 gpurun_out/sizes_n8.err
