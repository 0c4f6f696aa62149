mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py -x -q -m gpu > gpurun_out/t_tf32.log 2>&1; echo "tf32 tests rc=$?"; tail -4 gpurun_out/t_tf32.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2b.json | cut -c1-1200; tail -3 gpurun_out/bench_n2b.err
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu -k "2-3 or 2-0" > gpurun_out/t_dist2.log 2>&1; echo "dist rc=$?"; tail -4 gpurun_out/t_dist2.log
