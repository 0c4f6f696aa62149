mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_spmm.log 2>&1; echo "spmm tests rc=$?"; tail -12 gpurun_out/t_spmm.log
timeout 600 python tools/bench_spmm.py 8192 1024 0.01 3 > gpurun_out/spmm_r02.jsonl 2>&1; cat gpurun_out/spmm_r02.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
