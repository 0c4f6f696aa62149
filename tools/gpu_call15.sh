#!/bin/bash
# round 2, call 17 (2 GPUs): alternating A / B pull pieces, gated pipelined ingest
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 1200 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_grid.py tests/test_gpu_cpp_facade.py -m gpu -x -q > gpurun_out/t_dist2c.log 2>&1; echo "dist tests rc=$?"; tail -8 gpurun_out/t_dist2c.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2c.json | cut -c1-300; tail -3 gpurun_out/bench_n2c.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2c.json') if l.startswith('{')][-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'check', d['check'], 'dmma', d['dmma_fp64'].get('value'))
P
