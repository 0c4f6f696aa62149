#!/bin/bash
# round 2, call 20 (8 GPUs): headline at N=8 with the final kernels; world-8 parity of the tcgen05 and fp32 paths
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8c.json 2> gpurun_out/bench_n8c.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/bench_n8c.err
python - <<'P'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8c.json') if l.startswith('{')][-1])
    print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d.get('e2e'), 'check', d.get('check'), 'dmma', (d.get('dmma_fp64') or {}).get('value'))
except Exception as e: print('ERR', e)
P
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "nccl and (8-0 or 8-3)" > gpurun_out/t_dist8c.log 2>&1; echo "dist8 rc=$?"; tail -4 gpurun_out/t_dist8c.log
: > gpurun_out/sizes_r02_n8c.jsonl
for cfg in "4096 512 0" "65536 2048 0 2"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 tools/bench_sizes.py $cfg 2>> gpurun_out/sizes_n8c.err | grep '^{' >> gpurun_out/sizes_r02_n8c.jsonl
done
cat gpurun_out/sizes_r02_n8c.jsonl | cut -c1-400; tail -3 gpurun_out/sizes_n8c.err
