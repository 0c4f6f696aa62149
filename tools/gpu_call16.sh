#!/bin/bash
# round 2, call 18 (2 GPUs): own blocks ordered with their piece; e2e with 4 and 8 pieces
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "nccl and 2-0" > gpurun_out/t_dist2d.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/t_dist2d.log
for ch in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$ch bench.py --gpus 2 --steps 5 --warmup 3 --e2e-chunks $ch > gpurun_out/bench_n2_e$ch.json 2> gpurun_out/bench_n2_e$ch.err; echo "bench n2 rc=$?"
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_n2_e$ch.json') if l.startswith('{')][-1])
print('pieces $ch: N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'check', d['check']['e2e_max_rel_err_vs_host_fp64'])
P
done
