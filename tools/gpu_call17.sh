#!/bin/bash
# round 2, call 19 (4 GPUs): headline at N=4 and configs[4] on 4 GPUs with the final kernels
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m pytest tests/test_gpu_spmm.py -m gpu -x -q > gpurun_out/t_spmm5.log 2>&1; echo "spmm tests rc=$?"; tail -3 gpurun_out/t_spmm5.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4c.json 2> gpurun_out/bench_n4c.err; echo "bench n4 rc=$?"; tail -2 gpurun_out/bench_n4c.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/cfg5_n4c.json 2> gpurun_out/cfg5_n4c.err; echo "cfg5 n4 rc=$?"; tail -2 gpurun_out/cfg5_n4c.err
python - <<'P'
import json
for f in ('gpurun_out/bench_n4c.json','gpurun_out/cfg5_n4c.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d.get('e2e'), 'check', d.get('check'))
    except Exception as e: print(f, 'ERR', e)
P
