#!/bin/bash
# round 2, last call: full suite on the rebuilt library; the K-split item order of the CTA-pair GEMM (opt-in) under the bit-exact
# Ozaki-II / full-size tests, A/B against the default order on this box, DRAM bytes of its launch; the bench line with the faster order
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests -m gpu -q > gpurun_out/t_full_suite4.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/t_full_suite4.log
MATREL_OZ2_KSPLIT=2 timeout -k 5 100 python -m pytest tests/test_gpu_ozaki2.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/t_ksplit.log 2>&1; KRC=$?; echo "ksplit tests rc=$KRC"; tail -4 gpurun_out/t_ksplit.log
KS=0
if [ "$KRC" = "0" ]; then
  timeout -k 5 100 python tools/ab_ksplit.py 16384 1024 6 > gpurun_out/ab_ksplit.json 2> gpurun_out/ab_ksplit.err; echo "ab rc=$?"; cat gpurun_out/ab_ksplit.json; tail -2 gpurun_out/ab_ksplit.err
  KS=$(python -c "
import json
try:
    d = json.load(open('gpurun_out/ab_ksplit.json'))
    print(1 if d['bit_identical_block'] and d['speedup_gemm'] >= 1.02 and d['speedup_wall'] >= 1.01 else 0)
except Exception:
    print(0)")
fi
echo "chosen oz2_ksplit=$KS"
MATREL_OZ2_KSPLIT=$KS timeout -k 5 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench_h.json; tail -2 gpurun_out/bench_h.err
if [ "$KRC" = "0" ]; then
  MATREL_OZ2_KSPLIT=1 timeout -k 5 90 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum --clock-control none -k regex:ozaki2_gemm_2sm -c 1 --csv --log-file gpurun_out/ncu_ksplit_16384.csv python tools/run_multiply.py 16384 1024 1 > gpurun_out/ncu_ksplit.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/ncu_ksplit_16384.csv | cut -c1-300 | tail -9
fi
