mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py -x -q -m gpu > gpurun_out/t_spmm2.log 2>&1; echo "spmm tests rc=$?"; tail -6 gpurun_out/t_spmm2.log
timeout 600 python tools/bench_spmm.py 8192 1024 0.01 3 > gpurun_out/spmm_r02b.jsonl 2>&1; cat gpurun_out/spmm_r02b.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_oz2_r02.csv python tools/run_multiply.py 16384 1024 3 > gpurun_out/launches_oz2.log 2>&1; tail -2 gpurun_out/launches_oz2.log
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_oz2_r02.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:70], r[iv])
P
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/t_full.log 2>&1; echo "fullsize rc=$?"; tail -8 gpurun_out/t_full.log
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 2 > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo "cfg5 rc=$?"; cut -c1-2500 gpurun_out/bench_cfg5_n1.json; tail -5 gpurun_out/bench_cfg5_n1.err
