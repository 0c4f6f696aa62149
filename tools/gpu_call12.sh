#!/bin/bash
# round 2, call 14: fp64-chunk CRT kernel, wide-tile absmax, moduli chosen from K
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki2.py tests/test_gpu_ozaki.py tests/test_gpu_tf32.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/t_oz2d.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/t_oz2d.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file gpurun_out/launches_oz2_r02c.csv python tools/run_multiply.py 16384 1024 3 > gpurun_out/launches_oz2c.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_oz2_r02c.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:]:
    if 'copy_words' not in r[ik]: print(r[ik][:60], r[iv])
P
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_e.json; tail -2 gpurun_out/bench_e.err
