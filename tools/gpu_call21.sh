#!/bin/bash
# round 2, re-entry: full suite after the facade fix + the new grid operators (transpose, scalar maps, rowSum / colSum, project /
# selection), smoke, and one `ncu --set full` capture of the headline kernel at the bench size (DRAM traffic per launch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_full_suite3.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/t_full_suite3.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke3.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ozaki2_gemm_2sm -c 1 -f -o gpurun_out/prof_oz2sm_16384_r02 python tools/run_multiply.py 16384 1024 1 > gpurun_out/ncu_oz2sm_16384.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_oz2sm_16384.log
ls -l gpurun_out/prof_oz2sm_16384_r02.ncu-rep
