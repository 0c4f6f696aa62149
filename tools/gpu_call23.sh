#!/bin/bash
# round 2, leftover seconds: memcheck over the single-process grid operators added at the end (C++ smoke binaries, no Python)
mkdir -p gpurun_out
g++ -std=c++17 -O1 -I include tests/cpp/grid_smoke.cpp -Lmatrel_b200 -lmatrel_b200 -Wl,-rpath,$PWD/matrel_b200 -o /tmp/grid_smoke || exit 1
timeout -k 3 40 compute-sanitizer --tool memcheck --error-exitcode 9 /tmp/grid_smoke 1 > gpurun_out/sanitizer_memcheck_gridops.log 2>&1; echo "memcheck grid_smoke rc=$?"
tail -4 gpurun_out/sanitizer_memcheck_gridops.log
