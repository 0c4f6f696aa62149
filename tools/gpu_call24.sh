#!/bin/bash
# round 2, leftover seconds: memcheck of the C++ facade smoke (incl. the grid classes), synccheck of the grid smoke
mkdir -p gpurun_out
g++ -std=c++17 -O1 -I include tests/cpp/facade_smoke.cpp -Lmatrel_b200 -lmatrel_b200 -Wl,-rpath,$PWD/matrel_b200 -o /tmp/facade_smoke || exit 1
g++ -std=c++17 -O1 -I include tests/cpp/grid_smoke.cpp -Lmatrel_b200 -lmatrel_b200 -Wl,-rpath,$PWD/matrel_b200 -o /tmp/grid_smoke || exit 1
timeout -k 3 25 compute-sanitizer --tool memcheck --error-exitcode 9 /tmp/facade_smoke > gpurun_out/sanitizer_memcheck_facade.log 2>&1; echo "memcheck facade_smoke rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_facade.log
timeout -k 3 25 compute-sanitizer --tool synccheck --error-exitcode 9 /tmp/grid_smoke 1 > gpurun_out/sanitizer_synccheck_gridops.log 2>&1; echo "synccheck grid_smoke rc=$?"; tail -2 gpurun_out/sanitizer_synccheck_gridops.log
