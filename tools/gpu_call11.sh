mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_full_suite.log 2>&1; echo "full suite rc=$?"; tail -6 gpurun_out/t_full_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_d.json; tail -2 gpurun_out/bench_d.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_d.json 2> gpurun_out/bench_ref_d.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_d.json
