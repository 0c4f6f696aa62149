mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki2.py tests/test_gpu_spmm.py -x -q -m gpu > gpurun_out/t_oz2b.log 2>&1; echo "oz2+spmm tests rc=$?"; tail -12 gpurun_out/t_oz2b.log
# 1-SM vs 2-SM int8 GEMM inside the engine (gemm_variant 2 disables the pairing)
timeout 300 python - > gpurun_out/oz2_variants.txt 2>&1 <<'P'
import sys; sys.path.insert(0, '.')
import matrel_b200 as mb
n, blk = 16384, 1024
with mb.MatfastSession(device=0) as s:
    A, B = s.rand(n, n, blk, 42), s.rand(n, n, blk, 43)
    s.set_option("time_kernels", 1)
    for variant, name in ((2, "1sm"), (-1, "2sm")):
        s.set_option("gemm_variant", variant)
        for rep in range(4):
            s.reset_stats()
            C = A.matrixMultiply(n, n, B, n, n, blk)
            st = s.stats()
            print(name, rep, "tc_gemm_ms %.3f  TOPS %.1f  total_ms %.3f" % (st["tc_gemm_ms_total"], st["tc_int8_ops"] / st["tc_gemm_ms_total"] / 1e9, st["gemm_ms_total"]))
            del C
P
cat gpurun_out/oz2_variants.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm2_kernel -c 1 -o gpurun_out/prof_spmm2_r02 python tools/bench_spmm.py 8192 1024 0.01 1 > gpurun_out/ncu_spmm2.log 2>&1; tail -3 gpurun_out/ncu_spmm2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki2_gemm_2sm -c 1 -o gpurun_out/prof_oz2sm_r02 python tools/run_multiply.py 8192 1024 1 > gpurun_out/ncu_oz2sm.log 2>&1; tail -3 gpurun_out/ncu_oz2sm.log
timeout 600 ncu --set full --clock-control none -k regex:"crt_kernel|residue_kernel" -c 3 -o gpurun_out/prof_oz2aux_r02 python tools/run_multiply.py 8192 1024 1 > gpurun_out/ncu_oz2aux.log 2>&1; tail -3 gpurun_out/ncu_oz2aux.log
