#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (run under gpurun); summaries go to gpurun_out/.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
run() {  # tool, tag, pytest -k expression
  timeout 600 compute-sanitizer --tool "$1" --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ozaki.py tests/test_gpu_ozaki2.py tests/test_gpu_tf32.py \
      -m gpu -x -q -k "$3" > "$OUT/sanitizer_$1_$2.log" 2>&1
  echo "== $1 $2: exit $? ; $(grep -E 'ERROR SUMMARY|passed|failed' "$OUT/sanitizer_$1_$2.log" | tr '\n' ' ')"
}
run memcheck dmma   "multiply_dense_vs_oracle and (131 or 300 or 40-40)"
run memcheck ozaki  "ozaki_multiply_vs_oracle and (131 or 300)"
run memcheck crt    "crt_multiply_vs_oracle and (131 or 300) or crt_is_bit_exact"
run memcheck tf32   "tf32x3_multiply and 300"
run memcheck sparse "sparse_blocks or sparse_op_sparse or aggregates or project"
run memcheck spsp   "sparse_times_sparse"
run racecheck dmma  "multiply_dense_vs_oracle and (131 or 256-256-256-128-0.5)"
run synccheck ozaki "ozaki_multiply_vs_oracle and 256-256-256"
run synccheck crt   "crt_multiply_vs_oracle and 256-256-256"
