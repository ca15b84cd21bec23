#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
TRACE_POS=512 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_mma.log | tail -22
MT3_DEC_GEMM_MODE=0 TRACE_POS=512 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_fma.log | tail -12
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed|e2e " gpurun_out/bench_$name.err | head -2
}
run_bench default
MT3_DEC_GEMM_MODE=0 run_bench dec_fma
MT3_CLUSTER_POLICY=1 run_bench policy_spread
MT3_CLUSTER_POLICY=2 run_bench policy_lb
