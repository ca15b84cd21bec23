#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
TRACE_POS=512 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_fused.log | tail -14
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed|e2e " gpurun_out/bench_$name.err | head -2; tail -2 gpurun_out/bench_$name.err | grep -i -E "error|Traceback" 
}
run_bench default
MT3_DEC_FUSE=0 run_bench unfused
