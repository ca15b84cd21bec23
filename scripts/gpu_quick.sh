#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest variants"; timeout 900 python -m pytest tests -q -m gpu -x -k "variants or graph_equivalence" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_variants.log
MT3_DEC_INTERLEAVE=2 TRACE_POS=512 TRACE_ROWS=40 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_interleave2.log | head -48
tail -14 gpurun_out/trace_step_interleave2.log
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed|e2e " gpurun_out/bench_$name.err | head -2; tail -2 gpurun_out/bench_$name.err | grep -i -E "error|Traceback" 
}
MT3_DEC_INTERLEAVE=2 run_bench interleave2
MT3_DEC_INTERLEAVE=3 run_bench interleave3
MT3_DEC_INTERLEAVE=4 run_bench interleave4
run_bench default
