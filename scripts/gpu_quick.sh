#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
TRACE_POS=512 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_attn2.log | tail -13
MT3_DEC_INTERLEAVE=2 TRACE_POS=512 TRACE_ROWS=24 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_attn2_il2.log | head -30
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed|e2e |microbench" gpurun_out/bench_$name.err | head -3; tail -2 gpurun_out/bench_$name.err | grep -i -E "error|Traceback" 
}
run_bench default
MT3_DEC_INTERLEAVE=2 run_bench interleave2
MT3_DEC_INTERLEAVE=4 run_bench interleave4
