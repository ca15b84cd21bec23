#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest tensor-core decode variants"; timeout 900 python -m pytest tests -q -m gpu -x -k "tensor_core_variants or fused" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_dec_tc.log
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed" gpurun_out/bench_$name.err | head -1; tail -2 gpurun_out/bench_$name.err | grep -i -E "error|Traceback" 
}
export MT3_DEC_GEMM_MODE=1 MT3_DEC_TC=0
run_bench mma
MT3_DEC_INTERLEAVE=2 run_bench mma_il2
MT3_DEC_INTERLEAVE=4 run_bench mma_il4
MT3_DEC_INTERLEAVE=2 TRACE_POS=512 TRACE_ROWS=30 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_mma_il2.log | head -36
