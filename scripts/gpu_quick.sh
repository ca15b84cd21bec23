#!/bin/bash
# Short gpurun call: A/B of selective programmatic dependent launch (MT3_PDL bit mask) in the decode graph.
set -u
mkdir -p gpurun_out
echo "== pytest graph/PDL equivalence"; timeout 900 python -m pytest tests -q -m gpu -k "graph_equivalence or variants" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_pdl.log
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
  grep -E "timed|e2e " gpurun_out/bench_$name.err | head -2
}
MT3_PDL=2 run_bench pdl2_attn
MT3_PDL=4 run_bench pdl4_gemm
MT3_PDL=6 run_bench pdl6_attn_gemm
MT3_PDL=1 run_bench pdl1_all
run_bench default
