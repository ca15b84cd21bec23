#!/bin/bash
# Short gpurun call: decode-variant parity + A/B bench of the persistent decode kernel.
set -u
mkdir -p gpurun_out
echo "== pytest decode variants"; timeout 900 python -m pytest tests -q -m gpu -s -k "variants or graph_equivalence" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_variants.log
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 "$@" 2> gpurun_out/bench_$name.err | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-200
  grep -E "timed|e2e|microbench" gpurun_out/bench_$name.err; tail -3 gpurun_out/bench_$name.err | grep -v bench
}
MT3_DEC_MEGA=1 run_bench tf32x3_mega --gemm-mode tf32x3 --no-cpu-baseline
run_bench tf32x3 --gemm-mode tf32x3 --no-cpu-baseline
