#!/bin/bash
# Short gpurun call: decode-variant parity + A/B bench of the GEMM-chain decode path.
set -u
mkdir -p gpurun_out
echo "== attention tool"; timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -2
echo "== pytest decode variants"; timeout 900 python -m pytest tests -q -m gpu -s -k "variants or graph_equivalence or tensor_core" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_variants.log
run_bench () {
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/bench_$name.err | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-200
  grep -E "timed|e2e|microbench" gpurun_out/bench_$name.err; tail -3 gpurun_out/bench_$name.err | grep -v bench
}
MT3_DEC_CHAIN=1 run_bench chain
MT3_DEC_CHAIN=1 MT3_PDL=1 run_bench chain_pdl
run_bench default
echo "== ncu launch list chain"
MT3_DEC_CHAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_chain.csv \
   python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline > gpurun_out/ncu_bench_chain.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_chain.csv 2>&1 | tail -20 | tee gpurun_out/launch_summary_chain.txt
