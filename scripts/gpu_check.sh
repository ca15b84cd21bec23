#!/bin/bash
# One gpurun call: bring-up tools, parity tests, smoke, bench lines (A/B switches), ncu launch list.
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tcgen05 bring-up tools"; timeout 300 ./mt3_b200/csrc/tools/gemm_tc_test quick 2>&1 | tail -2 | tee gpurun_out/gemm_tc_test.log
timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -3 | tee gpurun_out/attn_tc_test.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
run_bench () {  # name, extra args...
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 "$@" 2> gpurun_out/bench_$name.err | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-200
  grep -E "timed|e2e|cpu port|microbench" gpurun_out/bench_$name.err
}
run_bench tf32x3 --gemm-mode tf32x3
MT3_DEC_STREAMS=2 run_bench tf32x3_s2 --gemm-mode tf32x3 --no-cpu-baseline
MT3_DEC_STREAMS=4 run_bench tf32x3_s4 --gemm-mode tf32x3 --no-cpu-baseline
MT3_DEC_STREAMS=4 MT3_PDL=1 run_bench tf32x3_s4_pdl --gemm-mode tf32x3 --no-cpu-baseline
echo "== ncu launch list (tf32x3, 4 decode steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>&1 | tail -24 | tee gpurun_out/launch_summary.txt
