#!/bin/bash
# One gpurun call: bring-up tools, parity tests, smoke, bench line + reference arm, ncu launch list and
# ncu --set full captures of the kernels DESIGN.md quotes.  Everything lands in gpurun_out/.
# Usage: scripts/gpu_check.sh [tests|notests]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${1:-tests}" = "tests" ]; then
  echo "== tcgen05 bring-up tools"; timeout 300 ./mt3_b200/csrc/tools/gemm_tc_test quick 2>&1 | tail -2 | tee gpurun_out/gemm_tc_test.log
  timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -3 | tee gpurun_out/attn_tc_test.log
  echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
fi
run_bench () {  # name, extra args...
  local name=$1; shift
  echo "== bench $name"
  timeout 900 python bench.py "$@" 2> gpurun_out/bench_$name.err | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-200
  grep -E "timed|e2e|cpu port|microbench" gpurun_out/bench_$name.err
}
run_bench default --steps 5 --warmup 3
run_bench reference --impl reference --steps 2 --warmup 1
NCU_ARGS="--steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --no-alt-kv"
echo "== ncu launch list (4 decode steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches.csv \
   python bench.py $NCU_ARGS > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>&1 | tail -24 | tee gpurun_out/launch_summary.txt
echo "== ncu --set full"
prof () {  # name, kernel regex, launch-skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c $4 \
     -o gpurun_out/prof_$1 -f python bench.py $NCU_ARGS > gpurun_out/ncu_full_$1.log 2>&1
  tail -1 gpurun_out/ncu_full_$1.log | cut -c1-160
}
prof dec_attention dec_attention_bulk 400 2    # the roofline leg's launches at cache length 512 (24-bit K/V rows)
prof dec_gemm sgemm_dec_cluster 60 5           # decode-step GEMMs (single and fused dual launch)
prof enc_gemm gemm_tf32 45 3                   # encoder GEMMs (3xTF32)
prof enc_attention enc_attention_tc 9 2
prof logmel logmel2048 1 1
echo "== decode-step timeline"; TRACE_POS=512 timeout 300 python scripts/trace_step.py > gpurun_out/trace_step.log 2>&1; tail -12 gpurun_out/trace_step.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
