#!/bin/bash
# One gpurun call: parity tests, then A/B bench lines of the decode-step switches and a decode-step timeline.
# Usage: scripts/gpu_ab.sh [tests|notests]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${1:-tests}" = "tests" ]; then
  echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/pytest_gpu.log
fi
run_bench () {  # name, env..., -- args
  local name=$1; shift
  echo "== bench $name"
  ( timeout 600 env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json )
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$name.json"))
    print("   %-14s ms/step %.1f  value %.1f  e2e %.1f  kv_f32 %s  job frac %.3f" % ("$name", d["ms_per_step"], d["value"], d["e2e"]["value"],
          d.get("ms_per_step_kv_f32"), d["roofline"]["job"]["frac"]))
    k = d["roofline"]["other_kernels_us_per_launch"]
    print("   " + ", ".join("%s=%.1f" % (a.replace("dec_", "").replace("attention", "attn"), b) for a, b in k.items() if a.startswith("dec_")))
except Exception as e:
    print("   $name FAILED", e)
    print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
for spec in ${AB_LIST:-default:MT3_X=0}; do
  run_bench "${spec%%:*}" "${spec#*:}"
done
echo "== decode-step timeline"
TRACE_POS=512 TRACE_ROWS=12 timeout 300 python scripts/trace_step.py > gpurun_out/trace_step.log 2>&1; tail -14 gpurun_out/trace_step.log
TRACE_KV=f32 TRACE_POS=512 TRACE_ROWS=0 timeout 300 python scripts/trace_step.py > gpurun_out/trace_step_kvf32.log 2>&1; tail -12 gpurun_out/trace_step_kvf32.log
