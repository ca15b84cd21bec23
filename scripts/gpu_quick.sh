#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step.log | tail -80
echo "== bench default (sanity after the tracing hooks)"
timeout 600 python bench.py --steps 3 --warmup 3 --ref-budget-s 5 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
grep -E "timed|e2e |microbench|cpu" gpurun_out/bench_default.err | head -6
