#!/bin/bash
set -u
mkdir -p gpurun_out
MT3_DEC_STREAMS=2 TRACE_POS=512 TRACE_ROWS=140 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_streams2.log | awk 'NR<=30 || (NR>=68 && NR<=92)' 
tail -14 gpurun_out/trace_step_streams2.log
echo "== bench default"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
grep -E "timed|e2e " gpurun_out/bench_default.err | head -2
