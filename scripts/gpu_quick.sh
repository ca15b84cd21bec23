#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest decode"; timeout 900 python -m pytest tests -q -m gpu -x -k "variants or fused or decoder_teacher or graph_equivalence or tiny_golden or tensor_core_variants" 2>&1 | tail -4
TRACE_POS=512 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_c16.log | tail -12
for c in 1 0; do
MT3_DEC_CLUSTER16=$c timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c16_$c.err | tail -1 > gpurun_out/bench_c16_$c.json
echo "cluster16=$c: $(grep -E 'timed' gpurun_out/bench_c16_$c.err)"
done
