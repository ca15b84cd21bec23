#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -12 | tee gpurun_out/attn_tc_test.log
echo "== pytest tensor core"; timeout 900 python -m pytest tests -q -m gpu -x -k "tensor_core_encoder or inference_model" 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
grep -E "timed|microbench" gpurun_out/bench_default.err
