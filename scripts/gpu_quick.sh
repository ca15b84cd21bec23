#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -7 | tee gpurun_out/attn_tc_test.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py --ref-budget-s 5 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
grep -E "timed|e2e |microbench" gpurun_out/bench_default.err | sed 's/.*kernel microbench: .*enc_qkv/enc_qkv/'
