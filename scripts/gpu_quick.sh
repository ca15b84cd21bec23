#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
grep -E "timed|e2e |cpu port" gpurun_out/bench_default.err; cut -c1-260 gpurun_out/bench_default.json
