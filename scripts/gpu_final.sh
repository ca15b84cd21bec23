#!/bin/bash
# Short verification of the committed state in one gpurun call: default bench line, smoke, the MT3_PF_ATTN A/B
# (scripts/ab_prefetch.py) and the whole `-m gpu` suite.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== bench"; timeout 200 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-240 gpurun_out/bench_default.json; grep -E "timed|e2e" gpurun_out/bench_default.err
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== A/B"; timeout 120 python scripts/ab_prefetch.py > gpurun_out/ab_prefetch.txt 2> gpurun_out/ab_prefetch.err; tail -9 gpurun_out/ab_prefetch.txt
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
