#!/bin/bash
# One gpurun call: parity tests, smoke, bench line, ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tcgen05 gemm bring-up"; timeout 300 ./mt3_b200/csrc/tools/gemm_tc_test 2>&1 | tail -40 | tee gpurun_out/gemm_tc_test.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench simt"; timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_simt.err | tail -1 | tee gpurun_out/bench_simt.json; tail -12 gpurun_out/bench_simt.err
echo "== bench tf32x3"; timeout 600 python bench.py --steps 3 --warmup 3 --gemm-mode tf32x3 --no-cpu-baseline 2> gpurun_out/bench_tf32x3.err | tail -1 | tee gpurun_out/bench_tf32x3.json; tail -6 gpurun_out/bench_tf32x3.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
python scripts/summarize_launches.py gpurun_out/launches.csv 2>&1 | tail -30
