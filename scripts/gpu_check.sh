#!/bin/bash
# One gpurun call: bring-up tool, parity tests, smoke, bench lines, ncu launch list + full captures.
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tcgen05 gemm bring-up"; timeout 300 ./mt3_b200/csrc/tools/gemm_tc_test 2>&1 | tail -30 | tee gpurun_out/gemm_tc_test.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
run_bench () {  # name, extra args...
  local name=$1; shift
  echo "== bench $name"
  timeout 600 python bench.py --steps 3 --warmup 3 "$@" 2> gpurun_out/bench_$name.err | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-260
  grep -E "timed|e2e|cpu port" gpurun_out/bench_$name.err
}
run_bench simt --no-cpu-baseline
MT3_PDL=1 run_bench simt_pdl --no-cpu-baseline
run_bench tf32x3 --gemm-mode tf32x3
MT3_PDL=1 run_bench tf32x3_pdl --gemm-mode tf32x3 --no-cpu-baseline
echo "== ncu launch list (tf32x3, 4 decode steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>&1 | tail -24 | tee gpurun_out/launch_summary.txt
echo "== ncu --set full: decode attention (roofline leg, cache length 512), decode GEMM, encoder GEMM + attention"
# attention launches before the roofline leg: 4 passes x 4 steps x 16 = 256 -> skip them, profile the debug launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_attention_bulk --launch-skip 262 -c 2 \
   -o gpurun_out/prof_dec_attention -f python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_full1.log 2>&1
prof () {  # name, kernel regex, launch-skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c $4 \
     -o gpurun_out/prof_$1 -f python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_full_$1.log 2>&1
  tail -1 gpurun_out/ncu_full_$1.log | cut -c1-200
}
prof dec_gemm sgemm_dec 250 4          # second pass: QKV / self-out / q / cross-out of layer 0..
prof enc_gemm gemm_tf32 60 6           # second pass: input projection + first layer's GEMMs
prof enc_attention enc_attention_tc 9 2
prof logmel logmel2048 1 1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
