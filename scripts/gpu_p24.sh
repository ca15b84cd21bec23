#!/bin/bash
# One gpurun call for the 24-bit K/V row format: parity tests, bench line with all three formats, ncu of the p24 kernel.
set -u
mkdir -p gpurun_out
echo "== pytest (K/V formats)"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "kv_cache or p24 or decode_variants" 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/pytest_p24.log
echo "== bench default (fp16 headline, alt legs f32 / p24)"
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_p24run.err | tail -1 | tee gpurun_out/bench_p24run.json | cut -c1-300
grep -E "timed|e2e|microbench" gpurun_out/bench_p24run.err
echo "== bench --kv p24 (roofline leg of the p24 kernel)"
timeout 900 python bench.py --steps 3 --warmup 3 --kv p24 --no-alt-kv --no-cpu-baseline 2> gpurun_out/bench_kvp24.err | tail -1 | tee gpurun_out/bench_kvp24.json | cut -c1-300
echo "== ncu --set full p24 attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_attention_bulk --launch-skip 400 -c 2 \
   -o gpurun_out/prof_dec_attention_p24 -f python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --no-alt-kv --kv p24 > gpurun_out/ncu_full_p24.log 2>&1
tail -1 gpurun_out/ncu_full_p24.log | cut -c1-160
echo "== decode-step timeline (p24)"; TRACE_POS=512 TRACE_KV=p24 timeout 300 python scripts/trace_step.py > gpurun_out/trace_step_p24.log 2>&1; tail -12 gpurun_out/trace_step_p24.log
