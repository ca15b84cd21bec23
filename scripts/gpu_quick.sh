#!/bin/bash
# Short gpurun call: bring-up tools, tensor-core parity tests, one ncu --set full capture.
set -u
mkdir -p gpurun_out
echo "== tcgen05 attention bring-up"; timeout 300 ./mt3_b200/csrc/tools/attn_tc_test 2>&1 | tail -14 | tee gpurun_out/attn_tc_test.log
echo "== pytest tensor-core tests"; timeout 900 python -m pytest tests -q -m gpu -s -k "tensor_core or graph_equivalence or inference_model" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_tc.log
echo "== ncu full: cluster decode GEMM"
MT3_TC_ATTENTION=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_dec_cluster --launch-skip 250 -c 4 \
   -o gpurun_out/prof_dec_gemm_cluster -f python bench.py --steps 1 --warmup 1 --dec-steps 4 --no-cpu-baseline --gemm-mode tf32x3 > gpurun_out/ncu_full_dec_gemm_cluster.log 2>&1
tail -1 gpurun_out/ncu_full_dec_gemm_cluster.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
