#!/bin/bash
# Multi-GPU check (run with gpurun --gpus N): the driver's launch line for bench.py, both arms.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus_n$N.txt
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-400
grep -E "timed|e2e|broadcast|gather" gpurun_out/bench_n$N.err | head; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --impl reference --gpus $N --steps 1 --warmup 1 2> gpurun_out/bench_ref_n$N.err | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-300
echo "== distributed transcribe (InferenceModel over NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
   scripts/dist_check.py 2>&1 | tail -6 | tee gpurun_out/dist_check_n$N.log
echo "== pytest multi-GPU long-form parity"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_multi_n$N.log
echo "== long-form workload (BASELINE configs[4]), 1 and $N GPUs"
timeout 600 python bench.py --workload longform --steps 3 --warmup 1 2> gpurun_out/bench_longform_n1.err | tail -1 | tee gpurun_out/bench_longform_n1.json | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
   bench.py --workload longform --gpus $N --steps 3 --warmup 1 2> gpurun_out/bench_longform_n$N.err | tail -1 | tee gpurun_out/bench_longform_n$N.json | cut -c1-260
