#!/bin/bash
# Trimmed multi-GPU check (gpurun --gpus N is charged N x the box time): the driver's launch line for bench.py,
# the sharded InferenceModel check and the long-form workload on N GPUs.  The full version is scripts/gpu_multi.sh.
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus_n$N.txt
echo "== bench --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-alt-kv 2> gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-400
grep -E "timed|e2e|broadcast|gather" gpurun_out/bench_n$N.err | head; tail -3 gpurun_out/bench_n$N.err
echo "== distributed transcribe (InferenceModel over NCCL)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
   scripts/dist_check.py 2>&1 | tail -6 | tee gpurun_out/dist_check_n$N.log
echo "== long-form workload (BASELINE configs[4]) on $N GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
   bench.py --workload longform --gpus $N --steps 3 --warmup 1 2> gpurun_out/bench_longform_n$N.err | tail -1 | tee gpurun_out/bench_longform_n$N.json | cut -c1-260
