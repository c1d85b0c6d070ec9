#!/bin/bash
# 8-GPU bench (gpurun --gpus 8 -- bash tools/gpu_n8.sh): default workload under torchrun, all export variants
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n8.log 2>&1
grep "^{" gpurun_out/bench_n8.log | cut -c1-300
