#!/bin/bash
# Multi-GPU session (gpurun --gpus N -- bash tools/gpu_multi.sh N): bench.py under torchrun (lists + matrix-free) and N=1 on the same box
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
tail -3 gpurun_out/bench_n$N.log | cut -c1-1200
echo "== bench matrixfree N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload matrixfree --steps 3 --warmup 1 > gpurun_out/bench_mf_n$N.log 2>&1
grep '^{' gpurun_out/bench_mf_n$N.log | cut -c1-400
echo "== bench N=1 (same box)"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n1_samebox.log 2>&1; tail -1 gpurun_out/bench_n1_samebox.log | cut -c1-400
