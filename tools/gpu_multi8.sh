#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L > gpurun_out/gpus.txt
echo "== bench lists N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | cut -c1-400
echo "== bench matrixfree N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload matrixfree --steps 3 --warmup 1 > gpurun_out/bench_mf_n$N.log 2>&1
grep '^{' gpurun_out/bench_mf_n$N.log | cut -c1-700
tail -3 gpurun_out/bench_mf_n$N.log | cut -c1-300
