#!/bin/bash
# 2-GPU session (gpurun --gpus 2 -- bash tools/gpu_peer.sh): peer-store export test + bench.py under torchrun at N=2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
echo "== pytest peer export"
timeout 900 python -m pytest tests/test_gpu_peer_export.py tests/test_gpu_parity.py -m gpu -q -x -s -k "peer or host_buffer" --timeout 800 > gpurun_out/pytest_peer.log 2>&1; tail -25 gpurun_out/pytest_peer.log
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n$N.log 2>&1
tail -3 gpurun_out/bench_n$N.log | cut -c1-6000
