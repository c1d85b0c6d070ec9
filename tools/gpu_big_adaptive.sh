#!/bin/bash
# BASELINE.json configs[3] at full size (s5m2 refined twice = 125 280 triangles, 1.57e10 ordered pairs, automatic error control),
# list-free: gpurun [--gpus N] -- bash tools/gpu_big_adaptive.sh N
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
ARGS="bench.py --gpus $N --workload matrixfree --mf-mesh s5m2 --scale 0.0005 --sphere-level 2 --level -1 --steps 2 --warmup 3"
if [ "$N" = "1" ]; then timeout 600 python $ARGS > gpurun_out/bench_big_adaptive_n$N.log 2>&1
else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 $ARGS > gpurun_out/bench_big_adaptive_n$N.log 2>&1; fi
grep '^{' gpurun_out/bench_big_adaptive_n$N.log | cut -c1-1500 || tail -20 gpurun_out/bench_big_adaptive_n$N.log
