#!/bin/bash
# BASELINE.json configs[3]: s5m2.dat (scale 0.0005) with automatic error control, sharded by predicted cost
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --mesh s5m2 --scale 0.0005 --level -1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_s5m2_ad_n$N.log 2>&1
grep '^{' gpurun_out/bench_s5m2_ad_n$N.log | cut -c1-500; tail -3 gpurun_out/bench_s5m2_ad_n$N.log | cut -c1-300
timeout 600 python bench.py --mesh s5m2 --scale 0.0005 --level -1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_s5m2_ad_n1.log 2>&1
grep '^{' gpurun_out/bench_s5m2_ad_n1.log | cut -c1-1500
