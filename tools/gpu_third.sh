#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu_all.log 2>&1; tail -12 gpurun_out/pytest_gpu_all.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_ours.log 2>&1; tail -2 gpurun_out/bench_ours.log
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -2 gpurun_out/bench_ref.log
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
