#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu_all.log 2>&1; tail -8 gpurun_out/pytest_gpu_all.log
echo "== bench matrixfree level 3 (quick) and level 5"
timeout 600 python bench.py --workload matrixfree --sphere-level 3 --steps 3 --warmup 1 > gpurun_out/bench_mf3.log 2>&1; tail -1 gpurun_out/bench_mf3.log | cut -c1-700
timeout 900 python bench.py --workload matrixfree --sphere-level 5 --steps 2 --warmup 1 > gpurun_out/bench_mf5.log 2>&1; tail -1 gpurun_out/bench_mf5.log | cut -c1-900
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/bench_ours.log 2>&1; tail -1 gpurun_out/bench_ours.log | cut -c1-3000
echo "== ncu full (final regular kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 2 -c 1 -o gpurun_out/prof_grouped_v3 python tools/gpu_variants.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
