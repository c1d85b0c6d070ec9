#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
for v in 0 1 2 3; do I2_VARIANT=$v timeout 300 python tools/gpu_variants.py >> gpurun_out/variants.log 2>&1; done
for mb in 3 5; do I2_MINBLOCKS=$mb I2_VARIANT=3 timeout 300 python tools/gpu_variants.py >> gpurun_out/variants.log 2>&1; done
cat gpurun_out/variants.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu_all.log 2>&1; tail -8 gpurun_out/pytest_gpu_all.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_ours.log 2>&1; tail -1 gpurun_out/bench_ours.log
