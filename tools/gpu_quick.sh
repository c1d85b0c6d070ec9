#!/bin/bash
# quick GPU check: kernel timing + full GPU test suite
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 300 python tools/gpu_variants.py >> gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --durations=12 > gpurun_out/pytest_gpu_all.log 2>&1; tail -22 gpurun_out/pytest_gpu_all.log
grep -h "median gpu" gpurun_out/pytest_gpu_all.log | head
