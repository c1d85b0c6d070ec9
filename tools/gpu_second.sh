#!/bin/bash
# GPU session: tests + variant timings (+ optional ncu of the grouped kernel)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu_all.log 2>&1; tail -25 gpurun_out/pytest_gpu_all.log
echo "== explore"; rm -f gpurun_out/explore.log
for mb in 4 3 5; do I2_MINBLOCKS=$mb timeout 600 python tools/gpu_explore.py time Vint16k >> gpurun_out/explore.log 2>&1; done
timeout 900 python tools/gpu_explore.py adaptive s5m2 0.0005 >> gpurun_out/explore.log 2>&1
cat gpurun_out/explore.log
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 2 -c 1 -o gpurun_out/prof_grouped python tools/gpu_explore.py time Vint16k > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
