#!/bin/bash
# First GPU session of a round: reference golden dumps, GPU tests, variant timings, bench (both arms), ncu evidence.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "== golden"; tools/gpu_golden.sh > gpurun_out/golden.log 2>&1; tail -3 gpurun_out/golden.log
echo "== pytest gpu (no -x)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_all.log 2>&1; tail -30 gpurun_out/pytest_gpu_all.log
echo "== explore"
for mb in 3 4 5; do I2_MINBLOCKS=$mb timeout 600 python tools/gpu_explore.py time Vint16k >> gpurun_out/explore.log 2>&1; done
timeout 600 python tools/gpu_explore.py adaptive s5m 0.0005 >> gpurun_out/explore.log 2>&1
timeout 900 python tools/gpu_explore.py adaptive s5m2 0.0005 >> gpurun_out/explore.log 2>&1
cat gpurun_out/explore.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_ours.log 2>&1; tail -3 gpurun_out/bench_ours.log
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -3 gpurun_out/bench_ref.log
echo "== ncu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 2 -c 1 -o gpurun_out/prof_regular python tools/gpu_explore.py time Vint16k > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -40
