#!/bin/bash
# Full single-GPU session (run with: gpurun -- bash tools/gpu_session.sh): kernel timing, GPU test suite, bench (both arms),
# ncu launch list of the bench command and one ncu --set full capture of the regular-pair kernel.  Outputs in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 300 python tools/gpu_variants.py >> gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --durations=15 > gpurun_out/pytest_gpu_all.log 2>&1; tail -24 gpurun_out/pytest_gpu_all.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_ours.log 2>&1; tail -1 gpurun_out/bench_ours.log | cut -c1-600
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 2 -c 1 -o gpurun_out/prof_grouped_v7 python tools/gpu_variants.py > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
