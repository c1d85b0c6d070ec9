#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/z_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/z_all_tests.log
tail -5 gpurun_out/z_all_tests.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/z_smoke.log 2>&1
tail -1 gpurun_out/z_smoke.log
timeout 900 python bench.py > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err
tail -c 300 gpurun_out/z_bench_n1.json
timeout 300 python bench.py --mesh s5m2 --scale 0.0005 --level -1 --no-cpu --no-largest > gpurun_out/z_bench_s5m2_ad_n1.json 2> gpurun_out/z_bench_s5m2_ad_n1.err
timeout 600 python bench.py --impl reference --mesh s5m2 --scale 0.0005 --level -1 --steps 1 --warmup 0 > gpurun_out/z_bench_s5m2_ad_ref.json 2> gpurun_out/z_bench_s5m2_ad_ref.err
tail -c 400 gpurun_out/z_bench_s5m2_ad_ref.json
