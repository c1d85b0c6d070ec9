#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mgpu.py -m gpu -q -x > gpurun_out/h_tests.log 2>&1
tail -4 gpurun_out/h_tests.log
for g in 0 1; do for mesh in s5m s5m2; do
I2_GRAPHS=$g timeout 300 python bench.py --mesh $mesh --scale 0.0005 --level -1 --no-cpu --no-largest > gpurun_out/h_bench_${mesh}_ad_graphs$g.json 2> gpurun_out/h_bench_${mesh}_ad_graphs$g.err
python - gpurun_out/h_bench_${mesh}_ad_graphs$g.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "launches", d["gpu_launches"], "checksum", d["checksum_sum_abs_J"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done; done
