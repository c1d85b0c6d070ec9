#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 0 1 0 1; do
I2_LEVEL1=$v timeout 300 python bench.py --mesh s5m2 --scale 0.0005 --level -1 --no-cpu --no-largest --no-e2e --steps 10 > gpurun_out/l_bench_s5m2_level1_$v.json 2> gpurun_out/l_bench_s5m2_level1_$v.err
python - gpurun_out/l_bench_s5m2_level1_$v.json $v <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("I2_LEVEL1=" + sys.argv[2], "ms", round(d["ms_per_step"], 4), "checksum", d["checksum_sum_abs_J"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_mgpu.py -m gpu -q -x > gpurun_out/l_tests.log 2>&1
tail -3 gpurun_out/l_tests.log
