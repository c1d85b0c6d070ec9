#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mgpu.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/k_tests.log 2>&1
tail -3 gpurun_out/k_tests.log
python tools/gpu_ab.py BASE=1 > gpurun_out/k_ab.log 2>&1; head -1 gpurun_out/k_ab.log
for cfg in "s5m 0.0005 -1" "s5m2 0.0005 -1"; do
set -- $cfg
timeout 300 python bench.py --mesh $1 --scale $2 --level $3 --no-cpu --no-largest --steps 10 > gpurun_out/k_bench_$1.json 2> gpurun_out/k_bench_$1.err
python - gpurun_out/k_bench_$1.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "checksum", d["checksum_sum_abs_J"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
