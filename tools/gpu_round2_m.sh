#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mgpu.py tests/test_gpu_abi_errors.py tests/test_cli.py -m gpu -q -x > gpurun_out/m_tests.log 2>&1
tail -3 gpurun_out/m_tests.log
timeout 600 python bench.py --no-cpu --no-largest > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/m_bench_n1.json").read().strip().splitlines()[-1])
print("N1 ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
PY
