#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, ".")
from integrator2_b200.meshio import load_fixture, write_dat
write_dat("/tmp/Vint16k.dat", load_fixture("Vint16k"))
PY
(cd /tmp && "$GRAFT_REPO_ROOT/integrator2_b200/host/integrator2test3D" -f /tmp/Vint16k.dat -r 0 -c > "$GRAFT_REPO_ROOT/gpurun_out/g_cli_gpus1.txt" 2>&1)
grep -E "Time for" gpurun_out/g_cli_gpus1.txt
bash tools/gpu_sanitize2.sh > gpurun_out/g_sanitize2.txt 2>&1
cat gpurun_out/g_sanitize2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mgpu.py tests/test_cli.py -m gpu -q -x > gpurun_out/g_tests.log 2>&1
tail -4 gpurun_out/g_tests.log
