#!/bin/bash
# compute-sanitizer passes over a small end-to-end run (G1: all classes, fixed + adaptive, matrix-free, classification)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
m = load_fixture("G1")
ctx = abi.Context(0)
ctx.set_mesh(m.vertices, m.cells)
lists = ctx.classify()
for cls in range(3):
    t = ctx.tasks_from_pairs(lists[cls])
    for level in (0, 2, 3, -1):
        r = ctx.integrate_class(cls, t, level)
        assert torch.isfinite(r["results"]).all()
ctx.apply_regular(0, m.n_cells)
cnt = ctx.host_prepare(m.vertices, m.cells)
ht = [torch.empty((n, 3), dtype=torch.int32, pin_memory=True) for n in cnt]
hr = [torch.empty((n, 3), dtype=torch.float64, pin_memory=True) for n in cnt]
ctx.host_run(-1, ht, hr)
print("sanitizer workload done")
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python /tmp/san.py > gpurun_out/sanitizer_$tool.out 2>&1
  tail -1 gpurun_out/sanitizer_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error" gpurun_out/sanitizer_$tool.log | head -5
done
echo "== new gpu tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cost_balanced" -s 2>&1 | tail -4
